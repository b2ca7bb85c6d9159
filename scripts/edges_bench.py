"""Edge front-end of prl::binarizeLocalOtsu on one A4 page: device time per kernel family, host-call latency, cv2 time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import prlib_b200
from oracle import c_oracle as CO, prl_oracle as O
page = CO.synth_page(0)
rows, cols = page.shape
ctx = prlib_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
d = torch.from_numpy(page).cuda()
out = torch.empty_like(d)
L = ctx._L
for i in range(4):
    if i == 1:
        ctx.timing_reset(); ctx.timing_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    ctx._check(L.prl_cuda_canny_edge_detection_dev(ctx._h, d.data_ptr(), rows, cols, cols, 19, 0.15, 0.01, 1, 3, out.data_ptr(), cols))
e1.record(); torch.cuda.synchronize()
t = ctx.timing(); ctx.timing_enable(False)
print("device-resident chain:", round(e0.elapsed_time(e1) / 3, 3), "ms per A4 page;", {k: (round(v["ms"] / 3, 3), v["launches"] // 3) for k, v in t.items()})
ctx.set_stream(None)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); e = ctx.canny_edge_detection(page, 19, 0.15, 0.01, 1, 3); ts.append(time.perf_counter() - t0)
t0 = time.perf_counter(); ref = O.local_otsu_edges(page); tc = time.perf_counter() - t0
print(f"host call: {1e3 * min(ts):.2f} ms; cv2 (1 thread): {1e3 * tc:.1f} ms; equal: {np.array_equal(e, ref)}; edge pixels {int((e > 0).sum())}")
ts = []
for _ in range(3):
    t0 = time.perf_counter(); m = prlib_b200.binarizeLocalOtsu(page); ts.append(time.perf_counter() - t0)
t0 = time.perf_counter(); mr = O.binarizeLocalOtsu(page); tc = time.perf_counter() - t0
print(f"binarizeLocalOtsu: GPU path {1e3 * min(ts):.1f} ms (host image in, host mask out, one call), cv2 {1e3 * tc:.1f} ms, equal {np.array_equal(m, mr)}")
