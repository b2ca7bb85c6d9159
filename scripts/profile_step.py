"""Runs the device-resident hot path a few times (for ncu / quick timing).  usage: profile_step.py [pages] [steps] [method] [window]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import prlib_b200
pages_n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
method = int(sys.argv[3]) if len(sys.argv) > 3 else 0
window = int(sys.argv[4]) if len(sys.argv) > 4 else 15
rows = int(sys.argv[5]) if len(sys.argv) > 5 else 3508
cols = int(sys.argv[6]) if len(sys.argv) > 6 else 2480
params = {0: (0.2,), 1: (-0.2,), 2: (0.5,), 3: (-0.1,), 4: (0.75, 0.2, 0.03, 2.0)}[method]
ctx = prlib_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
if os.environ.get("PRL_FUSED"):
    ctx.set_option("enable_fused", 1)
rc, orow, ocol = ctx.output_shape(method, rows, cols, window)
si = (cols + 15) // 16 * 16; so = (ocol + 15) // 16 * 16
pages = torch.empty((pages_n, rows, si), dtype=torch.uint8, device="cuda")
masks = torch.empty((pages_n, orow, so), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(pages.data_ptr(), pages_n, rows, cols, si, rows * si, 2024, 0)
torch.cuda.synchronize()
ctx.timing_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(steps):
    if i == 1:
        ctx.timing_reset(); e0.record()
    ctx.binarize_local_batch_dev(method, pages.data_ptr(), pages_n, rows, cols, si, rows * si, window, params, 0,
                                 masks.data_ptr(), so, orow * so)
e1.record(); torch.cuda.synchronize()
t = ctx.timing()
n = max(steps - 1, 1)
print({k: round(v["ms"] / v["launches"], 3) for k, v in t.items()}, "ms/step(events)", round(e0.elapsed_time(e1) / n, 3) if steps > 1 else None,
      "pages/s", round(pages_n * n / (e0.elapsed_time(e1) / 1e3)) if steps > 1 else None)
