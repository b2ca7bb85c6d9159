"""Global Otsu over 1024 A4 pages: the two-lane overlap of histogram and apply passes, group size sweep (diagnostic)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prlib_b200
ctx = prlib_b200.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n, rows, cols = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 3508, 2480
step = (cols + 15) // 16 * 16
buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
out = torch.empty_like(buf); thr = torch.zeros(n, dtype=torch.int32, device="cuda")
ref = None
for group in (0, -1, 16, 32, 64, 128, 256):
    if group > 0 and 4 * group > n: continue
    ctx.set_option("otsu_group", group)
    f = lambda: ctx.otsu_global_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 255.0, out.data_ptr(), step, rows * step, thr.data_ptr())
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): f()
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    h = (int(thr.sum()), int(out[::97].to(torch.int64).sum()))
    if ref is None: ref = h
    print(json.dumps({"group": group, "pages": n, "ms": round(ms, 3), "frac_of_6558": round(3 * rows * cols * n / (ms / 1e3) / 1e9 / 6558.4, 3), "same": h == ref}), flush=True)
