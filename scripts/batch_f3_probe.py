"""Batched LocalOtsu / removeLines over A4 pages resident in HBM: pages per second (diagnostic).  usage: python scripts/batch_f3_probe.py [pages=128]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, prlib_b200
kv = dict(a.split("=") for a in sys.argv[1:])
n = int(kv.get("pages", 128)); rows, cols = 3508, 2480
ctx = prlib_b200.Context(0)
step = (cols + 15) // 16 * 16
buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
out = torch.empty_like(buf)
torch.cuda.synchronize()
for name, fn in (("binarizeLocalOtsu", lambda: ctx.binarize_local_otsu_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, out.data_ptr(), step, rows * step)),
                 ("removeLines", lambda: ctx.remove_lines_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, out.data_ptr(), step, rows * step))):
    fn()
    l0 = ctx.launch_count()
    t0 = time.perf_counter(); r = fn(); dt = time.perf_counter() - t0
    extra = "" if r is None else f", rectangles per page {r[0].mean():.0f}, status ok {int((r[1] == 0).sum())}/{n}"
    print(f"{name}: {n} A4 pages in {dt*1e3:.1f} ms = {n/dt:.0f} pages/s ({ctx.launch_count() - l0} launches){extra}", flush=True)
