"""64x64-tile Otsu timing: the three kernels (set_option("tiles_legacy", v): 3 = bulk-copy ring, 2 = lane-per-tile with
register-staged loads, 1 = round-1 warp-batched kernel, 0 = the library's choice) on the same pages, outputs compared.  usage: tile_bench.py [pages] [tw th]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import prlib_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
tw, th = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 64)
rows, cols = 3508, 2480
ctx = prlib_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
step = (cols + 15) // 16 * 16
buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
ref = None
for variant in (1, 2, 3, 2, 3, 0):
    ctx.set_option("tiles_legacy", variant)
    out = torch.zeros_like(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(7):
        if i == 2: e0.record()
        ctx.otsu_tiles_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, tw, th, 255.0, out.data_ptr(), step, rows * step)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if ref is None: ref = out.clone()
    print(f"tiles_legacy={variant}  {tw}x{th}  {ms:7.3f} ms / {n} pages  = {2 * rows * cols * n / ms / 1e6:7.1f} GB/s (2 B/px)  same={bool(torch.equal(ref, out))}", flush=True)
ctx.set_option("tiles_legacy", 0)
