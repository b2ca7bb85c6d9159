"""64x64-tile Otsu timing vs prefetch distance.  usage: tile_bench.py [pages] [pf ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import prlib_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pfs = [int(a) for a in sys.argv[2:]] or [0, 2, 4, 6, 8, 12]
rows, cols = 3508, 2480
ctx = prlib_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
step = (cols + 15) // 16 * 16
buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
out = torch.empty_like(buf)
ref = None
for pf in pfs:
    ctx.set_option("tile_prefetch", pf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(6):
        if i == 2: e0.record()
        ctx.otsu_tiles_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 64, 64, 255.0, out.data_ptr(), step, rows * step)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    if ref is None: ref = out.clone()
    print(f"pf={pf:2d}  {ms:7.3f} ms / {n} pages  = {2 * rows * cols * n / ms / 1e6:7.1f} GB/s (2 B/px)  same={bool(torch.equal(ref, out))}", flush=True)
