import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, prlib_b200
from prlib_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rows, cols = 3508, 2480
ctx = prlib_b200.Context(0)
d = torch.empty((n, rows, cols), dtype=torch.uint8, device="cuda")
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.synth_pages_dev(d.data_ptr(), n, rows, cols, cols, rows * cols, 2024, 0); torch.cuda.synchronize()
hp = torch.empty((n, rows, cols), dtype=torch.uint8).pin_memory(); hp.copy_(d)
hm = torch.empty((n, rows - 1, cols - 1), dtype=torch.uint8).pin_memory()
if len(sys.argv) > 2:
    prlib_b200.set_global_option('batch_chunk_pages', int(sys.argv[2]))
for i in range(4):
    t0 = time.perf_counter()
    prlib_b200.binarize_batch(hp.numpy(), capi.SAUVOLA, 15, (0.2,), 0, devices=[0], out=hm.numpy())
    dt = time.perf_counter() - t0
    print(f"n={n}: {1e3*dt:.1f} ms  {n/dt:.0f} pages/s  ({n*rows*cols*2/dt/1e9:.1f} GB/s both ways)")
