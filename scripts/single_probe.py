"""Where the single-image call spends its time (diagnostic): pageable / pinned copies of one A4 page each way, the kernels alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import prlib_b200
from prlib_b200 import capi
from oracle import c_oracle as CO
page = CO.synth_page(0)
ctx = prlib_b200.Context(0)
def med(f, n=30):
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e3 * sorted(ts)[n // 2]
d = torch.empty((3508, 2480), dtype=torch.uint8, device="cuda")
hp = torch.from_numpy(page)                     # pageable
hpin = torch.from_numpy(page).pin_memory()
def h2d_pageable(): d.copy_(hp); torch.cuda.synchronize()
def h2d_pinned(): d.copy_(hpin, non_blocking=True); torch.cuda.synchronize()
out_pageable = torch.empty((3507, 2479), dtype=torch.uint8)
out_pinned = torch.empty((3507, 2479), dtype=torch.uint8).pin_memory()
dm = torch.empty((3507, 2479), dtype=torch.uint8, device="cuda")
def d2h_pageable(): out_pageable.copy_(dm); torch.cuda.synchronize()
def d2h_pinned(): out_pinned.copy_(dm, non_blocking=True); torch.cuda.synchronize()
print("H2D pageable %.3f ms, pinned %.3f ms; D2H pageable %.3f ms, pinned %.3f ms" % (med(h2d_pageable), med(h2d_pinned), med(d2h_pageable), med(d2h_pinned)))
step = 2480
ostep = 2479
dst = torch.empty((3507, ostep), dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
for fused in (1, 0):
    ctx.set_option("enable_fused", fused)
    def kern():
        ctx.binarize_local_batch_dev(0, d.data_ptr(), 1, 3508, 2480, step, 3508 * step, 15, (0.2,), 0, dst.data_ptr(), ostep, 3507 * ostep)
        torch.cuda.synchronize()
    print("kernels only (1 page, fused=%d): %.3f ms" % (fused, med(kern)))
ctx.set_stream(None)
for fused in (1, 0):
    ctx.set_option("enable_fused", fused)
    print("whole call, pageable numpy (fused=%d): %.3f ms" % (fused, med(lambda: ctx.binarize_local(page, 0, 15, (0.2,), 0))))
    print("whole call w=101 morph 2 (fused=%d): %.3f ms" % (fused, med(lambda: ctx.binarize_local(page, 0, 101, (0.01,), 2))))
