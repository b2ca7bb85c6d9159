"""One method over a device-resident batch, a few steps: the command ncu wraps.
usage: python scripts/prof_step.py [pages=64] [steps=3] [method=0] [window=15] [rows=3508] [cols=2480] [opt=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import prlib_b200

kv = dict(a.split("=") for a in sys.argv[1:])
n = int(kv.pop("pages", 64)); steps = int(kv.pop("steps", 3)); method = int(kv.pop("method", 0)); window = int(kv.pop("window", 15))
rows = int(kv.pop("rows", 3508)); cols = int(kv.pop("cols", 2480)); morph = int(kv.pop("morph", 0))
params = {0: (0.2,), 1: (-0.2,), 2: (0.5,), 3: (-0.1,), 4: (0.75, 0.2, 0.03, 2.0)}[method]
ctx = prlib_b200.Context(0)
for k, v in kv.items():
    ctx.set_option(k, int(v))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
step = (cols + 15) // 16 * 16
buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
rc, orow, ocol = ctx.output_shape(method, rows, cols, window)
ostep = (ocol + 15) // 16 * 16
out = torch.empty((n, orow, ostep), dtype=torch.uint8, device="cuda")
for _ in range(steps):
    ctx.binarize_local_batch_dev(method, buf.data_ptr(), n, rows, cols, step, rows * step, window, params, morph, out.data_ptr(), ostep, orow * ostep)
torch.cuda.synchronize()
print("ok", n, steps, method, window)
