"""what kernel 2's exact path costs (diagnostic; dbg_skip_exact gives WRONG masks): python scripts/k2_skip_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prlib_b200
from prlib_b200 import capi
ctx = prlib_b200.Context(0)
ctx.set_option("enable_fused", 0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n, rows, cols, window = 256, 3508, 2480, 15
step = (cols + 15) // 16 * 16
buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
rc, orow, ocol = ctx.output_shape(0, rows, cols, window)
ostep = (ocol + 15) // 16 * 16
out = torch.empty((n, orow, ostep), dtype=torch.uint8, device="cuda")
for skip in (0, 1, 0, 1):
    for stages in (2, 3):
        ctx.set_option("dbg_skip_exact", skip); ctx.set_option("thr_stages", stages)
        f = lambda: ctx.binarize_local_batch_dev(0, buf.data_ptr(), n, rows, cols, step, rows * step, window, (0.2,), 0, out.data_ptr(), ostep, orow * ostep)
        for _ in range(3): f()
        torch.cuda.synchronize(); ctx.timing_reset(); ctx.timing_enable(True)
        for _ in range(10): f()
        torch.cuda.synchronize(); t = ctx.timing(); ctx.timing_enable(False)
        print(json.dumps({"skip_exact": skip, "stages": stages, **{k: round(v["ms"] / 10, 3) for k, v in t.items()}}), flush=True)
