"""kernel-1 band sweep (diagnostic): usage python scripts/k1_probe.py [pages=256]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prlib_b200
from prlib_b200 import capi
kv = dict(a.split("=") for a in sys.argv[1:])
n = int(kv.get("pages", 256)); rows, cols = int(kv.get("rows", 3508)), int(kv.get("cols", 2480)); window = int(kv.get("window", 15))
ctx = prlib_b200.Context(0)
ctx.set_option("enable_fused", 0)      # kernel 1 + kernel 2, whatever the window
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
step = (cols + 15) // 16 * 16
buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
rc, orow, ocol = ctx.output_shape(0, rows, cols, window)
ostep = (ocol + 15) // 16 * 16
out = torch.empty((n, orow, ostep), dtype=torch.uint8, device="cuda")
ref = None
for bands in [int(b) for b in kv.get("bands", "0,1,2,3,4,5,6,7,8,10,12,16").split(",")]:
    ctx.set_option("k1_bands", bands)
    f = lambda: ctx.binarize_local_batch_dev(0, buf.data_ptr(), n, rows, cols, step, rows * step, window, (0.2,), 0, out.data_ptr(), ostep, orow * ostep)
    for _ in range(2): f()
    torch.cuda.synchronize(); ctx.timing_reset(); ctx.timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): f()
    e1.record(stream)
    torch.cuda.synchronize(); t = ctx.timing(); ctx.timing_enable(False)
    h = int(out[:4].to(torch.int64).sum())
    if ref is None: ref = h
    print(json.dumps({"bands": bands, "step_ms": round(e0.elapsed_time(e1) / 5, 3), **{k: round(v["ms"] / 5, 3) for k, v in t.items()}, "same": h == ref}), flush=True)
