"""fused default path vs two-kernel path over batch sizes (diagnostic): python scripts/path_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prlib_b200
ctx = prlib_b200.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
rows, cols, window = 3508, 2480, 15
step = (cols + 15) // 16 * 16
for n in (1, 2, 4, 8, 32, 64, 256):
    buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
    ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
    rc, orow, ocol = ctx.output_shape(0, rows, cols, window)
    ostep = (ocol + 15) // 16 * 16
    out = torch.empty((n, orow, ostep), dtype=torch.uint8, device="cuda")
    res = {"pages": n}
    for fused in (1, 0):
        ctx.set_option("enable_fused", fused)
        f = lambda: ctx.binarize_local_batch_dev(0, buf.data_ptr(), n, rows, cols, step, rows * step, window, (0.2,), 0, out.data_ptr(), ostep, orow * ostep)
        for _ in range(3): f()
        torch.cuda.synchronize(); ctx.timing_reset(); ctx.timing_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): f()
        e1.record(stream); torch.cuda.synchronize(); t = ctx.timing(); ctx.timing_enable(False)
        res["fused" if fused else "two_kernel"] = {"ms": round(e0.elapsed_time(e1) / 10, 4), **{k: round(v["ms"] / 10, 4) for k, v in t.items()}}
    print(json.dumps(res), flush=True)
    del buf, out
