"""kernel 1 + kernel 2 per method over 256 A4 pages, two-kernel path (diagnostic): python scripts/k2_ab.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prlib_b200
from prlib_b200 import capi
ctx = prlib_b200.Context(0)
ctx.set_option("enable_fused", 0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
def run(tag, method, params, window, n=256, rows=3508, cols=2480):
    step = (cols + 15) // 16 * 16
    buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
    ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
    rc, orow, ocol = ctx.output_shape(method, rows, cols, window)
    ostep = (ocol + 15) // 16 * 16
    out = torch.empty((n, orow, ostep), dtype=torch.uint8, device="cuda")
    f = lambda: ctx.binarize_local_batch_dev(method, buf.data_ptr(), n, rows, cols, step, rows * step, window, params, 0, out.data_ptr(), ostep, orow * ostep)
    for _ in range(3): f()
    torch.cuda.synchronize(); ctx.timing_reset(); ctx.timing_enable(True)
    for _ in range(10): f()
    torch.cuda.synchronize(); t = ctx.timing(); ctx.timing_enable(False)
    import hashlib
    h = hashlib.sha1(out[:2, :, :ocol].contiguous().cpu().numpy().tobytes()).hexdigest()[:12]
    print(json.dumps({"tag": tag, **{k: round(v["ms"] / 10, 3) for k, v in t.items()}, "sha_p01": h}), flush=True)
    del buf, out
run("sauvola w15", capi.SAUVOLA, (0.2,), 15)
run("niblack w15", capi.NIBLACK, (-0.2,), 15)
run("wj w15", capi.WOLFJOLION, (0.5,), 15)
run("nick w15", capi.NICK, (-0.1,), 15)
run("feng w15", capi.FENG, (0.12, 0.25, 0.04, 2.0), 15)
run("sauvola w101 A4", capi.SAUVOLA, (0.01,), 101)
run("nick w101 A3-600", capi.NICK, (-0.1,), 101, n=32, rows=9921, cols=7016)
