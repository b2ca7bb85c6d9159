"""kernel-2 timing probes (diagnostic): usage python scripts/k2_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prlib_b200
from prlib_b200 import capi
ctx = prlib_b200.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n, rows, cols = 256, 3508, 2480
step = (cols + 15) // 16 * 16
buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
def run(tag, method, params, window, **opts):
    for k, v in opts.items(): ctx.set_option(k, v)
    rc, orow, ocol = ctx.output_shape(method, rows, cols, window)
    ostep = (ocol + 15) // 16 * 16
    out = torch.empty((n, orow, ostep), dtype=torch.uint8, device="cuda")
    f = lambda: ctx.binarize_local_batch_dev(method, buf.data_ptr(), n, rows, cols, step, rows * step, window, params, 0, out.data_ptr(), ostep, orow * ostep)
    for _ in range(2): f()
    torch.cuda.synchronize(); ctx.timing_reset(); ctx.timing_enable(True)
    for _ in range(5): f()
    torch.cuda.synchronize(); t = ctx.timing(); ctx.timing_enable(False)
    print(json.dumps({"tag": tag, **{k: round(v["ms"] / 5, 3) for k, v in t.items()}, "white": round(float((out[:4, :, :ocol] == 255).float().mean()), 5)}), flush=True)
    for k in opts: ctx.set_option(k, 0)
run("sauvola ns3", capi.SAUVOLA, (0.2,), 15)
run("sauvola ns2", capi.SAUVOLA, (0.2,), 15, thr_stages=2)
run("sauvola ns2 bands2", capi.SAUVOLA, (0.2,), 15, thr_stages=2, k1_bands=2)
run("sauvola ns2 bands3", capi.SAUVOLA, (0.2,), 15, thr_stages=2, k1_bands=3)
run("sauvola ns2 bands5", capi.SAUVOLA, (0.2,), 15, thr_stages=2, k1_bands=5)
run("niblack ns2", capi.NIBLACK, (-0.2,), 15, thr_stages=2)
run("wj ns2", capi.WOLFJOLION, (0.5,), 15, thr_stages=2)
run("sauvola w=31 ns2", capi.SAUVOLA, (0.2,), 31, thr_stages=2)
run("sauvola w=101 ns3", capi.SAUVOLA, (0.01,), 101)
run("sauvola w=101 ns2", capi.SAUVOLA, (0.01,), 101, thr_stages=2)
