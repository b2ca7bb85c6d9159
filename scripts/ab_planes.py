"""A/B of the two plane layouts (compact low-word planes + anchors vs full int64 planes) on the BASELINE configs.
usage: python scripts/ab_planes.py [--quick]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import prlib_b200
from prlib_b200 import capi

quick = "--quick" in sys.argv
ctx = prlib_b200.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)


def pages(n, rows, cols):
    step = (cols + 15) // 16 * 16
    buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
    ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
    torch.cuda.synchronize()
    return buf, step


def timed(fn, steps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ctx.timing_reset(); ctx.timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps): fn()
    e1.record(stream); torch.cuda.synchronize()
    t = ctx.timing(); ctx.timing_enable(False)
    return e0.elapsed_time(e1) / steps, {k: round(v["ms"] / steps, 3) for k, v in t.items()}


def run(name, method, params, window, n, rows, cols, morph=0):
    buf, step = pages(n, rows, cols)
    rc, orow, ocol = ctx.output_shape(method, rows, cols, window)
    ostep = (ocol + 15) // 16 * 16
    outs = {}
    for mode in ("int64", "compact-noTMA-K2", "compact"):
        ctx.set_option("disable_compact", 1 if mode == "int64" else 0)
        ctx.set_option("thr_no_tma", 1 if mode == "compact-noTMA-K2" else 0)
        out = torch.empty((n, orow, ostep), dtype=torch.uint8, device="cuda")
        ms, fam = timed(lambda: ctx.binarize_local_batch_dev(method, buf.data_ptr(), n, rows, cols, step, rows * step, window, params, morph,
                                                              out.data_ptr(), ostep, orow * ostep))
        outs[mode] = out[:, :, :ocol].clone()
        print(json.dumps({"config": name, "planes": mode, "ms_per_step": round(ms, 3), "pages_per_sec": round(n / ms * 1e3, 1), "kernels_ms": fam}), flush=True)
    print(json.dumps({"config": name, "masks_equal": bool(torch.equal(outs["int64"], outs["compact"]) and torch.equal(outs["int64"], outs["compact-noTMA-K2"]))}), flush=True)


A4 = (3508, 2480); A3 = (9921, 7016)
n2 = 64 if quick else 256
run("Sauvola w=15 A4", capi.SAUVOLA, (0.2,), 15, n2, *A4)
run("Niblack w=15 A4", capi.NIBLACK, (-0.2,), 15, n2, *A4)
run("Wolf-Jolion w=15 A4", capi.WOLFJOLION, (0.5,), 15, n2, *A4)
run("NICK w=101 A3-600", capi.NICK, (-0.1,), 101, 8 if quick else 32, *A3)
run("Feng w=101 A3-600", capi.FENG, (0.75, 0.2, 0.03, 2.0), 101, 8 if quick else 32, *A3)
run("Sauvola w=101 morph2 A4", capi.SAUVOLA, (0.01,), 101, n2, *A4, morph=2)
run("Sauvola w=15 A4 x8 pages (latency mode)", capi.SAUVOLA, (0.2,), 15, 8, *A4)
