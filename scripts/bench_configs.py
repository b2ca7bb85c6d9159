"""Device-resident throughput of every BASELINE.json config (not only the headline one bench.py reports).
usage: python scripts/bench_configs.py [--quick] > profiles/r01_configs.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import prlib_b200
from prlib_b200 import capi

PEAK = 6558.4
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
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
    return e0.elapsed_time(e1) / steps, {k: v["ms"] / steps for k, v in t.items()}


def local(name, method, params, window, n, rows, cols, morph=0):
    buf, step = pages(n, rows, cols)
    rc, orow, ocol = ctx.output_shape(method, rows, cols, window)
    ostep = (ocol + 15) // 16 * 16
    out = torch.empty((n, orow, ostep), dtype=torch.uint8, device="cuda")
    ms, fam = timed(lambda: ctx.binarize_local_batch_dev(method, buf.data_ptr(), n, rows, cols, step, rows * step, window, params, morph,
                                                          out.data_ptr(), ostep, orow * ostep))
    h = window // 2; Hp, Wp = rows + 2 * h, cols + 2 * h
    k1 = rows * cols + 16 * Hp * Wp; k2 = 16 * Hp * Wp + 2 * orow * ocol
    alg = k1 + k2 + (16 * Hp * Wp if method == capi.WOLFJOLION else 0)
    pps = n / (ms / 1e3)
    return {"config": name, "pages": n, "rows": rows, "cols": cols, "window": window, "morph": morph, "ms_per_step": ms, "pages_per_sec": pps,
            "MP_per_sec": pps * rows * cols / 1e6, "algorithmic_bytes_per_page": alg, "achieved_GBs": alg * pps / 1e9,
            "frac_of_measured_peak": alg * pps / 1e9 / PEAK, "kernel_ms_per_step": fam}


def otsu(name, n, rows, cols, tiles):
    buf, step = pages(n, rows, cols)
    out = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
    thr = torch.zeros(n, dtype=torch.int32, device="cuda")
    if tiles:
        fn = lambda: ctx.otsu_tiles_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 64, 64, 255.0, out.data_ptr(), step, rows * step)
        alg = 2 * rows * cols
    else:
        fn = lambda: ctx.otsu_global_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 255.0, out.data_ptr(), step, rows * step, thr.data_ptr())
        alg = 3 * rows * cols
    ms, fam = timed(fn)
    pps = n / (ms / 1e3)
    return {"config": name, "pages": n, "rows": rows, "cols": cols, "ms_per_step": ms, "pages_per_sec": pps, "MP_per_sec": pps * rows * cols / 1e6,
            "algorithmic_bytes_per_page": alg, "achieved_GBs": alg * pps / 1e9, "frac_of_measured_peak": alg * pps / 1e9 / PEAK,
            "kernel_ms_per_step": fam}


A4 = (3508, 2480); A3 = (9921, 7016)
n2 = 64 if quick else 256
res = []
res.append(local("2: Sauvola w=15 k=0.2, A4", capi.SAUVOLA, (0.2,), 15, n2, *A4))
res.append(local("2: Niblack w=15 k=-0.2, A4", capi.NIBLACK, (-0.2,), 15, n2, *A4))
res.append(local("2: Wolf-Jolion w=15 k=0.5, A4", capi.WOLFJOLION, (0.5,), 15, n2, *A4))
res.append(local("3: NICK w=101 k=-0.1, A3-600", capi.NICK, (-0.1,), 101, 8 if quick else 32, *A3))
res.append(local("3: Feng w=101 defaults, A3-600", capi.FENG, (0.75, 0.2, 0.03, 2.0), 101, 8 if quick else 32, *A3))
res.append(otsu("4: Global Otsu, A4", 128 if quick else 1024, *A4, tiles=False))
res.append(otsu("4: 64x64-tile Otsu, A4", 128 if quick else 1024, *A4, tiles=True))
res.append(local("defaults: Sauvola w=101 k=0.01 morph=2, A4", capi.SAUVOLA, (0.01,), 101, n2, *A4, morph=2))
print(json.dumps({"peak_GBs": PEAK, "results": res}, indent=1))
