"""Device-resident throughput of every BASELINE.json config (not only the headline one bench.py reports).
usage: python scripts/bench_configs.py [--quick] > profiles/r02_configs.json
bench.py imports run_all() for the `other_configs` key of its N=1 line (quick sizes)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_all(ctx_device=0, quick=False):
    import torch
    import prlib_b200
    from prlib_b200 import capi

    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    dev = torch.device(f"cuda:{ctx_device}")
    ctx = prlib_b200.Context(ctx_device)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)

    def pages(n, rows, cols):
        step = (cols + 15) // 16 * 16
        buf = torch.empty((n, rows, step), dtype=torch.uint8, device=dev)
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
        return e0.elapsed_time(e1) / steps, {k: round(v["ms"] / steps, 4) for k, v in t.items()}

    def anchor_rows(Hp, Wp):
        for sh in (3, 2, 1, 0):
            if (((1 << sh) - 1) * Wp + 3.0 * Hp) * 65025.0 < 4294967296.0:
                return (Hp + (1 << sh) - 1) >> sh
        return Hp

    def local(name, method, params, window, n, rows, cols, morph=0, fused=0):
        buf, step = pages(n, rows, cols)
        rc, orow, ocol = ctx.output_shape(method, rows, cols, window)
        ostep = (ocol + 15) // 16 * 16
        out = torch.empty((n, orow, ostep), dtype=torch.uint8, device=dev)
        ctx.set_option("enable_fused", fused)
        ms, fam = timed(lambda: ctx.binarize_local_batch_dev(method, buf.data_ptr(), n, rows, cols, step, rows * step, window, params, morph,
                                                              out.data_ptr(), ostep, orow * ostep))
        ctx.set_option("enable_fused", 1)
        h = window // 2; Hp, Wp = rows + 2 * h, cols + 2 * h
        wj = method == capi.WOLFJOLION
        alg64 = rows * cols + 16 * Hp * Wp + 16 * Hp * Wp + 2 * orow * ocol + (16 * Hp * Wp if wj else 0)
        algc = rows * cols + 8 * Hp * Wp + 2 * anchor_rows(Hp, Wp) * Wp + 8 * Hp * Wp + 2 * orow * ocol + (8 * Hp * Wp if wj else 0)
        pps = n / (ms / 1e3)
        white = float((out[:2, :, :ocol] == 255).float().mean())
        del buf, out
        return {"config": name, "pages": n, "rows": rows, "cols": cols, "window": window, "morph": morph, "ms_per_step": round(ms, 4),
                "pages_per_sec": round(pps, 1), "MP_per_sec": round(pps * rows * cols / 1e6, 1),
                "algorithmic_bytes_per_page": algc, "achieved_GBs": round(algc * pps / 1e9, 1), "frac_of_measured_peak": round(algc * pps / 1e9 / peak, 4),
                "frac_int64_model": round(alg64 * pps / 1e9 / peak, 4), "white_fraction": round(white, 5), "kernel_ms_per_step": fam}

    def otsu(name, n, rows, cols, tiles):
        buf, step = pages(n, rows, cols)
        out = torch.empty((n, rows, step), dtype=torch.uint8, device=dev)
        thr = torch.zeros(n, dtype=torch.int32, device=dev)
        if tiles:
            fn = lambda: ctx.otsu_tiles_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 64, 64, 255.0, out.data_ptr(), step, rows * step)
            alg = 2 * rows * cols
        else:
            fn = lambda: ctx.otsu_global_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 255.0, out.data_ptr(), step, rows * step, thr.data_ptr())
            alg = 3 * rows * cols
        ms, fam = timed(fn)
        pps = n / (ms / 1e3)
        del buf, out
        return {"config": name, "pages": n, "rows": rows, "cols": cols, "ms_per_step": round(ms, 4), "pages_per_sec": round(pps, 1),
                "MP_per_sec": round(pps * rows * cols / 1e6, 1), "algorithmic_bytes_per_page": alg, "achieved_GBs": round(alg * pps / 1e9, 1),
                "frac_of_measured_peak": round(alg * pps / 1e9 / peak, 4), "kernel_ms_per_step": fam}

    A4 = (3508, 2480); A3 = (9921, 7016)
    n2 = 64 if quick else 256
    n3 = 8 if quick else 32
    n4 = 128 if quick else 1024
    res = []
    if not quick:
        res.append(local("2: Sauvola w=15 k=0.2, A4", capi.SAUVOLA, (0.2,), 15, n2, *A4))
    res.append(local("2: Niblack w=15 k=-0.2, A4", capi.NIBLACK, (-0.2,), 15, n2, *A4))
    res.append(local("2: Wolf-Jolion w=15 k=0.5, A4", capi.WOLFJOLION, (0.5,), 15, n2, *A4))
    res.append(local("3: NICK w=101 k=-0.1, A3-600", capi.NICK, (-0.1,), 101, n3, *A3))
    res.append(local("3: Feng w=101 defaults, A3-600", capi.FENG, (0.75, 0.2, 0.03, 2.0), 101, n3, *A3))
    res.append(otsu("4: Global Otsu, A4", n4, *A4, tiles=False))
    res.append(otsu("4: 64x64-tile Otsu, A4", n4, *A4, tiles=True))
    res.append(local("defaults: Sauvola w=101 k=0.01 morph=2, A4", capi.SAUVOLA, (0.01,), 101, n2, *A4, morph=2))
    torch.cuda.synchronize()
    ctx.close()
    return {"peak_GBs": peak, "bytes_model": "local methods: compact planes (frac_int64_model = SURVEY 8(d) int64 planes); Global Otsu 3*H*W; tile Otsu 2*H*W",
            "sizes": "quick" if quick else "BASELINE", "results": res}


if __name__ == "__main__":
    print(json.dumps(run_all(0, "--quick" in sys.argv), indent=1))
