#!/usr/bin/env python
"""bench.py -- throughput of the local-statistics binarization hot path on B200.

Metric (BASELINE.json): megapixels/s (and pages/s) of Sauvola w=15, k=0.2, morph 0 over synthetic
A4 300-dpi pages (2480 x 3508 u8, synthpage-v2 seed 2024), plus the fraction of the HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pages P] [--impl ours|reference]

One "step" = one pass of the hot path over a batch of P pages PER GPU (weak scaling; pages are
independent, no data-path collective).
  value              : whole-job MP/s of the integral-image pipeline north_star names (kernel 1 -> kernel 2, S/Q planes
                       through HBM), pages already resident in HBM (device-timed, max over ranks)
  value_default_path : the same call with the library's defaults (windows <= 31 take the fused small-window kernel)
  e2e                : the same metric through the host C-ABI call prl_cuda_binarize_batch with pinned HOST buffers,
                       H2D and D2H inside the timed region
  roofline / cpu_baseline / e2e_single_call / e2e_dispatcher / pcie : see DESIGN.md section "Measurement"
--impl reference times the reference's OWN C++ (oracle/_ref: binarizeSauvola.cpp compiled unmodified against the cv::
facade, OpenCV primitives executed by the cv2 wheel) on the host cores; the op-for-op port stands in if _ref is absent.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS = 3508, 2480          # A4 at 300 dpi
WINDOW, K_COEF = 15, 0.2
SEED = 2024
METRIC = "megapixels/sec, Sauvola w=15 k=0.2 A4 300dpi u8 pages (input pixels)"
WORKLOAD = "Sauvola w=15 k=0.2 morph=0, synthpage-v2 A4 2480x3508 u8 (BASELINE configs[1])"


def geometry(rows, cols, w):
    h = w // 2
    Hp, Wp = rows + 2 * h, cols + 2 * h
    return dict(h=h, Hp=Hp, Wp=Wp, out_rows=Hp - w, out_cols=Wp - w)


def anchor_rows(Hp, Wp):
    """rows of high-word anchors of the compact plane layout (csrc/common.cuh: prl_anchor_shift)"""
    for sh in (3, 2, 1, 0):
        if (((1 << sh) - 1) * Wp + 3.0 * Hp) * 65025.0 < 4294967296.0:
            return (Hp + (1 << sh) - 1) >> sh
    return Hp


def algorithmic_bytes(rows, cols, w, model="compact"):
    """Compulsory traffic of the integral-image pipeline per page (each kernel reads its inputs and writes its outputs once).
    int64   : SURVEY.md 8(d), S and Q as int64 planes:  K1 = H*W + 16*Hp*Wp,  K2 = 16*Hp*Wp + 2*Hout*Wout
    compact : what is benchmarked since round 2 (SURVEY 7 'exact 32-bit planes'): one interleaved plane of {S, Q} low
              words (8 B per padded pixel) + 2 B per pixel of high words on every A-th row:
              K1 = H*W + 8*Hp*Wp + 2*ceil(Hp/A)*Wp,  K2 = 8*Hp*Wp + 2*Hout*Wout"""
    g = geometry(rows, cols, w)
    px = g["Hp"] * g["Wp"]
    if model == "int64":
        return rows * cols + 16 * px, 16 * px + 2 * g["out_rows"] * g["out_cols"]
    return rows * cols + 8 * px + 2 * anchor_rows(g["Hp"], g["Wp"]) * g["Wp"], 8 * px + 2 * g["out_rows"] * g["out_cols"]


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_config(pages_per_gpu, world):
    """The SAME dict on both arms (the driver compares them)."""
    k1, _ = algorithmic_bytes(ROWS, COLS, WINDOW)
    return {"workload": WORKLOAD, "pages_per_gpu": pages_per_gpu, "rows": ROWS, "cols": COLS, "window": WINDOW, "k": K_COEF,
            "morph": 0, "parallelism": f"page-sharded x{world}, no collective",
            "l2": f"no flush needed: {pages_per_gpu * ROWS * COLS / 1e9:.2f} GB of pages + "
                  f"{pages_per_gpu * (k1 - ROWS * COLS) / 1e9:.1f} GB of S/Q planes per step >> 126 MB L2"}


def golden_digests():
    """sha1 of the Sauvola w=15 k=0.2 masks of synthpage-v2 pages 0 and 1 as produced by the REFERENCE's own C++
    (tests/golden/ref_golden.json, written by tests/golden/make_ref_golden.py from oracle/_ref)."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "ref_golden.json")) as f:
            im = json.load(f)["images"]
        return {0: im["a4_p0"]["masks"]["sauvola_w15_k0.2"]["sha1"], 1: im["a4_p1"]["masks"]["sauvola_w15_k0.2"]["sha1"]}
    except Exception:
        return {}


def sha1(a):
    import numpy as np
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median over the samples taken under load (the sampler also sees the idle gaps between the timed regions)
        busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": len(busy), "power_w_max": max(power)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own code on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
def cpu_kind():
    from oracle import ref as R
    return "reference" if R.available() else "port"


def _cpu_worker(args):
    idx_list, rows, cols, window, k, start_at, kind = args
    import cv2
    cv2.setNumThreads(1)
    from oracle import c_oracle as CO
    if kind == "reference":
        from oracle import ref as R
        fn = lambda pg: R.binarizeSauvola(pg, window, k, 0)               # prl::binarizeSauvola, the reference's own object code
    else:
        from oracle import prl_oracle as O
        fn = lambda pg: O.binarize_local(pg, O.SAUVOLA, window, (k,), 0)
    pages = [CO.synth_page(i % 256, rows, cols, SEED) for i in idx_list]   # untimed: inputs resident in host memory
    fn(pages[0][:64, :64])                                                # imports / first-call set-up outside the timed region
    while time.time() < start_at:                                         # common start line
        time.sleep(0.001)
    t0 = time.time()
    white = 0
    for pg in pages:
        out = fn(pg)
        white += int(out[0, 0])
    return t0, time.time(), len(pages)


def cpu_reference_pass(pool, cores, n_pages, rows, cols, kind, first_page=0):
    """One bounded pass: n_pages pages dealt over `cores` single-threaded workers; returns (MP/s, seconds, pages)."""
    per = [n_pages // cores + (1 if c < n_pages % cores else 0) for c in range(cores)]
    start_at = time.time() + 0.5 + 0.07 * max(per)                        # workers generate their pages first
    jobs, nxt = [], first_page
    for c in range(cores):
        jobs.append((list(range(nxt, nxt + per[c])), rows, cols, WINDOW, K_COEF, start_at, kind))
        nxt += per[c]
    res = [r for r in pool.map(_cpu_worker, [j for j in jobs if j[0]])]
    t0 = min(r[0] for r in res); t1 = max(r[1] for r in res)
    n = sum(r[2] for r in res)
    return n * rows * cols / 1e6 / (t1 - t0), t1 - t0, n


def make_pool(cores):
    import multiprocessing as mp
    return mp.get_context("spawn").Pool(cores)


CPU_NOTE = {"reference": "oracle/_ref = the reference's own binarizeSauvola.cpp compiled unmodified against a cv:: facade; every OpenCV "
                         "primitive runs in the cv2 4.13 wheel; one single-threaded worker process per host core",
            "port": "oracle/prl_oracle.py = the reference's OpenCV call sequence restated op for op (oracle/_ref not built); one "
                    "single-threaded worker process per host core"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import c_oracle as CO
    CO.build()
    kind = cpu_kind()
    cores = len(os.sched_getaffinity(0))
    pool = make_pool(cores)
    # calibrate, then size a step so that the whole run stays within ~2.5 minutes: up to the GPU arm's 256 pages per step
    _, dt, n = cpu_reference_pass(pool, cores, cores, ROWS, COLS, kind)
    pps = n / dt
    budget_s = 150.0
    per_step = int(budget_s * pps / max(args.steps + max(args.warmup, 1), 1))
    per_step = max(cores, min(args.pages, per_step // cores * cores))
    for _ in range(max(args.warmup, 1) - 1):
        cpu_reference_pass(pool, cores, per_step, ROWS, COLS, kind)
    times, n_tot = [], 0
    for s in range(args.steps):
        _, dt, n = cpu_reference_pass(pool, cores, per_step, ROWS, COLS, kind, first_page=s * per_step)
        times.append(dt); n_tot += n
    pool.close()
    total = sum(times)
    mps = n_tot * ROWS * COLS / 1e6 / total
    sample = (f"{per_step} synthpage-v2 A4 pages per step ({per_step / cores:.1f} per core; the GPU arm's step is {args.pages} pages), "
              f"{args.steps} steps; {CPU_NOTE[kind]}")
    line = {
        "impl": "reference", "metric": METRIC,
        "value": mps, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "pages_per_sec": n_tot / total,
        "config": make_config(args.pages, args.gpus),
        "cpu_baseline": {"value": mps, "unit": "MP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mps, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import prlib_b200
    from prlib_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    host_group = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        host_group = dist.new_group(backend="gloo")        # host-side barrier for the phases in which other ranks' GPUs must stay idle
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    numa_note = None
    if world > 1:
        # one process per GPU: run (and first-touch the pinned staging buffers) on the cores next to this GPU, otherwise
        # every rank's host buffers land on one socket and the end-to-end path crosses the inter-socket link
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa_note = f"rank pinned to the {len(cpus)} cores nearest its GPU (nvmlDeviceGetCpuAffinity)"
        except Exception as ex:
            numa_note = f"no CPU affinity set ({type(ex).__name__})"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")

    n_pages = args.pages
    g = geometry(ROWS, COLS, WINDOW)
    step_in = (COLS + 15) // 16 * 16
    step_out = (g["out_cols"] + 15) // 16 * 16
    ctx = prlib_b200.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)        # a real (non-legacy) stream: kernels, events and timing all live on it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    # synthetic pages, generated on the device (page index = global page id: rank shards are contiguous)
    from prlib_b200.sharding import shard_range
    lo, _ = shard_range(rank, world, n_pages * world)
    pages = torch.empty((n_pages, ROWS, step_in), dtype=torch.uint8, device=dev)
    masks = torch.empty((n_pages, g["out_rows"], step_out), dtype=torch.uint8, device=dev)
    ctx.synth_pages_dev(pages.data_ptr(), n_pages, ROWS, COLS, step_in, ROWS * step_in, SEED, lo)
    torch.cuda.synchronize()

    def run_batch(dst):
        ctx.binarize_local_batch_dev(capi.SAUVOLA, pages.data_ptr(), n_pages, ROWS, COLS, step_in, ROWS * step_in,
                                     WINDOW, (K_COEF,), 0, dst.data_ptr(), step_out, g["out_rows"] * step_out)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(group=host_group)

    def timed_steps(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    # ---- headline: the integral-image pipeline (kernel 1 -> kernel 2), fused small-window path switched off
    ctx.set_option("enable_fused", 0)
    for _ in range(args.warmup):
        run_batch(masks)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ctx.timing_reset(); ctx.timing_enable(True)
    l0 = ctx.launch_count()
    ms_total = timed_steps(lambda: run_batch(masks), args.steps)
    launches = ctx.launch_count() - l0
    ktimes = ctx.timing()
    ctx.timing_enable(False)

    # the timed output against the reference's own masks (pages 0 and 1 live on rank 0)
    golden = {}
    if rank == 0:
        want = golden_digests()
        for p, d in want.items():
            if p < n_pages:
                golden[f"page{p}"] = sha1(masks[p, :, :g["out_cols"]].cpu().numpy()) == d
    golden_ok = bool(golden) and all(golden.values())

    # ---- the library's default path for this window: fused small-window kernel (the planes never reach HBM);
    # scored against the problem-minimum bytes H*W + Hout*Wout
    default_path = None
    try:
        masks_f = torch.empty_like(masks)
        ctx.set_option("enable_fused", 1)
        for _ in range(max(2, args.warmup)):
            run_batch(masks_f)
        ctx.timing_reset(); ctx.timing_enable(True)
        ms_f = timed_steps(lambda: run_batch(masks_f), args.steps)
        ktimes_f = ctx.timing(); ctx.timing_enable(False)
        default_path = {"ms_per_step": ms_f / args.steps,
                        "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in ktimes_f.items()},
                        "masks_equal_two_kernel_path": bool(torch.equal(masks_f[:, :, :g["out_cols"]], masks[:, :, :g["out_cols"]])),
                        "pages_handed_back_to_two_kernel_path": int(ctx.fused_redo_count())}
        del masks_f
    except Exception as ex:
        default_path = {"error": str(ex)}

    # ---- end to end through the host C-ABI (pinned host buffers; H2D + D2H inside the timed region), library defaults
    host_pages = torch.empty((n_pages, ROWS, COLS), dtype=torch.uint8, pin_memory=True)
    host_pages.copy_(pages[:, :, :COLS])
    host_masks = torch.empty((n_pages, g["out_rows"], g["out_cols"]), dtype=torch.uint8, pin_memory=True)
    hp, hm = host_pages.numpy(), host_masks.numpy()
    ctx.set_stream(None)
    e2e_warm = max(1, min(args.warmup, 2))
    for _ in range(e2e_warm):
        prlib_b200.binarize_batch(hp, capi.SAUVOLA, WINDOW, (K_COEF,), 0, devices=[local_rank], out=hm)
    barrier()
    e2e_steps = args.steps
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prlib_b200.binarize_batch(hp, capi.SAUVOLA, WINDOW, (K_COEF,), 0, devices=[local_rank], out=hm)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    unpack_threads = int(capi.load().prl_cuda_batch_unpack_threads())     # > 0: the masks crossed PCIe as bits (library default where the host has the cores)
    # the e2e masks must equal the device-resident ones
    same = bool(torch.equal(host_masks[:2].to(dev), masks[:2, :, :g["out_cols"]])) and \
        bool(torch.equal(host_masks[-1:].to(dev), masks[-1:, :, :g["out_cols"]]))
    # the same call with the masks crossing PCIe as bytes (round 1's return path; "batch_unpack_threads" = 0)
    prlib_b200.set_global_option("batch_unpack_threads", 0)
    host_masks.zero_()
    prlib_b200.binarize_batch(hp, capi.SAUVOLA, WINDOW, (K_COEF,), 0, devices=[local_rank], out=hm)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prlib_b200.binarize_batch(hp, capi.SAUVOLA, WINDOW, (K_COEF,), 0, devices=[local_rank], out=hm)
    torch.cuda.synchronize()
    bytes_s = time.perf_counter() - t0
    prlib_b200.set_global_option("batch_unpack_threads", -1)
    same = same and bool(torch.equal(host_masks[:2].to(dev), masks[:2, :, :g["out_cols"]]))
    # extra: the same call with 1-bit-per-pixel output (prl_cuda_binarize_batch_packed, PIX layout): D2H is 8x smaller
    wpl = (g["out_cols"] + 31) // 32
    host_bits = torch.empty((n_pages, g["out_rows"], wpl), dtype=torch.int32, pin_memory=True)
    hb = host_bits.numpy().view(np.uint32)
    prlib_b200.binarize_batch(hp, capi.SAUVOLA, WINDOW, (K_COEF,), 0, devices=[local_rank], out=hb, packed=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prlib_b200.binarize_batch(hp, capi.SAUVOLA, WINDOW, (K_COEF,), 0, devices=[local_rank], out=hb, packed=True)
    torch.cuda.synchronize()
    packed_s = time.perf_counter() - t0
    packed_same = bool(np.array_equal(prlib_b200.unpack_lept1(hb[:1], g["out_cols"]), hm[:1]))
    mask0 = hm[0].copy()                                          # (the copy-rate probe below overwrites the host buffers)

    # ---- the host<->device fabric itself, all ranks at once: what the copy engines give with nothing else running
    d_lin_in = torch.empty((n_pages, ROWS, COLS), dtype=torch.uint8, device=dev)
    d_lin_out = torch.empty((n_pages, g["out_rows"], g["out_cols"]), dtype=torch.uint8, device=dev)
    s_h2d, s_d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def copy_rate(h2d, d2h, reps=3):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s_h2d):
                    d_lin_in.copy_(host_pages, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_d2h):
                    host_masks.copy_(d_lin_out, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    copy_rate(True, True, 1)
    pcie_t = [copy_rate(True, False), copy_rate(False, True), copy_rate(True, True)]
    del d_lin_in, d_lin_out

    # ---- the reference-signature call: ONE image per call, host pointers in and out (prl::binarizeSauvola through the shim
    # ends in exactly this C-ABI call); pageable numpy buffers like a cv::Mat's
    single = None
    if rank == 0:
        try:
            gray = np.ascontiguousarray(hp[0])
            bgr = np.repeat(gray[:, :, None], 3, axis=2)
            out0 = None
            lat = {}
            for name, fn in (("gray", lambda: ctx.binarize_local(gray, capi.SAUVOLA, WINDOW, (K_COEF,), 0)),
                             ("bgr", lambda: ctx.binarize_image(bgr, capi.SAUVOLA, WINDOW, (K_COEF,), 0))):
                for _ in range(3):
                    out0 = fn()
                ts = []
                for _ in range(15):
                    t0 = time.perf_counter(); out0 = fn(); ts.append(time.perf_counter() - t0)
                lat[name] = statistics.median(ts)
            single = {"api": "prl_cuda_binarize_local / prl_cuda_binarize_local_image (one A4 image per call, pageable host memory, synchronous)",
                      "gray_ms": 1e3 * lat["gray"], "bgr_ms": 1e3 * lat["bgr"],
                      "gray_pages_per_sec": 1.0 / lat["gray"], "bgr_pages_per_sec": 1.0 / lat["bgr"],
                      "mask_equals_batch_path": bool(np.array_equal(out0, mask0))}
        except Exception as ex:
            single = {"error": f"{type(ex).__name__}: {ex}"}

    clocks = sampler.stop() if sampler else None

    # ---- N > 1: the product's own page dispatcher.  Rank 0 ALONE hands BASELINE configs[4] (8192 A4 pages = 32 x `p mod 256`)
    # to prl_cuda_binarize_batch(devices = 0..N-1) in ONE process while the other ranks idle at a host-side barrier.
    dispatcher = None
    if world > 1:
        del pages
        torch.cuda.empty_cache()
        host_barrier()
        if rank == 0:
            try:
                big_n, calls = 1024, 8
                big_in = torch.empty((big_n, ROWS, COLS), dtype=torch.uint8, pin_memory=True)
                big_out = torch.empty((big_n, g["out_rows"], g["out_cols"]), dtype=torch.uint8, pin_memory=True)
                for r in range(big_n // n_pages):
                    big_in[r * n_pages:(r + 1) * n_pages].copy_(host_pages)
                bi, bo = big_in.numpy(), big_out.numpy()
                devs = list(range(world))
                prlib_b200.binarize_batch(bi, capi.SAUVOLA, WINDOW, (K_COEF,), 0, devices=devs, out=bo)      # warm-up: workers, rings
                t0 = time.perf_counter()
                for _ in range(calls):
                    prlib_b200.binarize_batch(bi, capi.SAUVOLA, WINDOW, (K_COEF,), 0, devices=devs, out=bo)
                dt = time.perf_counter() - t0
                want = {p: sha1(masks[p, :, :g["out_cols"]].cpu().numpy()) for p in (0, 1, n_pages // 2, n_pages - 1)}
                ident = all(sha1(bo[r * n_pages + p]) == want[p] for r in range(big_n // n_pages) for p in want)
                dispatcher = {"api": "prl_cuda_binarize_batch(devices=[0..N-1]) called by ONE process (rank 0); the other ranks idle",
                              "n_dev": world, "pages": big_n * calls,
                              "batch": f"{calls} calls x {big_n} pinned pages = {big_n * calls} pages (page p = synthpage-v2 page p mod {n_pages})",
                              "pages_per_sec": big_n * calls / dt, "value": big_n * calls * ROWS * COLS / 1e6 / dt, "unit": "MP/s",
                              "seconds": dt, "byte_identical_to_n1": ident}
                del big_in, big_out
            except Exception as ex:
                dispatcher = {"error": f"{type(ex).__name__}: {ex}"}
        host_barrier()

    t_dev = torch.tensor([ms_total, e2e_s * 1e3, packed_s * 1e3, (default_path or {}).get("ms_per_step", 0.0)] + [t * 1e3 for t in pcie_t] +
                         [bytes_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, packed_ms = float(t_dev[0]), float(t_dev[1]), float(t_dev[2])
    if default_path and "ms_per_step" in default_path:
        default_path["ms_per_step"] = float(t_dev[3])
    pcie_ms = [float(t_dev[4]), float(t_dev[5]), float(t_dev[6])]
    bytes_ms = float(t_dev[7])

    if rank == 0:
        total_pages = n_pages * world
        mp_per_page = ROWS * COLS / 1e6
        value = total_pages * args.steps * mp_per_page / (ms_total / 1e3)
        e2e_value = total_pages * e2e_steps * mp_per_page / (e2e_ms / 1e3)
        peak, peak_src = measured_peak()
        k1b, k2b = algorithmic_bytes(ROWS, COLS, WINDOW, "compact")
        k1b64, k2b64 = algorithmic_bytes(ROWS, COLS, WINDOW, "int64")
        kern = {}
        for fam, nbytes, nb64 in (("integral", k1b, k1b64), ("threshold", k2b, k2b64)):
            if fam in ktimes and ktimes[fam]["launches"]:
                avg_ms = ktimes[fam]["ms"] / ktimes[fam]["launches"]
                pages_per_launch = n_pages * args.steps / ktimes[fam]["launches"]
                gbs = nbytes * pages_per_launch / (avg_ms / 1e3) / 1e9
                kern[fam] = {"avg_ms": avg_ms, "launches": ktimes[fam]["launches"], "algorithmic_bytes_per_launch": nbytes * pages_per_launch,
                             "achieved_gbs": gbs, "frac": gbs / peak,
                             "frac_int64_model": nb64 * pages_per_launch / (avg_ms / 1e3) / 1e9 / peak}
        dom = max(kern, key=lambda f: kern[f]["avg_ms"] * kern[f]["launches"]) if kern else None
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic = tj.get(dom, {}).get("dram_bytes_per_launch")
            if traffic is not None and tj[dom].get("pages_per_launch"):
                traffic = traffic * (n_pages / tj[dom]["pages_per_launch"])
            traffic_src = tj.get(dom, {}).get("source")
        except Exception:
            pass
        roofline = None
        if dom:
            pipe_gbs = (k1b + k2b) * total_pages / world * args.steps / (ms_total / 1e3) / 1e9
            roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                        "frac": kern[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                        "bytes_model": "compact planes (what runs): K1 = H*W + 8*Hp*Wp + 2*ceil(Hp/8)*Wp, K2 = 8*Hp*Wp + 2*Hout*Wout per page; "
                                       "frac_int64_model rates the same times against SURVEY 8(d)'s int64 planes (K1 = H*W + 16*Hp*Wp, "
                                       "K2 = 16*Hp*Wp + 2*Hout*Wout), i.e. against the best an int64-plane pipeline could do",
                        "kernels": kern,
                        "pipeline": {"algorithmic_bytes_per_page": k1b + k2b, "achieved_gbs": pipe_gbs, "frac": pipe_gbs / peak,
                                     "algorithmic_bytes_per_page_int64_model": k1b64 + k2b64,
                                     "frac_int64_model": pipe_gbs * (k1b64 + k2b64) / (k1b + k2b) / peak,
                                     "frac_of_nominal_8TBs": pipe_gbs / 8000.0}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            kind = "port"
            try:
                from oracle import c_oracle as CO
                CO.build()
                kind = cpu_kind()
                cores = len(os.sched_getaffinity(0))
                pool = make_pool(cores)
                cpu_reference_pass(pool, cores, cores, ROWS, COLS, kind)                      # warm-up (imports, page-in)
                mps, dt, n = cpu_reference_pass(pool, cores, 2 * cores, ROWS, COLS, kind, first_page=cores)
                pool.close()
                cpu = {"value": mps, "unit": "MP/s", "cores": cores, "kind": kind,
                       "sample": f"{n} synthpage-v2 A4 pages, 2 per core, {dt:.2f} s wall; {CPU_NOTE[kind]}",
                       "pages_per_sec": n / dt}
            except Exception as ex:  # the GPU numbers stand on their own
                cpu = {"value": None, "unit": "MP/s", "cores": len(os.sched_getaffinity(0)), "kind": kind, "sample": f"failed: {ex}"}
        in_bytes, out_bytes = n_pages * ROWS * COLS, n_pages * g["out_rows"] * g["out_cols"]
        pcie = {"h2d_alone_gbs": world * in_bytes / (pcie_ms[0] / 1e3) / 1e9, "d2h_alone_gbs": world * out_bytes / (pcie_ms[1] / 1e3) / 1e9,
                "both_each_way_gbs": world * min(in_bytes, out_bytes) / (pcie_ms[2] / 1e3) / 1e9,
                "note": f"aggregate over {world} GPU(s), all ranks copying at once: one linear pinned copy of the step's pages in / masks out per "
                        "direction, then both directions together; no kernels running"}
        pcie["e2e_ceiling_pages_per_sec"] = total_pages / (pcie_ms[2] / 1e3)          # masks returned as bytes: both directions loaded alike
        pcie["e2e_ceiling_bits_return_pages_per_sec"] = total_pages / (pcie_ms[0] / 1e3)   # masks returned as bits: H2D is the loaded direction
        bits_bytes = n_pages * g["out_rows"] * wpl * 4
        line = {
            "metric": METRIC,
            "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "pages_per_sec": total_pages * args.steps / (ms_total / 1e3),
            "config": make_config(n_pages, world),
            "golden_check": {"ok": golden_ok, "pages": golden,
                             "what": "sha1 of masks 0 and 1 of the timed batch == tests/golden/ref_golden.json (outputs of the reference's own C++)"},
            "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": bits_bytes if unpack_threads > 0 else out_bytes, "ms_per_step": e2e_ms / e2e_steps,
                    "pages_per_sec": total_pages * e2e_steps / (e2e_ms / 1e3), "masks_match_device_path": same,
                    "host_result_bytes_per_step": out_bytes,
                    "frac_of_pcie_ceiling": (total_pages * e2e_steps / (e2e_ms / 1e3)) /
                    pcie["e2e_ceiling_bits_return_pages_per_sec" if unpack_threads > 0 else "e2e_ceiling_pages_per_sec"],
                    "mask_return": (f"1 bit per pixel over PCIe, expanded to the caller's 0/255 bytes by {unpack_threads} library host threads per GPU "
                                    "inside the call") if unpack_threads > 0 else "0/255 bytes over PCIe (too few host cores per GPU for the 1-bit return path)",
                    "api": "prl_cuda_binarize_batch (pinned host pages -> H2D -> kernels -> D2H -> host masks of 0/255 bytes; 3-slot ring, library defaults)",
                    "bytes_return": {"value": total_pages * e2e_steps * mp_per_page / (bytes_ms / 1e3), "unit": "MP/s",
                                     "pages_per_sec": total_pages * e2e_steps / (bytes_ms / 1e3), "d2h_bytes_per_step": out_bytes,
                                     "frac_of_pcie_ceiling": (total_pages * e2e_steps / (bytes_ms / 1e3)) / pcie["e2e_ceiling_pages_per_sec"],
                                     "api": "the same call with \"batch_unpack_threads\" = 0: the byte masks themselves cross PCIe (round 1's path)"},
                    "host_affinity": numa_note},
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "pcie": pcie,
        }
        line["e2e"]["packed_1bpp"] = {"value": total_pages * e2e_steps * mp_per_page / (packed_ms / 1e3), "unit": "MP/s",
                                      "pages_per_sec": total_pages * e2e_steps / (packed_ms / 1e3),
                                      "d2h_bytes_per_step": n_pages * g["out_rows"] * wpl * 4, "equals_byte_masks": packed_same,
                                      "api": "prl_cuda_binarize_batch_packed (extra; the headline e2e returns 0/255 bytes like the reference)"}
        if default_path and "ms_per_step" in default_path:
            pmin = ROWS * COLS + g["out_rows"] * g["out_cols"]
            pps_f = total_pages / (default_path["ms_per_step"] / 1e3)
            default_path.update({"value": pps_f * mp_per_page, "unit": "MP/s", "pages_per_sec": pps_f,
                                 "problem_minimum_bytes_per_page": pmin, "frac_of_problem_minimum_roofline": pmin * pps_f / world / 1e9 / peak,
                                 "note": "what prl_cuda_binarize_local_batch_dev runs by default for windows <= 31: the fused small-window kernel "
                                         "(S/Q never reach HBM); `value` above is the integral-image pipeline north_star names"})
            line["value_default_path"] = default_path["value"]
        line["default_path"] = default_path
        line["e2e_single_call"] = single
        if dispatcher is not None:
            line["e2e_dispatcher"] = dispatcher
        if world == 1 and not args.no_other_configs:
            try:
                sys.path.insert(0, os.path.join(ROOT, "scripts"))
                import bench_configs
                del masks, host_pages, host_masks, host_bits
                torch.cuda.empty_cache()
                line["other_configs"] = bench_configs.run_all(ctx_device=local_rank, quick=True)
            except Exception as ex:
                line["other_configs"] = {"error": f"{type(ex).__name__}: {ex}"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pages", type=int, default=256, help="pages per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip BASELINE configs 3 and 4 (extra keys of the N=1 line)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
