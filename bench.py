#!/usr/bin/env python
"""bench.py -- throughput of the local-statistics binarization hot path on B200.

Metric (BASELINE.json): megapixels/s (and pages/s) of Sauvola w=15, k=0.2, morph 0 over synthetic
A4 300-dpi pages (2480 x 3508 u8, synthpage-v2 seed 2024), plus the fraction of the HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pages P] [--impl ours|reference]

One "step" = one pass of the hot path (kernel 1 integral -> kernel 2 threshold) over a batch of
P pages PER GPU (weak scaling; pages are independent, no data-path collective).
  value  : whole-job MP/s with the pages already resident in HBM (device-timed, max over ranks)
  e2e    : the same metric through the host C-ABI call prl_cuda_binarize_batch with pinned HOST
           buffers, H2D and D2H inside the timed region
  roofline / cpu_baseline : see DESIGN.md section "Measurement"
--impl reference times the reference's CPU implementation of the path (its OpenCV call sequence,
oracle/prl_oracle.py -- the C++ library itself cannot be built in this image) on the host cores.
"""
from __future__ import annotations

import argparse
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
METHOD_NAME = "sauvola"
SEED = 2024


def geometry(rows, cols, w):
    h = w // 2
    Hp, Wp = rows + 2 * h, cols + 2 * h
    return dict(h=h, Hp=Hp, Wp=Wp, out_rows=Hp - w, out_cols=Wp - w)


def algorithmic_bytes(rows, cols, w):
    """compulsory traffic of the integral-image pipeline, per page."""
    g = geometry(rows, cols, w)
    k1 = rows * cols + 16 * g["Hp"] * g["Wp"]
    k2 = 16 * g["Hp"] * g["Wp"] + 2 * g["out_rows"] * g["out_cols"]
    return k1, k2


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's OpenCV call sequence on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    idx_list, rows, cols, window, k, barrier_t0 = args
    import numpy as np  # noqa: F401
    import cv2
    cv2.setNumThreads(1)
    from oracle import c_oracle as CO
    from oracle import prl_oracle as O
    pages = [CO.synth_page(i, rows, cols, SEED) for i in idx_list]       # untimed: inputs resident in host memory
    while time.time() < barrier_t0:                                      # common start line
        time.sleep(0.001)
    t0 = time.time()
    white = 0
    for pg in pages:
        out = O.binarize_local(pg, O.SAUVOLA, window, (k,), 0)
        white += int(out[0, 0])
    return t0, time.time(), len(pages)


def cpu_reference_pass(pool, cores, pages_per_core, rows, cols, first_page=0):
    """One bounded pass: cores x pages_per_core pages, one worker per core; returns (MP/s, seconds)."""
    start_at = time.time() + 0.25 + 0.06 * pages_per_core * 1.0 + 0.2   # workers generate their pages first
    jobs = [([first_page + c * pages_per_core + j for j in range(pages_per_core)], rows, cols, WINDOW, K_COEF, start_at)
            for c in range(cores)]
    res = pool.map(_cpu_worker, jobs)
    t0 = min(r[0] for r in res); t1 = max(r[1] for r in res)
    n = sum(r[2] for r in res)
    return n * rows * cols / 1e6 / (t1 - t0), t1 - t0, n


def make_pool(cores):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    return ctx.Pool(cores)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import c_oracle as CO
    CO.build()
    cores = len(os.sched_getaffinity(0))
    pool = make_pool(cores)
    ppc = 1
    for _ in range(max(args.warmup, 1)):
        cpu_reference_pass(pool, cores, ppc, ROWS, COLS)
    times, n_tot = [], 0
    for s in range(args.steps):
        _, dt, n = cpu_reference_pass(pool, cores, ppc, ROWS, COLS, first_page=s * cores)
        times.append(dt); n_tot += n
    pool.close()
    total = sum(times)
    mps = n_tot * ROWS * COLS / 1e6 / total
    sample = f"{cores * ppc} synthpage-v2 A4 pages per step (one per core), {args.steps} steps"
    line = {
        "impl": "reference", "metric": "megapixels/sec, Sauvola w=15 k=0.2 A4 300dpi u8 pages (input pixels)",
        "value": mps, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "pages_per_sec": n_tot / total,
        "config": {"workload": "Sauvola w=15 k=0.2 morph=0, synthpage-v2 A4 2480x3508 u8 (BASELINE configs[1])",
                   "rows": ROWS, "cols": COLS, "window": WINDOW, "k": K_COEF,
                   "note": "reference arm = the reference's OpenCV call sequence (cv2 4.13, filter2D direct path) "
                           "restated op for op in oracle/prl_oracle.py; the C++ library cannot be built here "
                           "(no OpenCV C++ SDK / Leptonica)"},
        "cpu_baseline": {"value": mps, "unit": "MP/s", "cores": cores, "kind": "port", "sample": sample},
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
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
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

    def step():
        ctx.binarize_local_batch_dev(capi.SAUVOLA, pages.data_ptr(), n_pages, ROWS, COLS, step_in, ROWS * step_in,
                                     WINDOW, (K_COEF,), 0, masks.data_ptr(), step_out, g["out_rows"] * step_out)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ctx.timing_reset(); ctx.timing_enable(True)
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    ktimes = ctx.timing()
    ctx.timing_enable(False)

    # ---- extra (not the headline): the opt-in fused small-window path (integral planes never reach HBM),
    # same pages, same masks; scored against the problem-minimum bytes H*W + Hout*Wout
    fused = None
    try:
        masks_f = torch.empty_like(masks)
        ctx.set_option("enable_fused", 1)
        def step_f():
            ctx.binarize_local_batch_dev(capi.SAUVOLA, pages.data_ptr(), n_pages, ROWS, COLS, step_in, ROWS * step_in,
                                         WINDOW, (K_COEF,), 0, masks_f.data_ptr(), step_out, g["out_rows"] * step_out)
        for _ in range(2):
            step_f()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            step_f()
        f1.record(stream)
        barrier()
        fused = {"ms_per_step": f0.elapsed_time(f1) / args.steps,
                 "masks_equal_two_kernel_path": bool(torch.equal(masks_f[:, :, :g["out_cols"]], masks[:, :, :g["out_cols"]]))}
        del masks_f
    except Exception as ex:
        fused = {"error": str(ex)}
    finally:
        ctx.set_option("enable_fused", 0)

    # ---- end to end through the host C-ABI (pinned host buffers; H2D + D2H inside the timed region)
    host_pages = torch.empty((n_pages, ROWS, COLS), dtype=torch.uint8).pin_memory()
    host_pages.copy_(pages[:, :, :COLS])
    host_masks = torch.empty((n_pages, g["out_rows"], g["out_cols"]), dtype=torch.uint8).pin_memory()
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
    # the e2e masks must equal the device-resident ones (same kernels)
    same = bool(torch.equal(host_masks[:2].to(dev), masks[:2, :, :g["out_cols"]]))
    # extra: the same call with 1-bit-per-pixel output (prl_cuda_binarize_batch_packed, PIX layout): D2H is 8x smaller
    wpl = (g["out_cols"] + 31) // 32
    host_bits = torch.empty((n_pages, g["out_rows"], wpl), dtype=torch.int32).pin_memory()
    hb = host_bits.numpy().view(np.uint32)
    prlib_b200.binarize_batch(hp, capi.SAUVOLA, WINDOW, (K_COEF,), 0, devices=[local_rank], out=hb, packed=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prlib_b200.binarize_batch(hp, capi.SAUVOLA, WINDOW, (K_COEF,), 0, devices=[local_rank], out=hb, packed=True)
    torch.cuda.synchronize()
    packed_s = time.perf_counter() - t0
    packed_same = bool(np.array_equal(prlib_b200.unpack_lept1(hb[:1], g["out_cols"]), hm[:1]))

    clocks = sampler.stop() if sampler else None

    t_dev = torch.tensor([ms_total, e2e_s * 1e3, packed_s * 1e3, (fused or {}).get("ms_per_step", 0.0)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, packed_ms = float(t_dev[0]), float(t_dev[1]), float(t_dev[2])
    if fused and "ms_per_step" in fused:
        fused["ms_per_step"] = float(t_dev[3])

    if rank == 0:
        total_pages = n_pages * world
        mp_per_page = ROWS * COLS / 1e6
        value = total_pages * args.steps * mp_per_page / (ms_total / 1e3)
        e2e_value = total_pages * e2e_steps * mp_per_page / (e2e_ms / 1e3)
        peak, peak_src = measured_peak()
        k1b, k2b = algorithmic_bytes(ROWS, COLS, WINDOW)
        kern = {}
        for fam, nbytes in (("integral", k1b), ("threshold", k2b)):
            if fam in ktimes and ktimes[fam]["launches"]:
                avg_ms = ktimes[fam]["ms"] / ktimes[fam]["launches"]
                pages_per_launch = n_pages * args.steps / ktimes[fam]["launches"]
                gbs = nbytes * pages_per_launch / (avg_ms / 1e3) / 1e9
                kern[fam] = {"avg_ms": avg_ms, "launches": ktimes[fam]["launches"], "algorithmic_bytes_per_launch": nbytes * pages_per_launch,
                             "achieved_gbs": gbs, "frac": gbs / peak}
        dom = max(kern, key=lambda f: kern[f]["avg_ms"] * kern[f]["launches"]) if kern else None
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = None
        if dom:
            roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                        "frac": kern[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                        "bytes_model": "SURVEY 8(d) integral-image pipeline: K1 = H*W + 16*Hp*Wp, K2 = 16*Hp*Wp + 2*Hout*Wout per page",
                        "kernels": kern,
                        "pipeline": {"algorithmic_bytes_per_page": k1b + k2b,
                                     "achieved_gbs": (k1b + k2b) * total_pages / world * args.steps / (ms_total / 1e3) / 1e9,
                                     "frac": (k1b + k2b) * total_pages / world * args.steps / (ms_total / 1e3) / 1e9 / peak}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import c_oracle as CO
                CO.build()
                cores = len(os.sched_getaffinity(0))
                pool = make_pool(cores)
                cpu_reference_pass(pool, cores, 1, ROWS, COLS)                       # warm-up (imports, page-in)
                ppc = 2
                mps, dt, n = cpu_reference_pass(pool, cores, ppc, ROWS, COLS, first_page=cores)
                pool.close()
                cpu = {"value": mps, "unit": "MP/s", "cores": cores, "kind": "port",
                       "sample": f"{n} synthpage-v2 A4 pages, {ppc} per core, one cv2-single-thread worker per core, {dt:.2f} s wall",
                       "pages_per_sec": n / dt}
            except Exception as ex:  # the GPU numbers stand on their own
                cpu = {"value": None, "unit": "MP/s", "cores": len(os.sched_getaffinity(0)), "kind": "port", "sample": f"failed: {ex}"}
        line = {
            "metric": "megapixels/sec, Sauvola w=15 k=0.2 A4 300dpi u8 pages (input pixels)",
            "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "pages_per_sec": total_pages * args.steps / (ms_total / 1e3),
            "config": {"workload": "Sauvola w=15 k=0.2 morph=0, synthpage-v2 A4 2480x3508 u8 (BASELINE configs[1])",
                       "pages_per_gpu": n_pages, "rows": ROWS, "cols": COLS, "window": WINDOW, "k": K_COEF,
                       "parallelism": f"page-sharded x{world}, no collective",
                       "l2": f"no flush needed: {n_pages * ROWS * COLS / 1e9:.2f} GB of pages + {n_pages * (k1b - ROWS * COLS) / 1e9:.1f} GB of S/Q planes per step >> 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": n_pages * ROWS * COLS,
                    "d2h_bytes_per_step": n_pages * g["out_rows"] * g["out_cols"], "ms_per_step": e2e_ms / e2e_steps,
                    "pages_per_sec": total_pages * e2e_steps / (e2e_ms / 1e3), "masks_match_device_path": same,
                    "api": "prl_cuda_binarize_batch (pinned host pages -> H2D -> K1 -> K2 -> D2H -> host masks, 3-slot ring)",
                    "host_affinity": numa_note},
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        }
        line["e2e"]["packed_1bpp"] = {"value": total_pages * e2e_steps * mp_per_page / (packed_ms / 1e3), "unit": "MP/s",
                                      "pages_per_sec": total_pages * e2e_steps / (packed_ms / 1e3),
                                      "d2h_bytes_per_step": n_pages * g["out_rows"] * wpl * 4, "equals_byte_masks": packed_same,
                                      "api": "prl_cuda_binarize_batch_packed (extra; the headline e2e returns 0/255 bytes like the reference)"}
        if fused and "ms_per_step" in fused:
            pmin = ROWS * COLS + g["out_rows"] * g["out_cols"]
            pps_f = total_pages / (fused["ms_per_step"] / 1e3)
            fused.update({"value": pps_f * mp_per_page, "unit": "MP/s", "pages_per_sec": pps_f,
                          "problem_minimum_bytes_per_page": pmin, "frac_of_problem_minimum_roofline": pmin * pps_f / world / 1e9 / peak,
                          "note": "opt-in (set_option enable_fused): not the headline, which is the integral-image pipeline north_star names"})
        line["fused_small_window_path"] = fused
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
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
