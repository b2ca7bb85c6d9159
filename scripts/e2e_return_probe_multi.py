"""end-to-end rate of prl_cuda_binarize_batch by return path with every GPU of the box busy (diagnostic).
torchrun --nproc-per-node N scripts/e2e_return_probe_multi.py   (one process per GPU, gloo barrier, wall clock, max over ranks)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist, prlib_b200
from prlib_b200 import capi
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("gloo")
torch.cuda.set_device(local)
n, rows, cols = 256, 3508, 2480
ctx = prlib_b200.Context(local)
d = torch.empty((n, rows, cols), dtype=torch.uint8, device="cuda")
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.synth_pages_dev(d.data_ptr(), n, rows, cols, cols, rows * cols, 2024, 0)
hp = torch.empty((n, rows, cols), dtype=torch.uint8, pin_memory=True); hp.copy_(d); torch.cuda.synchronize(); del d
hm = torch.empty((n, rows - 1, cols - 1), dtype=torch.uint8, pin_memory=True)
if rank == 0:
    print(json.dumps({"host_cores": len(os.sched_getaffinity(0)), "gpus": world, "auto_threads": int(capi.load().prl_cuda_batch_unpack_threads())}), flush=True)
ref = None
cases = [(0, 1), (1, 1), (2, 1), (3, 1), (4, 1), (6, 1)]
if os.environ.get("PROBE_CASES"):
    cases = [tuple(int(v) for v in c.split(":")) for c in os.environ["PROBE_CASES"].split(",")]
for threads, nt in cases:
    prlib_b200.set_global_option("batch_unpack_threads", threads); prlib_b200.set_global_option("batch_unpack_nt", nt)
    f = lambda: prlib_b200.binarize_batch(hp.numpy(), capi.SAUVOLA, 15, (0.2,), 0, devices=[local], out=hm.numpy())
    f()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(4): f()
    dt = torch.tensor([(time.perf_counter() - t0) / 4], dtype=torch.float64)
    if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if ref is None: ref = hm.numpy()[::41].copy()
    same = bool(np.array_equal(hm.numpy()[::41], ref))
    if rank == 0:
        print(json.dumps({"unpack_threads_per_gpu": threads, "non_temporal": nt, "pages_per_sec_all_gpus": round(n * world / float(dt[0]), 1), "same": same}), flush=True)
