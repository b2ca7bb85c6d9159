"""end-to-end rate of prl_cuda_binarize_batch by return path (diagnostic): python scripts/e2e_return_probe.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, prlib_b200
from prlib_b200 import capi
n, rows, cols = 256, 3508, 2480
ctx = prlib_b200.Context(0)
d = torch.empty((n, rows, cols), dtype=torch.uint8, device="cuda")
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.synth_pages_dev(d.data_ptr(), n, rows, cols, cols, rows * cols, 2024, 0)
hp = torch.empty((n, rows, cols), dtype=torch.uint8, pin_memory=True); hp.copy_(d); torch.cuda.synchronize(); del d
hm = torch.empty((n, rows - 1, cols - 1), dtype=torch.uint8, pin_memory=True)
pageable = np.empty((n, rows - 1, cols - 1), np.uint8)
print(json.dumps({"host_cores": len(os.sched_getaffinity(0))}))
ref = None
for threads, nt, lag, out, tag in ((0, 1, 0, hm.numpy(), "bytes over PCIe, pinned"), (8, 1, 1, hm.numpy(), "bits"), (8, 1, 0, hm.numpy(), "bits"), (8, 1, 1, hm.numpy(), "bits"),
                                   (8, 1, 0, hm.numpy(), "bits"), (6, 1, 0, hm.numpy(), "bits"), (10, 1, 0, hm.numpy(), "bits"), (8, 0, 0, hm.numpy(), "bits"),
                                   (8, 1, 0, pageable, "bits, pageable masks")):
    prlib_b200.set_global_option("batch_unpack_lag", lag)
    prlib_b200.set_global_option("batch_unpack_threads", threads); prlib_b200.set_global_option("batch_unpack_nt", nt)
    f = lambda: prlib_b200.binarize_batch(hp.numpy(), capi.SAUVOLA, 15, (0.2,), 0, devices=[0], out=out)
    f(); f()
    t0 = time.perf_counter()
    for _ in range(4): f()
    dt = (time.perf_counter() - t0) / 4
    if ref is None: ref = out[::37].copy()
    print(json.dumps({"return": tag, "unpack_threads": threads, "non_temporal": nt, "submitter_waits": lag, "pages_per_sec": round(n / dt, 1), "same": bool(np.array_equal(out[::37], ref))}), flush=True)
