"""adaptive family over a batch of A4 pages resident in HBM (diagnostic): python scripts/batch_f4_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prlib_b200
ctx = prlib_b200.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
rows, cols = 3508, 2480
step = (cols + 15) // 16 * 16
native = dict(gray_first=1, blur=1, blur_ksize=5, assert_ksize=1, method=1, type=1, maxval=255.0, check_maxval=1, block_size=19, auto_block=1,
              delta=9.0, invert_if_dark=1)
cases = {"binarizeNativeAdaptive defaults (median 5, Gaussian mean 19)": native,
         "binarizeNativeAdaptive + bilateral 5": dict(native, bilateral_d=5, bilateral_sigma_color=150.0, bilateral_sigma_space=150.0),
         "NativeAdaptive, MEAN_C, auto block": dict(native, method=0, block_size=0),
         "adaptiveThreshold MEAN_C 19 alone": dict(gray_first=1, blur=0, method=0, type=0, maxval=255.0, block_size=19, delta=9.0)}
for n in (16, 64, 256):
    buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
    ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
    out = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
    for name, kw in cases.items():
        f = lambda: ctx.binarize_adaptive_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 1, out.data_ptr(), step, rows * step, **kw)
        for _ in range(2): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3): f()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps({"pages": n, "case": name, "ms": round(ms, 2), "pages_per_sec": round(n / ms * 1e3, 1)}), flush=True)
    del buf, out
