"""Runs every secondary kernel family once or twice on a small batch (for ncu captures).
usage: profile_ops.py [pages] [what ...]   what in {tiles, otsu, morph, fused, a3, wj, adaptive}"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import prlib_b200
from prlib_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
what = sys.argv[2:] or ["tiles", "otsu", "morph", "fused", "a3", "wj"]
ctx = prlib_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
rows, cols = 3508, 2480


def pages(n, rows, cols):
    step = (cols + 15) // 16 * 16
    buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
    ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
    return buf, step


def local(method, params, window, n, rows, cols, morph=0, reps=2):
    buf, step = pages(n, rows, cols)
    rc, orow, ocol = ctx.output_shape(method, rows, cols, window)
    ostep = (ocol + 15) // 16 * 16
    out = torch.empty((n, orow, ostep), dtype=torch.uint8, device="cuda")
    for _ in range(reps):
        ctx.binarize_local_batch_dev(method, buf.data_ptr(), n, rows, cols, step, rows * step, window, params, morph,
                                     out.data_ptr(), ostep, orow * ostep)
    torch.cuda.synchronize()


if "tiles" in what or "otsu" in what:
    buf, step = pages(n, rows, cols)
    out = torch.empty_like(buf)
    thr = torch.zeros(n, dtype=torch.int32, device="cuda")
    for _ in range(2):
        if "tiles" in what:
            ctx.otsu_tiles_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 64, 64, 255.0, out.data_ptr(), step, rows * step)
        if "otsu" in what:
            ctx.otsu_global_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 255.0, out.data_ptr(), step, rows * step, thr.data_ptr())
    torch.cuda.synchronize()
if "morph" in what:
    local(capi.SAUVOLA, (0.2,), 15, n, rows, cols, morph=2)
if "fused" in what:
    ctx.set_option("enable_fused", 1)
    local(capi.SAUVOLA, (0.2,), 15, n, rows, cols)
    ctx.set_option("enable_fused", 0)
if "wj" in what:
    local(capi.WOLFJOLION, (0.5,), 15, n, rows, cols)
if "a3" in what:
    local(capi.NICK, (-0.1,), 101, max(n // 8, 2), 9921, 7016)
if "adaptive" in what:
    buf, step = pages(n, rows, cols)
    out = torch.empty_like(buf)
    native = dict(gray_first=1, blur=1, blur_ksize=5, assert_ksize=1, method=1, type=1, maxval=255.0, check_maxval=1, block_size=19, auto_block=1,
                  delta=9.0, invert_if_dark=1)
    for kw in (native, dict(native, method=0, bilateral_d=5, bilateral_sigma_color=150.0, bilateral_sigma_space=150.0)):
        for _ in range(2):
            ctx.binarize_adaptive_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 1, out.data_ptr(), step, rows * step, **kw)
    torch.cuda.synchronize()
print("done")
