import sys, os
sys.path.insert(0, "/root/repo")
import torch
import prlib_b200
ctx = prlib_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
n, rows, cols = 128, 3508, 2480
si = (cols + 15) // 16 * 16
pages = torch.empty((n, rows, si), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(pages.data_ptr(), n, rows, cols, si, rows * si, 2024, 0)
for window in (15, 31, 51, 61, 65, 67, 75, 101, 151, 201):
    rc, orow, ocol = ctx.output_shape(0, rows, cols, window)
    so = (ocol + 15) // 16 * 16
    masks = torch.empty((n, orow, so), dtype=torch.uint8, device="cuda")
    ctx.timing_enable(True)
    for i in range(4):
        if i == 1: ctx.timing_reset()
        ctx.binarize_local_batch_dev(0, pages.data_ptr(), n, rows, cols, si, rows * si, window, (0.2,), 0, masks.data_ptr(), so, orow * so)
    torch.cuda.synchronize()
    t = ctx.timing()
    print(f"w={window}", {k: round(v["ms"] / v["launches"], 3) for k, v in t.items()}, flush=True)
