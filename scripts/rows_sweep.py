import sys; sys.path.insert(0, "/root/repo")
import torch, prlib_b200
ctx = prlib_b200.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
n, rows, cols = 256, 3508, 2480
si = (cols + 15) // 16 * 16
pages = torch.empty((n, rows, si), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(pages.data_ptr(), n, rows, cols, si, rows * si, 2024, 0)
for window in (15, 101):
    rc, orow, ocol = ctx.output_shape(0, rows, cols, window)
    so = (ocol + 15) // 16 * 16
    masks = torch.empty((n, orow, so), dtype=torch.uint8, device="cuda"); ref = None
    for rpc in (2, 4, 6, 8, 12, 16):
        ctx.set_option("thr_rows", rpc); ctx.timing_enable(True)
        for i in range(4):
            if i == 1: ctx.timing_reset()
            ctx.binarize_local_batch_dev(0, pages.data_ptr(), n, rows, cols, si, rows * si, window, (0.2,), 0, masks.data_ptr(), so, orow * so)
        torch.cuda.synchronize(); t = ctx.timing()
        if ref is None: ref = masks.clone()
        print(f"w={window} rows/CTA={rpc}", round(t["threshold"]["ms"] / t["threshold"]["launches"], 3), "same", bool(torch.equal(ref, masks)), flush=True)
