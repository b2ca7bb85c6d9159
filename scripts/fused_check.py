"""Fused small-window path vs two-kernel path on the bench workload: timing per family, mask equality, redo count."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import prlib_b200
pages_n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rows, cols = 3508, 2480
ctx = prlib_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
si = (cols + 15) // 16 * 16
pages = torch.empty((pages_n, rows, si), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(pages.data_ptr(), pages_n, rows, cols, si, rows * si, 2024, 0)
for method, window, params in [(0, 15, (0.2,)), (1, 15, (-0.2,)), (3, 15, (-0.1,)), (4, 15, (0.75, 0.2, 0.03, 2.0)), (0, 31, (0.2,)), (0, 7, (0.2,))]:
    rc, orow, ocol = ctx.output_shape(method, rows, cols, window)
    so = (ocol + 15) // 16 * 16
    res = []
    for fused in (1, 0):
        ctx.set_option("enable_fused", fused)
        masks = torch.zeros((pages_n, orow, so), dtype=torch.uint8, device="cuda")
        ctx.timing_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(4):
            if i == 1:
                ctx.timing_reset(); e0.record()
            ctx.binarize_local_batch_dev(method, pages.data_ptr(), pages_n, rows, cols, si, rows * si, window, params, 0,
                                         masks.data_ptr(), so, orow * so)
        e1.record(); torch.cuda.synchronize()
        t = ctx.timing()
        print(method, window, "fused" if fused else "planes", {k: round(v["ms"] / v["launches"], 3) for k, v in t.items()},
              "ms/step", round(e0.elapsed_time(e1) / 3, 3), "redo", ctx.fused_redo_count(), flush=True)
        res.append(masks)
    print("   equal:", bool(torch.equal(res[0], res[1])), flush=True)
