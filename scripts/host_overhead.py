import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prlib_b200
n = int(sys.argv[1]); rows, cols = 3508, 2480
ctx = prlib_b200.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
rc, orow, ocol = ctx.output_shape(0, rows, cols, 15)
si = 2480; so = (ocol + 15) // 16 * 16
pages = torch.empty((n, rows, si), dtype=torch.uint8, device="cuda"); masks = torch.empty((n, orow, so), dtype=torch.uint8, device="cuda")
ctx.synth_pages_dev(pages.data_ptr(), n, rows, cols, si, rows * si, 2024, 0); torch.cuda.synchronize()
for i in range(5):
    t0 = time.perf_counter()
    ctx.binarize_local_batch_dev(0, pages.data_ptr(), n, rows, cols, si, rows * si, 15, (0.2,), 0, masks.data_ptr(), so, orow * so)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"n={n} call returned after {1e3*(t1-t0):.2f} ms, gpu done after {1e3*(t2-t0):.2f} ms")
