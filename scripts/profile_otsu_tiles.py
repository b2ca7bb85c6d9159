import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prlib_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rows, cols = 3508, 2480
ctx = prlib_b200.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
buf = torch.empty((n, rows, cols), dtype=torch.uint8, device="cuda"); out = torch.empty_like(buf)
ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, cols, rows * cols, 2024, 0)
for _ in range(2):
    ctx.otsu_tiles_batch_dev(buf.data_ptr(), n, rows, cols, cols, rows * cols, 64, 64, 255.0, out.data_ptr(), cols, rows * cols)
torch.cuda.synchronize()
