"""PCIe ceiling for the e2e path: pinned H2D, D2H, and both at once (two streams), 8.7 MB pages."""
import torch, time
n = 128; B = 3508 * 2480
h_in = torch.empty((n, B), dtype=torch.uint8).pin_memory(); h_out = torch.empty((n, B), dtype=torch.uint8).pin_memory()
d_in = torch.empty((n, B), dtype=torch.uint8, device="cuda"); d_out = torch.empty((n, B), dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunk):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for rep in range(3):
        for i in range(0, n, chunk):
            if h2d:
                with torch.cuda.stream(s1): d_in[i:i + chunk].copy_(h_in[i:i + chunk], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[i:i + chunk].copy_(d_out[i:i + chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return 3 * n * B / dt / 1e9
for chunk in (1, 8, 32):
    print(f"chunk {chunk:3d} pages: H2D {run(1, 0, chunk):6.1f} GB/s   D2H {run(0, 1, chunk):6.1f} GB/s   both {run(1, 1, chunk):6.1f} GB/s each way "
          f"= {run(1, 1, chunk) * 1e9 / B:7.0f} pages/s ceiling", flush=True)
