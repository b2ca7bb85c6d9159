"""Global Otsu: does the atomics-bound histogram pass of one chunk overlap with the DRAM-bound apply pass of another?
Two contexts on two streams, chunks interleaved with a half-chunk phase shift."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import prlib_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rows, cols = 3508, 2480
step = (cols + 15) // 16 * 16
ctxs = [prlib_b200.Context(0), prlib_b200.Context(0)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for c, s in zip(ctxs, streams): c.set_stream(s.cuda_stream)
buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
ctxs[0].synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
out = torch.empty_like(buf)
thr = torch.zeros(n, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()

def call(ctx, p0, m):
    ctx.otsu_global_batch_dev(buf[p0].data_ptr(), m, rows, cols, step, rows * step, 255.0, out[p0].data_ptr(), step, rows * step, thr[p0:].data_ptr())

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream())
    for s in streams: s.wait_event(e0)
    for _ in range(reps): fn()
    for s in streams:
        e = torch.cuda.Event(); e.record(s); torch.cuda.default_stream().wait_event(e)
    e1.record(torch.cuda.default_stream()); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

print(f"one stream: {timed(lambda: call(ctxs[0], 0, n)):.3f} ms per {n} pages", flush=True)
for chunk in (n // 4, n // 8, n // 16):
    def both():
        h = n // 2
        qa = list(range(0, h, chunk)); qb = [0] + list(range(chunk // 2, h, chunk))
        for i in range(max(len(qa), len(qb))):
            if i < len(qa): call(ctxs[0], qa[i], min(chunk, h - qa[i]))
            if i < len(qb):
                e = qb[i + 1] if i + 1 < len(qb) else h
                call(ctxs[1], h + qb[i], e - qb[i])
    print(f"two streams, chunks of {chunk}: {timed(both):.3f} ms per {n} pages", flush=True)
