"""Does running kernel 1 (store-bound) of one half-batch next to kernel 2 (load-bound) of the other help?
Two contexts on two streams, each with half of the pages, the second one started half a phase later."""
import sys, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import prlib_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rows, cols, window = 3508, 2480, 15
si = (cols + 15) // 16 * 16
ctxs = [prlib_b200.Context(0), prlib_b200.Context(0)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for c, s in zip(ctxs, streams): c.set_stream(s.cuda_stream)
rc, orow, ocol = ctxs[0].output_shape(0, rows, cols, window)
so = (ocol + 15) // 16 * 16
pages = torch.empty((n, rows, si), dtype=torch.uint8, device="cuda")
ctxs[0].synth_pages_dev(pages.data_ptr(), n, rows, cols, si, rows * si, 2024, 0)
masks = torch.empty((n, orow, so), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()

def run(ctx, p0, np_, reps, chunk):
    for _ in range(reps):
        for q in range(p0, p0 + np_, chunk):
            m = min(chunk, p0 + np_ - q)
            ctx.binarize_local_batch_dev(0, pages[q].data_ptr(), m, rows, cols, si, rows * si, window, (0.2,), 0,
                                         masks[q].data_ptr(), so, orow * so)

def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream()); 
    for s in streams: s.wait_event(e0)
    fn()
    evs = []
    for s in streams:
        e = torch.cuda.Event(); e.record(s); evs.append(e)
    for e in evs: torch.cuda.default_stream().wait_event(e)
    e1.record(torch.cuda.default_stream()); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

reps = 4
t1 = timed(lambda: run(ctxs[0], 0, n, reps, n)) / reps
print(f"one stream, {n} pages per launch: {t1:.3f} ms per {n} pages", flush=True)
for chunk in (n // 2, n // 4, n // 8, n // 16):
    def both():
        # interleave launches of the two contexts from one host thread: A gets chunk 0, then B chunk 0, A chunk 1, ...
        # B runs half a chunk behind A, so that its kernel 1 (stores) meets A's kernel 2 (loads)
        h = n // 2
        for _ in range(reps):
            qa = list(range(0, h, chunk))
            qb = [0] + list(range(chunk // 2, h, chunk))
            for i in range(max(len(qa), len(qb))):
                if i < len(qa):
                    m = min(chunk, h - qa[i])
                    ctxs[0].binarize_local_batch_dev(0, pages[qa[i]].data_ptr(), m, rows, cols, si, rows * si, window, (0.2,), 0,
                                                     masks[qa[i]].data_ptr(), so, orow * so)
                if i < len(qb):
                    e = qb[i + 1] if i + 1 < len(qb) else h
                    ctxs[1].binarize_local_batch_dev(0, pages[h + qb[i]].data_ptr(), e - qb[i], rows, cols, si, rows * si, window, (0.2,), 0,
                                                     masks[h + qb[i]].data_ptr(), so, orow * so)
    t2 = timed(both) / reps
    print(f"two streams, chunks of {chunk} pages: {t2:.3f} ms per {n} pages  ({t1 / t2:.3f}x)", flush=True)
