"""Two-kernel path: one context over 256 A4 pages vs several contexts (streams) over page groups side by side (diagnostic)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prlib_b200
n, rows, cols, window = 256, 3508, 2480, 15
step = (cols + 15) // 16 * 16
main = prlib_b200.Context(0)
buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda")
main.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 2024, 0)
rc, orow, ocol = main.output_shape(0, rows, cols, window)
ostep = (ocol + 15) // 16 * 16
out = torch.empty((n, orow, ostep), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for lanes, groups in ((1, 1), (2, 2), (2, 4), (2, 8), (3, 6), (4, 8)):
    ctxs, streams = [], []
    for i in range(lanes):
        c = prlib_b200.Context(0); c.set_option("enable_fused", 0)
        s = torch.cuda.Stream(); c.set_stream(s.cuda_stream)
        ctxs.append(c); streams.append(s)
    per = n // groups
    def run():
        for gi in range(groups):
            c = ctxs[gi % lanes]; p0 = gi * per
            np_ = per if not (gi == 0 and lanes > 1 and groups > lanes) else per
            c.binarize_local_batch_dev(0, buf.data_ptr() + p0 * rows * step, per, rows, cols, step, rows * step, window, (0.2,), 0,
                                       out.data_ptr() + p0 * orow * ostep, ostep, orow * ostep)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record(cur)
    for s in streams: s.wait_event(e0)
    for _ in range(5): run()
    for s in streams:
        ev = torch.cuda.Event(); ev.record(s); cur.wait_event(ev)
    e1.record(cur); torch.cuda.synchronize()
    print(json.dumps({"lanes": lanes, "groups": groups, "ms_per_256_pages": round(e0.elapsed_time(e1) / 5, 3)}), flush=True)
    for c in ctxs: c.set_stream(None); c.close()
