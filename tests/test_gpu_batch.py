"""GPU tests of the batch paths: device-resident batch entry points (torch owns the HBM buffers),
the host batch loader / page dispatcher, and size-independent properties at BASELINE sizes."""
import numpy as np
import pytest

import prlib_b200
from prlib_b200 import capi
from oracle import c_oracle as CO
from oracle import prl_oracle as O
from util import sha

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev_pages(ctx, n, rows, cols, first=0, seed=2024):
    step = (cols + 15) // 16 * 16
    buf = torch.empty((n, rows, step), dtype=torch.uint8, device="cuda:0")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.synth_pages_dev(buf.data_ptr(), n, rows, cols, step, rows * step, seed, first)
    return buf, step


def test_device_synth_matches_oracle_generator(ctx, golden):
    buf, step = _dev_pages(ctx, 3, 3508, 2480)
    torch.cuda.synchronize()
    host = buf.cpu().numpy()
    assert sha(host[0, :, :2480]) == golden["images"]["a4_p0"]["sha1"]
    assert sha(host[1, :, :2480]) == golden["images"]["a4_p1"]["sha1"]
    buf, step = _dev_pages(ctx, 2, 333, 501, first=5, seed=7)
    torch.cuda.synchronize()
    assert np.array_equal(buf.cpu().numpy()[1, :, :501], CO.synth_page(6, 333, 501, seed=7))
    ctx.set_stream(None)


@pytest.mark.parametrize("method,window,params", [(0, 15, (0.2,)), (1, 15, (-0.2,)), (2, 15, (0.5,)), (3, 21, (-0.1,)),
                                                  (4, 21, (0.75, 0.2, 0.03, 2.0))])
def test_batch_dev_matches_oracle_per_page(ctx, method, window, params):
    n, rows, cols = 6, 700, 1000
    buf, step = _dev_pages(ctx, n, rows, cols, first=10)
    rc, orow, ocol = ctx.output_shape(method, rows, cols, window)
    ostep = (ocol + 15) // 16 * 16
    out = torch.zeros((n, orow, ostep), dtype=torch.uint8, device="cuda:0")
    ctx.binarize_local_batch_dev(method, buf.data_ptr(), n, rows, cols, step, rows * step, window, params, 0,
                                 out.data_ptr(), ostep, orow * ostep)
    torch.cuda.synchronize()
    got = out.cpu().numpy()[:, :, :ocol]
    for p in range(n):
        page = CO.synth_page(10 + p, rows, cols)
        assert np.array_equal(got[p], CO.binarize_local(page, method, window, params, 0)), p
    ctx.set_stream(None)


def test_batch_dev_large_batch_single_band_path(ctx, golden):
    # >= 222 pages -> one CTA per page (no band carries): the throughput configuration of kernel 1
    n, rows, cols = 240, 200, 300
    buf, step = _dev_pages(ctx, n, rows, cols)
    rc, orow, ocol = ctx.output_shape(0, rows, cols, 15)
    ostep = (ocol + 15) // 16 * 16
    out = torch.zeros((n, orow, ostep), dtype=torch.uint8, device="cuda:0")
    ctx.binarize_local_batch_dev(0, buf.data_ptr(), n, rows, cols, step, rows * step, 15, (0.2,), 0,
                                 out.data_ptr(), ostep, orow * ostep)
    torch.cuda.synchronize()
    got = out.cpu().numpy()[:, :, :ocol]
    for p in (0, 1, 117, 239):
        assert np.array_equal(got[p], CO.binarize_local(CO.synth_page(p, rows, cols), 0, 15, (0.2,), 0)), p
    ctx.set_stream(None)


@pytest.mark.parametrize("method,window,params", [(0, 15, (0.2,)), (1, 15, (-0.2,)), (3, 15, (-0.1,)), (4, 15, (0.75, 0.2, 0.03, 2.0)),
                                                  (0, 21, (0.34,)), (3, 31, (-0.2,)), (0, 3, (0.2,)), (1, 9, (0.7,))])
def test_fused_small_window_path_equals_planes_path_and_oracle(ctx, method, window, params):
    """Big batch + small window -> the fused strip kernel (integral planes never reach HBM).  It must give the
    masks of kernel 1 + kernel 2 byte for byte, and those of the oracle, including on adversarial pages."""
    n, rows, cols = 150, 260, 1000
    rng = np.random.default_rng(window * 100 + method)
    host = np.stack([CO.synth_page(p, rows, cols) for p in range(n)])
    host[3] = rng.integers(0, 256, (rows, cols), dtype=np.uint8)                       # noise
    host[4] = (np.add.outer(np.arange(rows), np.arange(cols)) % 256).astype(np.uint8)  # ramp
    host[5] = 0                                                                        # all black
    host[6] = 255                                                                      # saturated
    host[7][:, :90] = 0; host[7][:40, :] = 0                                           # scanner-style black border
    host[8] = (rng.integers(0, 2, (rows, cols)) * 255).astype(np.uint8)                # pure black / white
    sparse = np.zeros((rows, cols), np.uint8); sparse[rng.integers(0, rows, 60), rng.integers(0, cols, 60)] = rng.integers(1, 256, 60)
    host[9] = sparse
    step = cols + 8                                                                     # pitch != width, 4-aligned
    buf = torch.zeros((n, rows, step), dtype=torch.uint8, device="cuda:0")
    buf[:, :, :cols] = torch.from_numpy(host).to("cuda:0")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    rc, orow, ocol = ctx.output_shape(method, rows, cols, window)
    outs = []
    for fused_off in (0, 1):
        ctx.set_option("enable_fused", 1 - fused_off)
        out = torch.zeros((n, orow, ocol), dtype=torch.uint8, device="cuda:0")              # dense, odd pitch
        ctx.timing_reset(); ctx.timing_enable(True)
        ctx.binarize_local_batch_dev(method, buf.data_ptr(), n, rows, cols, step, rows * step, window, params, 0,
                                     out.data_ptr(), ocol, orow * ocol)
        torch.cuda.synchronize()
        fam = ctx.timing(); ctx.timing_enable(False)
        assert ("fused" in fam) == (fused_off == 0), fam                                 # the intended path really ran
        outs.append(out.cpu().numpy())
    ctx.set_option("enable_fused", 0)
    assert np.array_equal(outs[0], outs[1])
    for p in list(range(12)) + [77, 149]:
        assert np.array_equal(outs[0][p], CO.binarize_local(host[p], method, window, params, 0)), p
    ctx.set_stream(None)


def test_fused_path_hands_pages_back_when_its_fixup_list_overflows(ctx):
    """Without its FP64 estimate tier the fused path lists every pixel its FP32 estimate cannot settle: with
    fused_page_cap = 0 each such page is redone by kernel 1 + kernel 2, with 128 the brute-force fixup finishes the
    pages that have few of them; the masks stay those of the oracle either way."""
    n, rows, cols = 140, 300, 1000
    rng = np.random.default_rng(5)
    host = rng.integers(0, 256, (n, rows, cols), dtype=np.uint8)
    host[::2] //= 16                                    # dark pages: many thresholds next to a rounding boundary
    buf = torch.from_numpy(host).to("cuda:0")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    rc, orow, ocol = ctx.output_shape(1, rows, cols, 15)
    outs = []
    for cap in (0, 128):
        out = torch.zeros((n, orow, ocol), dtype=torch.uint8, device="cuda:0")
        ctx.set_option("enable_fused", 1); ctx.set_option("fused_page_cap", cap); ctx.set_option("fused_no_tier2", 1)
        before = ctx.fused_redo_count()
        ctx.binarize_local_batch_dev(1, buf.data_ptr(), n, rows, cols, cols, rows * cols, 15, (-0.2,), 0, out.data_ptr(), ocol, orow * ocol)
        torch.cuda.synchronize()
        redone = ctx.fused_redo_count() - before
        assert redone > 0 if cap == 0 else redone < n       # cap 128: some pages are finished by the fixup kernel
        outs.append(out.cpu().numpy())
    ctx.set_option("enable_fused", 0); ctx.set_option("fused_page_cap", 128); ctx.set_option("fused_no_tier2", 0)
    assert np.array_equal(outs[0], outs[1])
    for p in (0, 1, 2, 3, 138, 139):
        assert np.array_equal(outs[0][p], CO.binarize_local(host[p], 1, 15, (-0.2,), 0)), p
    ctx.set_stream(None)


def test_fused_path_with_morphology_tail(ctx):
    n, rows, cols = 140, 300, 1000
    buf, step = _dev_pages(ctx, n, rows, cols, first=20)
    rc, orow, ocol = ctx.output_shape(0, rows, cols, 15)
    out = torch.zeros((n, orow, ocol), dtype=torch.uint8, device="cuda:0")
    ctx.set_option("enable_fused", 1)
    ctx.binarize_local_batch_dev(0, buf.data_ptr(), n, rows, cols, step, rows * step, 15, (0.2,), 2, out.data_ptr(), ocol, orow * ocol)
    torch.cuda.synchronize()
    ctx.set_option("enable_fused", 0)
    got = out.cpu().numpy()
    for p in (0, 70, 139):
        assert np.array_equal(got[p], CO.binarize_local(CO.synth_page(20 + p, rows, cols), 0, 15, (0.2,), 2)), p
    ctx.set_stream(None)


def test_batch_dev_workspace_chunking_is_invisible(ctx):
    n, rows, cols = 9, 600, 800
    buf, step = _dev_pages(ctx, n, rows, cols, first=3)
    rc, orow, ocol = ctx.output_shape(2, rows, cols, 15)
    ostep = (ocol + 15) // 16 * 16
    outs = []
    for limit in (1 << 40, 2 * (rows + 14) * 816 * 16 * 2 + 1000):   # everything at once vs 2 pages per chunk
        ctx.set_workspace_limit(limit)
        out = torch.zeros((n, orow, ostep), dtype=torch.uint8, device="cuda:0")
        ctx.binarize_local_batch_dev(2, buf.data_ptr(), n, rows, cols, step, rows * step, 15, (0.5,), 2,
                                     out.data_ptr(), ostep, orow * ostep)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy()[:, :, :ocol].copy())
    ctx.set_workspace_limit(48 << 30)
    assert np.array_equal(outs[0], outs[1])
    for p in (0, 8):
        assert np.array_equal(outs[0][p], CO.binarize_local(CO.synth_page(3 + p, rows, cols), 2, 15, (0.5,), 2))
    ctx.set_stream(None)


def test_a4_batch_golden_and_properties(ctx, golden):
    """BASELINE config 2 shape (A4, Sauvola/Niblack/WJ w=15): pages 0-1 against golden digests, and
    size-independent properties over the whole batch."""
    n, rows, cols = 16, 3508, 2480
    buf, step = _dev_pages(ctx, n, rows, cols)
    rc, orow, ocol = ctx.output_shape(0, rows, cols, 15)
    ostep = (ocol + 15) // 16 * 16
    out = torch.zeros((n, orow, ostep), dtype=torch.uint8, device="cuda:0")
    ctx.binarize_local_batch_dev(0, buf.data_ptr(), n, rows, cols, step, rows * step, 15, (0.2,), 0,
                                 out.data_ptr(), ostep, orow * ostep)
    torch.cuda.synchronize()
    got = out.cpu().numpy()[:, :, :ocol]
    assert sha(got[0]) == golden["images"]["a4_p0"]["masks"]["sauvola_w15_k0.2"]["sha1"]
    assert sha(got[1]) == golden["images"]["a4_p1"]["masks"]["sauvola_w15_k0.2"]["sha1"]
    assert set(np.unique(got).tolist()) <= {0, 255}
    white = (got == 255).mean(axis=(1, 2))
    assert np.all((white > 0.90) & (white < 0.95))          # ~9.2 % ink on every synthpage-v2 page
    # idempotence of the deterministic pipeline: a second run gives byte-identical masks
    out2 = torch.zeros_like(out)
    ctx.binarize_local_batch_dev(0, buf.data_ptr(), n, rows, cols, step, rows * step, 15, (0.2,), 0,
                                 out2.data_ptr(), ostep, orow * ostep)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    # integral planes at full size: last element == page sum (checksum of checksums)
    pad = 7
    pitch = (cols + 2 * pad + 15) // 16 * 16
    Hp = rows + 2 * pad
    S = torch.empty((2, Hp, pitch), dtype=torch.int64, device="cuda:0")
    Q = torch.empty_like(S)
    ctx.integral_batch_dev(buf.data_ptr(), 2, rows, cols, step, rows * step, pad, S.data_ptr(), Q.data_ptr(), pitch, Hp * pitch)
    torch.cuda.synchronize()
    e = golden["images"]["a4_p0"]["integral_pad7"]
    assert int(S[0, Hp - 1, cols + 2 * pad - 1]) == e["S_last"] and int(Q[0, Hp - 1, cols + 2 * pad - 1]) == e["Q_last"]
    Sh = S[0, :, :cols + 2 * pad].cpu().numpy()
    assert sha(Sh) == e["S_sha1"]
    ctx.set_stream(None)


def test_otsu_batch_dev(ctx, golden):
    n, rows, cols = 5, 3508, 2480
    buf, step = _dev_pages(ctx, n, rows, cols)
    out = torch.zeros((n, rows, step), dtype=torch.uint8, device="cuda:0")
    thr = torch.zeros(n, dtype=torch.int32, device="cuda:0")
    ctx.otsu_global_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 255.0, out.data_ptr(), step, rows * step, thr.data_ptr())
    torch.cuda.synchronize()
    e = golden["images"]["a4_p0"]
    assert int(thr[0]) == e["otsu_global"]["thr"] and sha(out[0, :, :cols].cpu().numpy()) == e["otsu_global"]["sha1"]
    for p in range(1, n):
        page = CO.synth_page(p)
        assert int(thr[p]) == O.otsu_threshold_cv(page)
    out.zero_()
    ctx.otsu_tiles_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, 64, 64, 255.0, out.data_ptr(), step, rows * step)
    torch.cuda.synchronize()
    assert sha(out[0, :, :cols].cpu().numpy()) == e["otsu_tiles64"]
    assert np.array_equal(out[3, :, :cols].cpu().numpy(), O.otsu_tiles(CO.synth_page(3)))
    ctx.set_stream(None)


@pytest.mark.parametrize("tw,th", [(16, 16), (32, 64), (48, 52), (64, 64), (128, 40), (112, 300), (64, 1000)])
def test_tile_otsu_lane_kernel_equals_batched_kernel_and_oracle(ctx, tw, th):
    """16-byte aligned pages with a width that is a multiple of 16 take a lane-per-tile kernel (per-lane histograms in
    shared memory; 64-wide tiles at least 48 high are fed by the bulk-copy ring, "tiles_legacy" = 3 forces that kernel for every
    width up to 64, "tiles_legacy" = 2 forces register-staged loads); "tiles_legacy" = 1 forces the round-1 warp-batched kernel.  Edge tiles (partial width and height), several
    pages per warp batch, maxValue below 255."""
    n, rows, cols = 7, 333, 1200
    host = np.stack([CO.synth_page(40 + p, rows, cols) for p in range(n)])
    host[2] = np.random.default_rng(tw * th).integers(0, 256, (rows, cols), dtype=np.uint8)
    host[3] = 255
    host[4][:, :600] = 0
    buf = torch.from_numpy(host).to("cuda:0")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    outs = {}
    for mv in (255.0, 200.0):
        for legacy in (0, 1, 2, 3):
            ctx.set_option("tiles_legacy", legacy)
            out = torch.zeros((n, rows, cols), dtype=torch.uint8, device="cuda:0")
            ctx.otsu_tiles_batch_dev(buf.data_ptr(), n, rows, cols, cols, rows * cols, tw, th, mv, out.data_ptr(), cols, rows * cols)
            torch.cuda.synchronize()
            outs[(mv, legacy)] = out.cpu().numpy()
        ctx.set_option("tiles_legacy", 0)
        assert np.array_equal(outs[(mv, 0)], outs[(mv, 1)]), (tw, th, mv)
        assert np.array_equal(outs[(mv, 0)], outs[(mv, 2)]), (tw, th, mv)
        assert np.array_equal(outs[(mv, 0)], outs[(mv, 3)]), (tw, th, mv)
        for p in range(n):
            assert np.array_equal(outs[(mv, 0)][p], O.otsu_tiles(host[p], tw, th, mv)), (tw, th, mv, p)
    ctx.set_stream(None)


@pytest.mark.parametrize("rows,cols,tw,th,n", [(70, 16, 16, 16, 3), (130, 64, 64, 64, 5), (65, 80, 64, 64, 40), (1, 4096, 64, 64, 2),
                                              (200, 48, 32, 7, 33), (64, 2048, 64, 64, 1)])
def test_tile_otsu_ring_kernel_run_shapes(ctx, rows, cols, tw, th, n):
    """The ring-fed kernel moves whole runs of tiles adjacent in x with one bulk copy per pixel row: pages one tile wide (every
    lane heads a run), two tiles wide with a narrow second tile, one pixel row high, exactly 32 tiles wide (one run per warp),
    more pages than a warp has lanes -- against the warp-batched kernel and the oracle."""
    rng = np.random.default_rng(rows * cols + tw)
    host = rng.integers(0, 256, (n, rows, cols), dtype=np.uint8)
    host[0] = CO.synth_page(7, rows, cols)
    buf = torch.from_numpy(host).to("cuda:0")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    outs = []
    for legacy in (3, 1):
        ctx.set_option("tiles_legacy", legacy)
        out = torch.zeros((n, rows, cols), dtype=torch.uint8, device="cuda:0")
        ctx.otsu_tiles_batch_dev(buf.data_ptr(), n, rows, cols, cols, rows * cols, tw, th, 255.0, out.data_ptr(), cols, rows * cols)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
    ctx.set_option("tiles_legacy", 0)
    ctx.set_stream(None)
    assert np.array_equal(outs[0], outs[1])
    for p in (0, n - 1):
        assert np.array_equal(outs[0][p], O.otsu_tiles(host[p], tw, th, 255.0)), p


def test_host_batch_dispatcher(golden):
    n, rows, cols = 10, 1200, 1600
    pages = np.stack([CO.synth_page(p, rows, cols) for p in range(n)])
    for method, window, params, morph in ((0, 15, (0.2,), 0), (2, 15, (0.5,), 0), (3, 21, (-0.1,), 1)):
        got = prlib_b200.binarize_batch(pages, method, window, params, morph, devices=[0])
        for p in range(n):
            assert np.array_equal(got[p], CO.binarize_local(pages[p], method, window, params, morph)), (method, p)
    # all visible devices (1 here, 8 on a full box) must give byte-identical masks
    got_all = prlib_b200.binarize_batch(pages, 0, 15, (0.2,), 0, devices=None)
    assert np.array_equal(got_all, prlib_b200.binarize_batch(pages, 0, 15, (0.2,), 0, devices=[0]))


def test_timing_hooks_count_launches(ctx):
    img = CO.synth_page(0, 300, 400)
    ctx.set_option("enable_fused", 0)
    ctx.timing_reset(); ctx.timing_enable(True)
    ctx.binarize_local(img, 0, 15, (0.2,), 0)
    t = ctx.timing()
    ctx.timing_enable(False)
    assert t["integral"]["launches"] == 1 and t["threshold"]["launches"] == 1
    assert t["integral"]["ms"] > 0 and ctx.launch_count() >= 2
    # default path for this window: the fused kernel, its fixup, the device-side overflow list and one (empty) hand-back round
    ctx.set_option("enable_fused", 1)
    ctx.timing_reset(); ctx.timing_enable(True)
    ctx.binarize_local(img, 0, 15, (0.2,), 0)
    t = ctx.timing()
    ctx.timing_enable(False)
    assert t["fused"]["launches"] == 1 and t["fused_fix"]["launches"] == 2
    assert t["integral"]["launches"] == 1 and t["threshold"]["launches"] == 1      # the indirect hand-back launches


def lept1_words(mask):
    """Leptonica 1 bpp words of a 0/255 mask, written the way src/formatConvert.cpp:57-69 fills a PIX:
    SET_DATA_BIT(line, x) sets bit 31 - (x & 31) of word x >> 5 for a dark pixel."""
    r, c = mask.shape
    wpl = (c + 31) // 32
    out = np.zeros((r, wpl), np.uint32)
    ys, xs = np.nonzero(mask == 0)
    np.bitwise_or.at(out, (ys, xs >> 5), (np.uint32(1) << (31 - (xs & 31)).astype(np.uint32)))
    return out


@pytest.mark.gpu
def test_packed_batch_and_device_pack(ctx):
    import torch
    rng = np.random.default_rng(11)
    pages = np.stack([CO.synth_page(p, 300, 421) for p in range(5)])
    pages[4] = rng.integers(0, 256, pages[4].shape, dtype=np.uint8)
    for method, window, params, morph in ((0, 15, (0.2,), 0), (3, 21, (-0.1,), 2)):
        bits = prlib_b200.binarize_batch(pages, method, window, params, morph, devices=[0], packed=True)
        masks = prlib_b200.binarize_batch(pages, method, window, params, morph, devices=[0])
        assert bits.dtype == np.uint32 and bits.shape == (5, masks.shape[1], (masks.shape[2] + 31) // 32)
        for p in range(5):
            assert np.array_equal(masks[p], CO.binarize_local(pages[p], method, window, params, morph))
            assert np.array_equal(bits[p], lept1_words(masks[p])), (method, p)
        assert np.array_equal(prlib_b200.unpack_lept1(bits, masks.shape[2]), masks)
    # device entry point, aligned and odd pitches, widths around the 16/32-byte group boundaries
    for cols in (33, 64, 100, 417, 421):
        m = np.where(rng.random((37, cols)) < 0.4, 0, 255).astype(np.uint8)
        for step in (cols, (cols + 15) // 16 * 16):
            buf = torch.zeros((2, 37, step), dtype=torch.uint8, device="cuda")
            buf[:, :, :cols] = torch.from_numpy(np.stack([m, 255 - m])).cuda()
            wpl = (cols + 31) // 32
            out = torch.full((2, 37, wpl), -1, dtype=torch.int32, device="cuda")
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            ctx.pack_mask_dev(buf.data_ptr(), 2, 37, cols, step, 37 * step, out.data_ptr())
            torch.cuda.synchronize()
            got = out.cpu().numpy().view(np.uint32)
            assert np.array_equal(got[0], lept1_words(m)) and np.array_equal(got[1], lept1_words(255 - m)), (cols, step)
    ctx.set_stream(None)


@pytest.mark.gpu
def test_every_device_of_the_box_and_the_page_dispatcher():
    """Each visible device alone and all of them together give the masks of device 0 (found on a 2-GPU box: a
    process-wide cache of cudaFuncSetAttribute left kernel 1 unconfigured on the second device)."""
    n_dev = capi.load().prl_cuda_device_count()
    pages = np.stack([CO.synth_page(p, 400, 700) for p in range(7)])
    want = prlib_b200.binarize_batch(pages, 0, 15, (0.2,), 0, devices=[0])
    for p in range(7):
        assert np.array_equal(want[p], CO.binarize_local(pages[p], 0, 15, (0.2,), 0))
    for d in range(1, n_dev):
        assert np.array_equal(prlib_b200.binarize_batch(pages, 0, 15, (0.2,), 0, devices=[d]), want), d
        c = prlib_b200.Context(d)
        assert np.array_equal(c.binarize_local(pages[0], 3, 21, (-0.1,), 2), CO.binarize_local(pages[0], 3, 21, (-0.1,), 2))
        assert np.array_equal(c.otsu_tiles(pages[1], 64, 64), O.otsu_tiles(pages[1], 64, 64))
        c.close()
    assert np.array_equal(prlib_b200.binarize_batch(pages, 0, 15, (0.2,), 0, devices=list(range(n_dev))), want)
    assert np.array_equal(prlib_b200.binarize_batch(pages, 2, 15, (0.5,), 1, devices=None),
                          prlib_b200.binarize_batch(pages, 2, 15, (0.5,), 1, devices=[0]))


@pytest.mark.gpu
def test_batch_loader_pinned_pageable_and_concurrent_callers():
    """The ctx-less batch entry point: page-locked buffers from prl_cuda_host_alloc, pageable buffers (bounced through
    library-owned pinned memory, or handed to the driver as they are), forced chunk sizes, and two host threads on
    the same device at once (the cached per-device worker is now taken in turns) all give the same masks."""
    import threading
    n, rows, cols = 9, 500, 731
    pages = np.stack([CO.synth_page(p, rows, cols) for p in range(n)])
    want = np.stack([CO.binarize_local(pages[p], 0, 15, (0.2,), 0) for p in range(n)])
    assert np.array_equal(prlib_b200.binarize_batch(pages, 0, 15, (0.2,), 0, devices=[0]), want)       # pageable, bounced
    pin_in = prlib_b200.PinnedArray(pages.shape); pin_out = prlib_b200.PinnedArray(want.shape)
    pin_in.array[...] = pages
    got = prlib_b200.binarize_batch(pin_in.array, 0, 15, (0.2,), 0, devices=[0], out=pin_out.array)
    assert np.array_equal(got, want)
    try:
        prlib_b200.set_global_option("batch_stage_pageable", 0)
        assert np.array_equal(prlib_b200.binarize_batch(pages, 0, 15, (0.2,), 0, devices=[0]), want)
        prlib_b200.set_global_option("batch_stage_pageable", 1)
        for chunk in (1, 2, 100):
            prlib_b200.set_global_option("batch_chunk_pages", chunk)
            # the masks' way back: as bytes, or as bits expanded by 1 / 4 / 7 host threads (9 chunks of one page: every host slot reused)
            for unpack in (0, 1, 4, 7):
                prlib_b200.set_global_option("batch_unpack_threads", unpack)
                assert np.array_equal(prlib_b200.binarize_batch(pages, 0, 15, (0.2,), 0, devices=[0]), want), (chunk, unpack)
                got = prlib_b200.binarize_batch(pin_in.array, 0, 15, (0.2,), 0, devices=[0], out=pin_out.array)
                assert np.array_equal(got, want), (chunk, unpack)
            prlib_b200.set_global_option("batch_unpack_threads", -1)
            assert np.array_equal(prlib_b200.unpack_lept1(
                prlib_b200.binarize_batch(pages, 0, 15, (0.2,), 0, devices=[0], packed=True), want.shape[2]), want), chunk
    finally:
        prlib_b200.set_global_option("batch_chunk_pages", 0)
        prlib_b200.set_global_option("batch_stage_pageable", 1)
        prlib_b200.set_global_option("batch_unpack_threads", -1)
    with pytest.raises(ValueError):
        prlib_b200.set_global_option("no_such_option", 1)
    # concurrent callers, same device, different methods and shapes
    jobs = [(0, 15, (0.2,), 0, pages), (3, 21, (-0.1,), 1, pages[:5, :300, :400].copy()), (2, 15, (0.5,), 0, pages),
            (1, 31, (-0.2,), 0, pages[:, :411, :].copy())]
    results = [None] * len(jobs)

    def run(i):
        m, w, prm, mo, pg = jobs[i]
        for _ in range(3):
            results[i] = prlib_b200.binarize_batch(pg, m, w, prm, mo, devices=[0])

    threads = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    for t in threads: t.start()
    for t in threads: t.join()
    for i, (m, w, prm, mo, pg) in enumerate(jobs):
        for p in range(pg.shape[0]):
            assert np.array_equal(results[i][p], CO.binarize_local(pg[p], m, w, prm, mo)), (i, p)
    pin_in.close(); pin_out.close()


@pytest.mark.gpu
def test_batch_return_paths_on_odd_shapes():
    """prl_cuda_binarize_batch with the masks back as bits (forced: 3 expansion threads) and as bytes: page shapes around the
    32-pixel word and 16-byte alignment boundaries, more chunks than host slots, every method that has a morphology tail"""
    rng = np.random.default_rng(21)
    try:
        for (n, rows, cols), method, window, params, morph in (((37, 97, 61), 1, 15, (-0.2,), 0), ((13, 130, 33), 0, 15, (0.2,), 1),
                                                               ((20, 64, 97), 3, 21, (-0.1,), 0), ((7, 211, 128), 4, 21, (0.12, 0.25, 0.04, 2.0), 2),
                                                               ((5, 300, 421), 2, 15, (0.5,), 0)):
            pages = rng.integers(0, 256, (n, rows, cols), dtype=np.uint8)
            pages[n // 2] = CO.synth_page(3, rows, cols)
            want = np.stack([CO.binarize_local(pages[p], method, window, params, morph) for p in range(n)])
            for chunk in (1, 4):
                prlib_b200.set_global_option("batch_chunk_pages", chunk)
                for unpack, lag, nt in ((3, 0, 1), (3, 1, 1), (2, 0, 0), (0, 0, 1)):      # bits (both hand-over schemes, both store kinds), bytes
                    prlib_b200.set_global_option("batch_unpack_threads", unpack)
                    prlib_b200.set_global_option("batch_unpack_lag", lag)
                    prlib_b200.set_global_option("batch_unpack_nt", nt)
                    got = prlib_b200.binarize_batch(pages, method, window, params, morph, devices=[0])
                    assert np.array_equal(got, want), ((n, rows, cols), method, chunk, unpack, lag, nt)
    finally:
        prlib_b200.set_global_option("batch_chunk_pages", 0)
        prlib_b200.set_global_option("batch_unpack_threads", -1)
        prlib_b200.set_global_option("batch_unpack_lag", 0)
        prlib_b200.set_global_option("batch_unpack_nt", 1)
