"""CPU suite, part 3: the N>1 host logic (page sharding + max-over-ranks timing) on world_size-2 gloo."""
import os
import socket

import numpy as np
import pytest

from prlib_b200.sharding import shard_range, shard_sizes


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 8, 255, 256, 8192):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(r, world, n) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = shard_sizes(world, n)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(2, 2, 10)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_pages, q):
    import hashlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import c_oracle as CO
    lo, hi = shard_range(rank, world, n_pages)
    # each rank "processes" its own pages (tiny synthetic pages through the CPU checker) ...
    digests = {}
    for p in range(lo, hi):
        page = CO.synth_page(p, 96, 128)
        digests[p] = hashlib.sha1(CO.binarize_local(page, 0, 15, (0.2,), 0).tobytes()).hexdigest()
    # ... and only metadata crosses ranks: counts and the max-over-ranks step time (bench.py's reduction)
    t = torch.tensor([float(rank + 1) * 10.0])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, digests)
    dist.barrier()
    if rank == 0:
        q.put((float(t.item()), int(cnt.item()), gathered))
    dist.destroy_process_group()


def test_two_rank_gloo_page_sharding_matches_single_rank():
    import hashlib
    import torch.multiprocessing as mp
    from oracle import c_oracle as CO
    n_pages, world = 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pages, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, total, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 20.0 and total == n_pages
    merged = {}
    for d in gathered:
        assert not (set(d) & set(merged))          # no page processed twice
        merged.update(d)
    assert sorted(merged) == list(range(n_pages))
    for p in range(n_pages):                        # identical to the 1-rank result, page by page
        page = CO.synth_page(p, 96, 128)
        assert merged[p] == hashlib.sha1(CO.binarize_local(page, 0, 15, (0.2,), 0).tobytes()).hexdigest()
