"""Page-level data parallelism: contiguous page ranges per GPU / rank, no collective on the data path."""
from __future__ import annotations


def shard_range(rank: int, world: int, n_pages: int) -> tuple[int, int]:
    """Pages [lo, hi) owned by `rank`: GPU g <- pages [floor(g*N/G), floor((g+1)*N/G))  (SURVEY.md section 8e;
    the same split prl_cuda_binarize_batch uses across its device threads)."""
    if world <= 0 or not (0 <= rank < world) or n_pages < 0:
        raise ValueError("bad rank/world/n_pages")
    return (rank * n_pages) // world, ((rank + 1) * n_pages) // world


def shard_sizes(world: int, n_pages: int) -> list[int]:
    return [shard_range(r, world, n_pages)[1] - shard_range(r, world, n_pages)[0] for r in range(world)]
