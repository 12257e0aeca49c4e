"""Data-parallel helpers for the image path: images are independent, so inference shards the batch
across ranks with NO data-path collective - one process per GPU, weights replicated, results gathered by the caller.
(The only collective of the whole design is the all-reduce of the tiny LoRA gradients in training.)"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `n_items` for `rank`; sizes differ by at most one, order is preserved."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(items: Sequence, rank: int, world: int) -> Sequence:
    lo, hi = shard_bounds(len(items), rank, world)
    return items[lo:hi]


def gather_in_order(local_results: List, group=None) -> List:
    """all_gather_object of per-rank result lists, concatenated in rank order (host-side result plumbing only)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return list(local_results)
    bucket = [None] * dist.get_world_size(group)
    dist.all_gather_object(bucket, list(local_results), group=group)
    return [r for part in bucket for r in part]


def analyze_distributed(analyzer, image_paths: Sequence[str], group=None, **kw) -> dict:
    """`CachedInteriorAnalyzer.analyze_images_batch` under torchrun (one process per GPU): every rank analyses its contiguous
    shard of `image_paths` on its own engine - no data-path collective - and the per-path result dicts are gathered on the
    host, so every rank returns the same dict the single-process call returns (/root/reference/main.py:371-469)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return analyzer.analyze_images_batch(list(image_paths), **kw)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = list(shard(list(image_paths), rank, world))
    local = analyzer.analyze_images_batch(mine, **kw) if mine else {}
    merged = {}
    for path, res in gather_in_order([(p, local[p]) for p in mine if p in local], group):
        merged[path] = res
    return merged
