"""Env sharding across the GPUs of one box, and the one reduction the job needs.

Envs are independent games, so the path shards with NO data-path collective:
rank r owns the contiguous global env ids [r*E, (r+1)*E) and keys each env's
Philox stream with its GLOBAL id, which makes every game -- and therefore the
reduced counters -- independent of the number of GPUs.  The only exchange is a
final all-reduce of a handful of int64 counters (NCCL on GPUs, gloo in the
CPU tests of this logic).
"""
from __future__ import annotations

from typing import Dict, Tuple

SUM_KEYS = ("plies", "games", "red_wins", "blue_wins", "draws", "swaps", "kernel_launches")
MAX_KEYS = ("max_length",)


def shard_range(global_envs: int, world: int, rank: int) -> Tuple[int, int]:
    """(first global env id, count) of `rank`'s shard; shards differ by at most one env."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank: %d/%d" % (world, rank))
    # the C ABI's helper (twixt_shard_range), so that a C++ host and this module cut the env range identically
    import ctypes as C
    from . import _lib
    first, count = C.c_int64(), C.c_int64()
    if _lib.load().twixt_shard_range(int(global_envs), world, rank, C.byref(first), C.byref(count)) != 0:
        raise ValueError(_lib.load().twixt_last_error().decode())
    return int(first.value), int(count.value)


def reduce_counters(counters: Dict[str, int], dist=None, device=None) -> Dict[str, int]:
    """All-reduce the playout counters over the default process group (SUM / MAX per key).

    `dist` is torch.distributed (initialised) or None for a single process."""
    out = dict(counters)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return out
    import torch
    sums = torch.tensor([int(counters.get(k, 0)) for k in SUM_KEYS], dtype=torch.int64, device=device)
    maxs = torch.tensor([int(counters.get(k, 0)) for k in MAX_KEYS], dtype=torch.int64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    for k, v in zip(SUM_KEYS, sums.tolist()):
        out[k] = int(v)
    for k, v in zip(MAX_KEYS, maxs.tolist()):
        out[k] = int(v)
    return out
