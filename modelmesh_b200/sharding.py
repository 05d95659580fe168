"""Host-side sharding arithmetic for multi-GPU runs (one process per GPU, torch.distributed for the plumbing).

Registry (model) sharding: rank r owns the contiguous model range shard_range(r, world, n_models); every rank holds the
whole instance table, so a decision needs no data-path exchange — each rank resolves the decisions of its own models.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(rank: int, world: int, n: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n items for `rank` of `world`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank * n // world, (rank + 1) * n // world


def owner_of(item: np.ndarray, world: int, n: int) -> np.ndarray:
    """Inverse of shard_range: which rank owns each item index."""
    item = np.asarray(item, dtype=np.int64)
    # rank r owns [r*n//w, (r+1)*n//w): the owner is the largest r with r*n//w <= item
    r = (item * world + world - 1) // max(n, 1)
    r = np.minimum(r, world - 1)
    lo = r * n // world
    r = np.where(lo > item, r - 1, r)
    hi = (r + 1) * n // world
    r = np.where(hi <= item, r + 1, r)
    return r.astype(np.int64)


def localize_decisions(dec: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """Decisions whose model lies in [lo, hi), with the model index rebased to the shard."""
    sel = (dec["model"] >= lo) & (dec["model"] < hi)
    out = dec[sel].copy()
    out["model"] -= lo
    return out


# ---- instance sharding: one 64-bit key per (decision, shard), combined by an unsigned minimum ----
# layout (csrc/place_core.cuh shard_key): 63 filter-dropped | 62..46 first rank | 45 open | 44..27 target+3 | 26..9 n_candidates
def combine_shard_keys(keys: np.ndarray) -> np.ndarray:
    """keys: uint64[world, n] -> uint64[n], what ncclAllReduce(ncclMin, ncclUint64) leaves on every rank."""
    return np.min(np.asarray(keys, dtype=np.uint64), axis=0)


def decode_shard_keys(best: np.ndarray):
    """-> (target int32[n], n_candidates int32[n], open bool[n]); open = needs the row-gather pass."""
    best = np.asarray(best, dtype=np.uint64)
    none = best == np.uint64(0xFFFFFFFFFFFFFFFF)
    target = ((best >> np.uint64(27)) & np.uint64(0x3FFFF)).astype(np.int64) - 3
    ncand = ((best >> np.uint64(9)) & np.uint64(0x3FFFF)).astype(np.int64)
    is_open = (((best >> np.uint64(45)) & np.uint64(1)) != 0) & ~none
    target = np.where(none, -1, target)
    ncand = np.where(none | is_open, 0, ncand)
    return target.astype(np.int32), ncand.astype(np.int32), is_open
