"""Length-balanced greedy (LPT) partition of independent pairs over ranks / devices.

Pairs are independent, so multi-GPU runs shard them with no data-path collective:
sort by estimated in-band cells descending, give each pair to the currently lightest shard.  The same
rule is implemented in C++ inside `ksw_b200_batch_upload` for in-process multi-device batches; this
module is the one-process-per-GPU (torchrun) form used by bench.py.
"""
from __future__ import annotations

import heapq
import numpy as np


def est_cells(qlen: np.ndarray, tlen: np.ndarray, w: int) -> np.ndarray:
    """Cheap work estimate: band width x number of anti-diagonals (exact counts are not needed to balance)."""
    qlen = qlen.astype(np.int64); tlen = tlen.astype(np.int64)
    mn = np.minimum(qlen, tlen)
    width = mn if w < 0 else np.minimum(mn, w + 1)
    return width * np.maximum(qlen + tlen - 1, 0)


def lpt_partition(work: np.ndarray, nshards: int):
    """Return a list of index arrays (one per shard), deterministic for a given input."""
    order = np.argsort(-work, kind="stable")
    if nshards <= 1:
        return [np.sort(order)]
    n = len(order)
    # exact LPT with a heap is O(n log k); for equal-work inputs it degenerates to round-robin
    heap = [(0, s) for s in range(nshards)]
    heapq.heapify(heap)
    assign = np.empty(n, np.int32)
    for i in order:
        load, s = heapq.heappop(heap)
        assign[i] = s
        heapq.heappush(heap, (load + int(work[i]) + 1, s))
    return [np.nonzero(assign == s)[0] for s in range(nshards)]


def shard_for_rank(qlen: np.ndarray, tlen: np.ndarray, w: int, rank: int, world: int) -> np.ndarray:
    return lpt_partition(est_cells(qlen, tlen, w), world)[rank]
