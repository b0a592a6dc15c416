"""Host-side particle sharding rules (mirrors set_slice()/do_allgather() in csrc/capi.cu).

No reference counterpart (the reference is single-GPU); the contract is "N-rank result == 1-rank
result".  Every rank owns the contiguous slice [rank*L, min(P, (rank+1)*L)) with L = ceil(P / n_ranks); the gathered
record buffer has n_ranks*L rows of REC doubles, rows >= P are padding.  One all-gather per SVN iteration.
"""
from __future__ import annotations

REC = 40  # doubles per particle record: x 6, b 6, H upper triangle 21, |delta_prev| 1, g 6 (csrc/common.cuh)
REC_X, REC_B, REC_H, REC_DNORM, REC_G = 0, 6, 12, 33, 34


def slice_of(P: int, rank: int, n_ranks: int):
    """-> (lo, hi, L): local particle range and the padded per-rank row count."""
    if n_ranks < 1 or not (0 <= rank < n_ranks):
        raise ValueError("bad rank")
    if P < n_ranks:
        raise ValueError("need at least one particle per rank")
    L = -(-P // n_ranks)
    lo = rank * L
    hi = min(P, lo + L)
    if hi <= lo:
        raise ValueError(f"P={P} leaves rank {rank} of {n_ranks} without particles")
    return lo, hi, L


def tri_index(r: int, c: int) -> int:
    """Row-major packing of the upper triangle of a symmetric 6x6 (tri() in csrc/common.cuh)."""
    if r > c:
        r, c = c, r
    return r * 6 - (r * (r - 1)) // 2 + (c - r)


def allgather_bytes_per_iteration(P: int, n_ranks: int) -> int:
    L = -(-P // n_ranks)
    return L * n_ranks * REC * 8
