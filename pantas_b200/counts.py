"""Result of one augment run in the flat layout of include/pantas_aug.h, and its
reduction across shards.

Every field is a commutative reduction over GAF records (sum of counters, min
of first-touch stamps, union-by-key of the two side tables), which is what makes
the byte-range sharding of the GAF exact (DESIGN.md "multi-GPU").
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

STAMP_UNSET = (1 << 63) - 1


@dataclass
class FlatResult:
    """What one context (one GPU / one shard) exports."""
    n_nodes: int
    n_edges: int
    sums: np.ndarray      # int64[3N + E + 4]
    stamps: np.ndarray    # int64[2N]
    novel: np.ndarray     # uint64[n, 3] rows {key, count, stamp}
    sparse: np.ndarray    # uint64[m, 3]


def merge_side(rows_list) -> np.ndarray:
    """Union by key: counts add, stamps take the minimum."""
    rows_list = [r.reshape(-1, 3) for r in rows_list if r.size]
    if not rows_list:
        return np.zeros((0, 3), dtype=np.uint64)
    rows = np.concatenate(rows_list, axis=0)
    keys, inv = np.unique(rows[:, 0], return_inverse=True)
    cnt = np.zeros(keys.shape[0], dtype=np.int64)
    np.add.at(cnt, inv, rows[:, 1].astype(np.int64))
    st = np.full(keys.shape[0], STAMP_UNSET, dtype=np.int64)
    np.minimum.at(st, inv, rows[:, 2].astype(np.int64))
    return np.stack([keys, cnt.astype(np.uint64), st.astype(np.uint64)], axis=1)


def merge_flat(parts) -> FlatResult:
    """Reduce per-shard results on the host (single-process multi-shard runs and tests)."""
    p0 = parts[0]
    sums = p0.sums.copy()
    stamps = p0.stamps.copy()
    for p in parts[1:]:
        sums += p.sums
        np.minimum(stamps, p.stamps, out=stamps)
    return FlatResult(p0.n_nodes, p0.n_edges, sums, stamps,
                      merge_side([p.novel for p in parts]), merge_side([p.sparse for p in parts]))


@dataclass
class Counts:
    """Decoded view used by the GFA writer (pantas_b200.gfa.write_augmented)."""
    nc: np.ndarray
    il0: np.ndarray        # IL[v][0]
    ol_len: np.ndarray     # OL[v][len(v)]
    il0_stamp: np.ndarray
    ol_stamp: np.ndarray
    rc: np.ndarray
    novel: np.ndarray
    sparse: np.ndarray
    rej: int
    n_lines: int

    @classmethod
    def from_flat(cls, r: FlatResult) -> "Counts":
        n, e = r.n_nodes, r.n_edges
        nc = r.sums[:n]
        return cls(nc=nc, il0=nc + r.sums[n:2 * n], ol_len=nc + r.sums[2 * n:3 * n],
                   il0_stamp=r.stamps[:n], ol_stamp=r.stamps[n:2 * n], rc=r.sums[3 * n:3 * n + e],
                   novel=r.novel.reshape(-1, 3), sparse=r.sparse.reshape(-1, 3),
                   rej=int(r.sums[3 * n + e]), n_lines=int(r.sums[3 * n + e + 1]))
