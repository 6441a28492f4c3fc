"""Read-support tags as arrays for the consumer of the augmented GFA, `pantas call`.

/root/reference/scripts/call.py re-parses the text `augment` printed: for every S line ``build_attrs`` (call.py:25-69)
turns ``NC:i`` into an int and clusters the ``IL:Z`` / ``OL:Z`` position histograms into at most two (position, count)
pairs -- the entries nearest the smallest and the largest position, each collapsed to its count-weighted mean position
(call.py:15-21) -- plus ``MAXIL`` / ``MAXOL``, the position with the largest count; for every L line ``RC:i`` and the
``ID:Z:N`` mark of a novel junction (call.py:145-181).  ``node_attrs`` / ``link_attrs`` give the same values straight
from the counters, without the text in between.

The clustering is a per-node reduction over a handful of entries and only nodes with deletion-derived keys have more
than one entry, so it runs on the host over exactly those nodes; the dense part is vectorised.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import floor

import numpy as np

from .gfa import LEN_ABSENT, POS_BIAS


def collapse_linkcounts(lc):
    """call.py:15-21 (float arithmetic kept operation for operation: the position is floor(sum_i pos_i * cnt_i / count))."""
    count = sum(x[1] for x in lc)
    r = [(x[0] * x[1]) / count for x in lc]
    return [int(floor(sum(r))), count]


def cluster(entries, d: int = 3):
    """call.py:34-63 for one IL / OL entry list [[pos, count], ...] in printed order."""
    v = [list(x) for x in entries]
    if len(v) >= 2:
        minv = min(v, key=lambda x: x[0])
        maxv = max(v, key=lambda x: x[0])
        if abs(minv[0] - maxv[0]) < d:
            v = [collapse_linkcounts(v)]
        else:
            k1, k2 = [minv.copy()], [maxv.copy()]
            for x in v:
                if x == minv or x == maxv:
                    continue
                if abs(x[0] - minv[0]) < abs(x[0] - maxv[0]):
                    k1.append(x)
                else:
                    k2.append(x)
            v = [collapse_linkcounts(k1), collapse_linkcounts(k2)]
    return v


@dataclass
class NodeAttrs:
    """Per node index (id - min_id); absent nodes have nc = 0 and no entries."""
    nc: np.ndarray                 # int64[N]                       NC
    il: dict                       # idx -> [[pos, count], ...]     IL after clustering, only nodes that print an IL tag
    ol: dict
    max_il: np.ndarray             # int64[N], -1 where no IL tag   MAXIL
    max_ol: np.ndarray


def node_attrs(graph, counts, d: int = 3) -> NodeAttrs:
    """What call.py's build_attrs computes from every S line's NC / IL / OL tags (call.py:25-69)."""
    n = graph.n_nodes
    node_len = np.asarray(graph.node_len).astype(np.int64)
    nc = counts.nc.astype(np.int64)
    il0 = counts.il0.astype(np.int64)
    oln = counts.ol_len.astype(np.int64)
    max_il = np.where(il0 > 0, 0, -1).astype(np.int64)
    max_ol = np.where(oln > 0, node_len, -1).astype(np.int64)
    il = {int(i): [[0, int(il0[i])]] for i in np.flatnonzero(il0 > 0)} if n < (1 << 22) else _Dense(il0, None)
    ol = {int(i): [[int(node_len[i]), int(oln[i])]] for i in np.flatnonzero(oln > 0)} if n < (1 << 22) else _Dense(oln, node_len)
    if counts.sparse.shape[0]:
        key = counts.sparse[:, 0]
        idx = (key >> np.uint64(32)).astype(np.int64)
        dirn = ((key >> np.uint64(31)) & np.uint64(1)).astype(np.int64)
        pos = (key & np.uint64(0x7FFFFFFF)).astype(np.int64) - POS_BIAS
        cnt = counts.sparse[:, 1].astype(np.int64)
        stamp = counts.sparse[:, 2].astype(np.int64)
        per = {}
        for k in range(key.shape[0]):
            per.setdefault((int(idx[k]), int(dirn[k])), []).append((int(stamp[k]), int(pos[k]), int(cnt[k])))
        for (i, dr), ent in per.items():
            if dr == 0 and il0[i] > 0:
                ent.append((int(counts.il0_stamp[i]), 0, int(il0[i])))
            if dr == 1 and oln[i] > 0:
                ent.append((int(counts.ol_stamp[i]), int(node_len[i]), int(oln[i])))
            ent.sort()                                   # printed order = dict insertion order (REF:391-394)
            v = cluster([[p, c] for _, p, c in ent], d)
            best = max(v, key=lambda x: x[1])[0]
            if dr == 0:
                il[i] = v
                max_il[i] = best
            else:
                ol[i] = v
                max_ol[i] = best
    return NodeAttrs(nc=nc, il=il, ol=ol, max_il=max_il, max_ol=max_ol)


class _Dense:
    """dict-like view for large graphs: entry i is [[pos_i, count_i]] unless overridden."""

    def __init__(self, count, pos):
        self.count, self.pos, self.over = count, pos, {}

    def __contains__(self, i):
        return i in self.over or self.count[i] > 0

    def __getitem__(self, i):
        if i in self.over:
            return self.over[i]
        if self.count[i] <= 0:
            raise KeyError(i)
        return [[0 if self.pos is None else int(self.pos[i]), int(self.count[i])]]

    def __setitem__(self, i, v):
        self.over[i] = v


@dataclass
class LinkAttrs:
    rc: np.ndarray                 # int64[E]  RC of the e-th distinct (from, to) key of the GFA's L lines
    novel: np.ndarray              # int64[k, 3] rows (from_idx, to_idx, RC) of the links call.py finds by their ID:Z:N tag, printed order


def link_attrs(graph, counts) -> LinkAttrs:
    """RC per known link and the novel junctions (call.py:166-181)."""
    nv = counts.novel
    if nv.shape[0]:
        order = np.argsort(nv[:, 2].astype(np.int64), kind="stable")
        k = nv[order, 0]
        rows = np.stack([(k >> np.uint64(32)).astype(np.int64), (k & np.uint64(0xFFFFFFFF)).astype(np.int64), nv[order, 1].astype(np.int64)], axis=1)
    else:
        rows = np.zeros((0, 3), dtype=np.int64)
    return LinkAttrs(rc=counts.rc.astype(np.int64), novel=rows)
