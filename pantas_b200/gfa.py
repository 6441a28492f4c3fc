"""GFA side of `pantas augment`: graph tables in, augmented GFA out.

Host code stays Python (BASELINE.json north_star).  Two functions mirror the two
GFA passes of the reference, /root/reference/scripts/alignments_augmentation_from_gaf.py
(REF:n below):

* ``load_graph``      REF:121-126 (``nodes_info``) plus the key set that REF:421
                      looks up, turned into the dense tables the device wants;
* ``write_augmented`` REF:377-427, byte for byte (tags appended to the stripped
                      original line, file order kept, bare ``L`` dropped, novel
                      links last in first-seen order).

The reference never imports ``gfautils`` on this path and ``gfautils.GFA.print``
re-orders records, so it is not used here either.
"""
from __future__ import annotations

import sys
from dataclasses import dataclass, field

import numpy as np

from .errors import PantasDataError, UnsupportedInput

LEN_ABSENT = 0xFFFFFFFF
MAX_ID = 0xFFFFFFFE
POS_BIAS = 1 << 30
STAMP_UNSET = (1 << 63) - 1


def _canonical_id(tok: str) -> int:
    """Decimal id as vg / build/annotate.cpp write it, else -1.

    Node identity in the reference is *string* identity (REF:126, 214): "07" and
    "7" are different nodes.  The device table is indexed by the integer, so only
    spellings that are their own canonical form may enter it.
    """
    if tok.isascii() and tok.isdigit() and (tok == "0" or tok[0] != "0") and len(tok) <= 10:
        v = int(tok)
        if v <= MAX_ID:
            return v
    return -1


@dataclass
class Graph:
    """Dense device-facing view of the annotated spliced pangenome."""
    min_id: int
    node_len: np.ndarray            # uint32[N]; LEN_ABSENT where no S line has that id
    edge_keys: np.ndarray           # uint64[E]; from_idx << 32 | to_idx, distinct, first-occurrence order
    link_edge: np.ndarray           # int64[#L lines]; edge index that L line prints, -1 if it prints 0
    n_s_lines: int = 0
    path: str = ""
    extra: dict = field(default_factory=dict)

    @property
    def n_nodes(self) -> int:
        return int(self.node_len.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.edge_keys.shape[0])


def load_graph(gfa_file: str, max_span_factor: float = 64.0) -> Graph:
    """First GFA pass (REF:121-126) + link table.

    Raises PantasDataError where the reference raises, UnsupportedInput for
    graphs whose S ids are not canonical non-negative decimals or whose id range
    is too sparse for a direct-indexed table.
    """
    ids: list[int] = []
    lens: list[int] = []
    l_from: list[int] = []
    l_to: list[int] = []
    with open(gfa_file, "r") as f:
        for line in f:
            if line.startswith("S"):                       # REF:123 (raw line)
                tokens = line.strip().split()
                try:
                    sid, seq = tokens[1], tokens[2]
                except IndexError:
                    raise PantasDataError("GFA S line with fewer than 3 fields (reference: IndexError)") from None
                v = _canonical_id(sid)
                if v < 0:
                    raise UnsupportedInput(f"segment id {sid!r} is not a canonical decimal integer")
                ids.append(v)
                lens.append(len(seq))
                continue
            s = line.strip()
            if s.startswith("L") and len(s) > 1:           # REF:417-421 will look this key up
                tokens = s.split()
                if len(tokens) >= 4:
                    l_from.append(_canonical_id(tokens[1]))
                    l_to.append(_canonical_id(tokens[3]))
                else:
                    l_from.append(-2)                      # the writer raises at this line, like REF:421
                    l_to.append(-2)
    if not ids:
        # no node at all: any record that passes the filters is a KeyError; keep one absent slot
        node_len = np.full(1, LEN_ABSENT, dtype=np.uint32)
        min_id = 0
    else:
        ida = np.asarray(ids, dtype=np.int64)
        min_id = int(ida.min())
        span = int(ida.max()) - min_id + 1
        if span > max(max_span_factor * len(ids), 1 << 20):
            raise UnsupportedInput(f"segment ids span {span} values for {len(ids)} segments: too sparse")
        node_len = np.full(span, LEN_ABSENT, dtype=np.uint32)
        lena = np.asarray(lens, dtype=np.int64)
        if lena.max() >= POS_BIAS:
            raise UnsupportedInput("segment longer than 2^30 bases")
        node_len[ida - min_id] = lena.astype(np.uint32)    # later S lines overwrite earlier ones (REF:126)
    n = node_len.shape[0]
    fa = np.asarray(l_from, dtype=np.int64) - min_id
    ta = np.asarray(l_to, dtype=np.int64) - min_id
    raw_from = np.asarray(l_from, dtype=np.int64)
    raw_to = np.asarray(l_to, dtype=np.int64)
    ok = (raw_from >= 0) & (raw_to >= 0) & (fa >= 0) & (fa < n) & (ta >= 0) & (ta < n)
    if ok.any():
        okf = np.where(ok, fa, 0)
        okt = np.where(ok, ta, 0)
        ok &= (node_len[okf] != LEN_ABSENT) & (node_len[okt] != LEN_ABSENT)
    keys = (fa.astype(np.uint64) << np.uint64(32)) | ta.astype(np.uint64)
    link_edge = np.full(len(l_from), -1, dtype=np.int64)
    if ok.any():
        pos = np.flatnonzero(ok)
        uniq, first = np.unique(keys[pos], return_index=True)
        order = np.argsort(first, kind="stable")             # distinct keys in first-occurrence order
        edge_keys = uniq[order]
        link_edge[pos[first[order]]] = np.arange(order.shape[0])   # weights.pop(): only the first L line gets the count
    else:
        edge_keys = np.zeros(0, dtype=np.uint64)
    link_edge[raw_from == -2] = -2
    return Graph(min_id=min_id, node_len=node_len, edge_keys=edge_keys.astype(np.uint64), link_edge=link_edge,
                 n_s_lines=len(ids), path=gfa_file)


def _hist_strings(graph: Graph, counts) -> tuple[dict, dict]:
    """Per node with deletion-derived keys: the ordered 'pos.count' lists (REF:391-394)."""
    il: dict[int, list] = {}
    ol: dict[int, list] = {}
    if counts.sparse.shape[0]:
        key = counts.sparse[:, 0]
        idx = (key >> np.uint64(32)).astype(np.int64)
        dirn = ((key >> np.uint64(31)) & np.uint64(1)).astype(np.int64)
        pos = (key & np.uint64(0x7FFFFFFF)).astype(np.int64) - POS_BIAS
        cnt = counts.sparse[:, 1].astype(np.int64)
        stamp = counts.sparse[:, 2].astype(np.int64)
        for k in range(key.shape[0]):
            d = il if dirn[k] == 0 else ol
            d.setdefault(int(idx[k]), []).append((int(stamp[k]), int(pos[k]), int(cnt[k])))
    return il, ol


def write_augmented(gfa_file: str, graph: Graph, counts, out=None) -> None:
    """Second GFA pass, REF:377-427."""
    out = out or sys.stdout
    nc = counts.nc
    il0 = counts.il0
    oln = counts.ol_len
    il0_stamp = counts.il0_stamp
    ol_stamp = counts.ol_stamp
    rc = counts.rc
    node_len = graph.node_len
    min_id = graph.min_id
    n = graph.n_nodes
    link_edge = graph.link_edge
    sp_il, sp_ol = _hist_strings(graph, counts)
    write = out.write
    k_link = 0
    with open(gfa_file, "r") as f:
        for line in f:
            line = line.strip()
            if line.startswith("S"):
                tokens = line.split()
                try:
                    sid = tokens[1]
                except IndexError:
                    raise PantasDataError("GFA S line without an id (reference: IndexError)") from None
                v = _canonical_id(sid)
                i = v - min_id
                if v < 0 or i < 0 or i >= n or node_len[i] == LEN_ABSENT:
                    raise PantasDataError(f"segment {sid!r} not in the node table (reference: KeyError)")
                if len(tokens) < 3:
                    continue                                # REF:395-416: neither branch prints
                a = int(il0[i])
                b = int(oln[i])
                if i in sp_il or i in sp_ol:
                    ent = list(sp_il.get(i, ()))
                    if a:
                        ent.append((int(il0_stamp[i]), 0, a))
                    ent.sort()
                    in_l = ",".join(f"{p}.{c}" for _, p, c in ent)
                    ent = list(sp_ol.get(i, ()))
                    if b:
                        ent.append((int(ol_stamp[i]), int(node_len[i]), b))
                    ent.sort()
                    out_l = ",".join(f"{p}.{c}" for _, p, c in ent)
                else:
                    in_l = f"0.{a}" if a else ""
                    out_l = f"{node_len[i]}.{b}" if b else ""
                s = f"{line}\tNC:i:{nc[i]}"
                if in_l:
                    s += f"\tIL:Z:{in_l}"
                if out_l:
                    s += f"\tOL:Z:{out_l}"
                write(s + "\n")
            elif line.startswith("L"):
                if len(line) == 1:
                    continue                                # REF:418-419
                e = link_edge[k_link]
                k_link += 1
                if e == -2:
                    raise PantasDataError("GFA L line with fewer than 4 fields (reference: IndexError)")
                w = int(rc[e]) if e >= 0 else 0             # REF:421 weights.pop(key, 0)
                write(f"{line}\tRC:i:{w}\n")
            else:
                write(line + "\n")                          # REF:424
    # REF:426-427: links seen in alignments but absent from the GFA, first-seen order
    if counts.novel.shape[0]:
        order = np.argsort(counts.novel[:, 2].astype(np.int64), kind="stable")
        for r in order:
            key = int(counts.novel[r, 0])
            a = (key >> 32) + min_id
            b = (key & 0xFFFFFFFF) + min_id
            write(f"L\t{a}\t+\t{b}\t+\t*\tRC:i:{int(counts.novel[r, 1])}\tID:Z:N\n")
