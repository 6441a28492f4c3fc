"""The two GFA passes of `pantas augment` on the device.

* ``DeviceGfa.load``   REF:121-126 (``nodes_info``) + the key set REF:421 looks up: the GFA file goes to the GPU once,
                       ``pt_gfa_parse`` (gfa_kernels.cuh, one thread per line) tokenises it, and the dense node table /
                       distinct link keys in first-occurrence order are put together with device-side index ops;
* ``DeviceGfa.write``  REF:377-427: ``pt_gfa_measure`` sizes every output line, a prefix sum places it, ``pt_gfa_format``
                       writes the stripped line plus its NC / IL / OL / RC tags; the bytes come back in one copy.

Same results and errors as the host passes of ``pantas_b200.gfa`` (which stay as the readable statement of the format
and as a cross-check in the tests); PyTorch is used for device / pinned buffers, the line index (``nonzero``), prefix
sums and ``unique`` -- index plumbing, no byte ever parsed or formatted by it.
"""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

from ._lib import GfaWriterArgs
from .errors import PantasDataError, UnsupportedInput
from .gfa import LEN_ABSENT, POS_BIAS, Graph

ID_INVALID = 0xFFFFFFFF
K_RAW_S, K_STR_S, K_STR_L = 1, 2, 4
NO_ERR = -1          # the error word as int64 (~0)

_PASS1_ERRORS = {
    1: (PantasDataError, "GFA S line with fewer than 3 fields (reference: IndexError)"),
    2: (UnsupportedInput, "segment id is not a canonical decimal integer"),
    3: (UnsupportedInput, "segment longer than 2^30 bases"),
}
_PASS2_ERRORS = {
    4: (PantasDataError, "GFA S line without an id (reference: IndexError)"),
    5: (PantasDataError, "segment not in the node table (reference: KeyError)"),
    6: (PantasDataError, "GFA L line with fewer than 4 fields (reference: IndexError)"),
}


class DeviceGfa:
    """A GFA file resident on one GPU: parsed once, written once per result."""

    def __init__(self, engine, torch_mod=None):
        import torch

        self.torch = torch
        self.eng = engine
        self.dev = engine.tdev
        self.graph: Graph | None = None

    # ------------------------------------------------------------------ pass 1
    @classmethod
    def load(cls, engine, gfa_file: str, max_span_factor: float = 64.0) -> "DeviceGfa":
        import torch

        self = cls(engine)
        # (a plain read + one pageable copy: pinning a buffer of the file's size first costs more than it saves for a
        # file that is read once)
        arr = np.fromfile(gfa_file, dtype=np.uint8)
        n = int(arr.shape[0])
        host = torch.from_numpy(arr) if n else torch.empty(1, dtype=torch.uint8)
        return self._load_from_host(host, n, gfa_file, max_span_factor)

    @classmethod
    def load_bytes(cls, engine, data: bytes, max_span_factor: float = 64.0) -> "DeviceGfa":
        import torch

        self = cls(engine)
        n = len(data)
        host = torch.frombuffer(bytearray(data), dtype=torch.uint8) if n else torch.empty(1, dtype=torch.uint8)
        return self._load_from_host(host, n, "<memory>", max_span_factor)

    def _load_from_host(self, host, n: int, name: str, max_span_factor: float) -> "DeviceGfa":
        torch = self.torch
        dev = self.dev
        lib, ctx = self.eng.lib, self.eng._ctx
        with torch.cuda.device(dev):
            buf = torch.zeros(n + 64, dtype=torch.uint8, device=dev)
            buf[:n].copy_(host[:n])
            b = buf[:n]
            # ---- line index: universal newlines ('\n', '\r\n', lone '\r'), like `for line in open(path, "r")`
            if n:
                brk = b == 10
                cr = b == 13
                if bool(cr.any()):
                    nxt = torch.empty_like(b)
                    nxt[:-1] = b[1:]
                    nxt[-1] = 0
                    brk |= cr & (nxt != 10)
                ends = torch.nonzero(brk).flatten()                      # position of every line break's last byte
                del brk, cr
            else:
                ends = torch.zeros(0, dtype=torch.int64, device=dev)
            n_brk = int(ends.shape[0])
            last_start = int(ends[-1].item()) + 1 if n_brk else 0
            tail = 1 if last_start < n else 0                            # an unterminated last line is a line too
            n_lines = n_brk + tail
            start = torch.empty(n_lines + 1, dtype=torch.int64, device=dev)
            end = torch.empty(max(n_lines, 1), dtype=torch.int64, device=dev)
            start[0] = 0
            if n_brk:
                start[1:n_brk + 1] = ends + 1
                end[:n_brk] = ends
            if tail:
                end[n_brk] = n
                start[n_lines] = n
            self.buf, self.nbytes, self.n_lines, self.start = buf, n, n_lines, start
            # ---- tokens, ids, sequence lengths: one thread per line (gfa_parse_kernel)
            i32 = dict(dtype=torch.int32, device=dev)
            self.a_rel = torch.empty(max(n_lines, 1), **i32)
            self.slen = torch.empty(max(n_lines, 1), **i32)
            self.kind = torch.empty(max(n_lines, 1), **i32)
            self.v1 = torch.empty(max(n_lines, 1), **i32)
            self.v2 = torch.empty(max(n_lines, 1), **i32)
            err = torch.full((1,), NO_ERR, dtype=torch.int64, device=dev)
            p = lambda t: ctypes.c_void_p(t.data_ptr())
            self.eng._check(lib.pt_gfa_parse(ctx, p(buf), p(start), p(end), n_lines, p(self.a_rel), p(self.slen), p(self.kind),
                                             p(self.v1), p(self.v2), p(err)))
            self._raise(err, _PASS1_ERRORS)
            kind = self.kind[:n_lines].long()
            v1 = self.v1[:n_lines].long() & 0xFFFFFFFF
            v2 = self.v2[:n_lines].long() & 0xFFFFFFFF
            line_no = torch.arange(n_lines, dtype=torch.int64, device=dev)
            # ---- node table: id -> length, later S lines overwrite earlier ones (REF:126)
            raw_s = (kind & K_RAW_S) != 0
            ids = v1[raw_s]
            n_s = int(ids.shape[0])
            if n_s == 0:
                min_id = 0
                node_len = torch.full((1,), LEN_ABSENT, dtype=torch.int64, device=dev)
            else:
                min_id = int(ids.min().item())
                span = int(ids.max().item()) - min_id + 1
                if span > max(max_span_factor * n_s, 1 << 20):
                    raise UnsupportedInput(f"segment ids span {span} values for {n_s} segments: too sparse")
                comb = (line_no[raw_s] << 32) | v2[raw_s]
                tmp = torch.full((span,), -1, dtype=torch.int64, device=dev)
                tmp.scatter_reduce_(0, ids - min_id, comb, "amax", include_self=True)
                node_len = torch.where(tmp >= 0, tmp & 0xFFFFFFFF, torch.full_like(tmp, LEN_ABSENT))
            n_nodes = int(node_len.shape[0])
            # ---- links: distinct (from, to) keys in first-occurrence order; the first L line of a key prints its count
            ntok = (kind >> 4) & 7
            is_l = ((kind & K_STR_L) != 0) & ((self.slen[:n_lines].long() & 0xFFFFFFFF) > 1)
            fa = v1 - min_id
            ta = v2 - min_id
            ok = is_l & (ntok >= 4) & (v1 != ID_INVALID) & (v2 != ID_INVALID) & (fa >= 0) & (fa < n_nodes) & (ta >= 0) & (ta < n_nodes)
            okl = torch.nonzero(ok).flatten()
            if okl.shape[0]:
                present = (node_len[fa[okl]] != LEN_ABSENT) & (node_len[ta[okl]] != LEN_ABSENT)
                okl = okl[present]
            link_edge = torch.full((max(n_lines, 1),), -1, dtype=torch.int32, device=dev)
            if okl.shape[0]:
                keys = (fa[okl] << 32) | ta[okl]
                uniq, inv = torch.unique(keys, return_inverse=True)
                first = torch.full((uniq.shape[0],), 1 << 62, dtype=torch.int64, device=dev)
                first.scatter_reduce_(0, inv, okl, "amin", include_self=True)
                order = torch.argsort(first)
                edge_keys = uniq[order]
                rank = torch.empty_like(order)
                rank[order] = torch.arange(order.shape[0], dtype=torch.int64, device=dev)
                is_first = okl == first[inv]
                link_edge[okl[is_first]] = rank[inv[is_first]].to(torch.int32)
            else:
                edge_keys = torch.zeros(0, dtype=torch.int64, device=dev)
            self.link_edge = link_edge
            self.node_len_dev = node_len.to(torch.int32)                  # uint32 bit patterns (LEN_ABSENT = 0xFFFFFFFF)
            self.edge_keys_dev = edge_keys.contiguous()
            self.min_id = min_id
            self.n_s_lines = n_s
            torch.cuda.current_stream(dev).synchronize()
        self.graph = Graph(min_id=min_id, node_len=_LazyHost(self.node_len_dev, np.uint32, n_nodes),
                           edge_keys=_LazyHost(self.edge_keys_dev, np.uint64, int(edge_keys.shape[0])),
                           link_edge=None, n_s_lines=n_s, path=name)      # (the per-line link table stays on the device)
        return self

    def _raise(self, err, table):
        w = int(err.item())
        if w == NO_ERR:
            return
        line, code = (w >> 8) & ((1 << 55) - 1), w & 0xFF
        exc, text = table.get(code, (PantasDataError, f"GFA error {code}"))
        raise exc(f"GFA line {line + 1}: {text}", code, line)

    def set_graph(self, novel_cap: int = 0, sparse_cap: int = 0):
        """Hand the node table and link keys to the engine (device pointers: the library copies device to device)."""
        eng = self.eng
        n_e = int(self.edge_keys_dev.shape[0])
        eng._check(eng.lib.pt_set_graph(eng._ctx, ctypes.c_void_p(self.node_len_dev.data_ptr()), int(self.node_len_dev.shape[0]),
                                        self.min_id, ctypes.c_void_p(self.edge_keys_dev.data_ptr()) if n_e else None, n_e,
                                        novel_cap, sparse_cap))
        eng.graph = self.graph

    # ------------------------------------------------------------------ pass 2
    def render_device(self, sums, stamps, novel: np.ndarray, sparse: np.ndarray):
        """The augmented GFA (REF:377-427) as (uint8 device tensor, its length, bytes of the novel-link lines that follow it).
        sums / stamps: device tensors in the pt_export_dense layout (after the cross-rank reduction, if any); novel / sparse:
        host rows {key, count, stamp}."""
        torch = self.torch
        dev = self.dev
        lib, ctx = self.eng.lib, self.eng._ctx
        n_nodes = int(self.node_len_dev.shape[0])
        with torch.cuda.device(dev):
            # ---- the few nodes with deletion-derived keys: their IL / OL tags in dict insertion order (REF:391-394)
            sp_slot = torch.full((n_nodes,), -1, dtype=torch.int32, device=dev)
            blob = bytearray()
            offs = [0]
            if sparse.shape[0]:
                key = sparse[:, 0]
                idx = (key >> np.uint64(32)).astype(np.int64)
                dirn = ((key >> np.uint64(31)) & np.uint64(1)).astype(np.int64)
                pos = (key & np.uint64(0x7FFFFFFF)).astype(np.int64) - POS_BIAS
                cnt = sparse[:, 1].astype(np.int64)
                stamp = sparse[:, 2].astype(np.int64)
                nodes = np.unique(idx)
                nt = torch.from_numpy(nodes).to(dev)
                nc = sums[nt]
                il0 = (nc + sums[n_nodes + nt]).cpu().numpy()
                oln = (nc + sums[2 * n_nodes + nt]).cpu().numpy()
                ist = stamps[nt].cpu().numpy()
                ost = stamps[n_nodes + nt].cpu().numpy()
                nlen = (self.node_len_dev[nt].long() & 0xFFFFFFFF).cpu().numpy()
                ent_il = {int(v): [] for v in nodes}
                ent_ol = {int(v): [] for v in nodes}
                for k in range(key.shape[0]):
                    (ent_il if dirn[k] == 0 else ent_ol)[int(idx[k])].append((int(stamp[k]), int(pos[k]), int(cnt[k])))
                for j, v in enumerate(nodes):
                    a, b = int(il0[j]), int(oln[j])
                    e = ent_il[int(v)]
                    if a:
                        e.append((int(ist[j]), 0, a))
                    e.sort()
                    s = ""
                    if e:
                        s += "\tIL:Z:" + ",".join(f"{p}.{c}" for _, p, c in e)
                    e = ent_ol[int(v)]
                    if b:
                        e.append((int(ost[j]), int(nlen[j]), b))
                    e.sort()
                    if e:
                        s += "\tOL:Z:" + ",".join(f"{p}.{c}" for _, p, c in e)
                    blob += s.encode()
                    offs.append(len(blob))
                sp_slot[nt] = torch.arange(nodes.shape[0], dtype=torch.int32, device=dev)
            sp_off = torch.tensor(offs, dtype=torch.int64, device=dev)
            sp_text = torch.frombuffer(bytearray(blob) if blob else bytearray(1), dtype=torch.uint8).to(dev)
            p = lambda t: ctypes.c_void_p(t.data_ptr())
            w = GfaWriterArgs(p(self.buf), p(self.start), p(self.a_rel), p(self.slen), p(self.kind), p(self.v1), p(self.link_edge),
                              p(self.node_len_dev), p(sums), p(sp_slot), p(sp_off), p(sp_text), n_nodes, self.n_lines, self.min_id)
            out_len = torch.zeros(max(self.n_lines, 1), dtype=torch.int64, device=dev)
            err = torch.full((1,), NO_ERR, dtype=torch.int64, device=dev)
            self.eng._check(lib.pt_gfa_measure(ctx, ctypes.byref(w), p(out_len), p(err)))
            self._raise(err, _PASS2_ERRORS)
            incl = torch.cumsum(out_len, 0)
            total = int(incl[-1].item()) if self.n_lines else 0
            off = incl - out_len
            out = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
            self.eng._check(lib.pt_gfa_format(ctx, ctypes.byref(w), p(off), p(out), p(err)))
            # REF:426-427: links seen in alignments but absent from the GFA, first-seen order
            extra = bytearray()
            if novel.shape[0]:
                order = np.argsort(novel[:, 2].astype(np.int64), kind="stable")
                for r in order:
                    k = int(novel[r, 0])
                    extra += f"L\t{(k >> 32) + self.min_id}\t+\t{(k & 0xFFFFFFFF) + self.min_id}\t+\t*\tRC:i:{int(novel[r, 1])}\tID:Z:N\n".encode()
        return out, total, bytes(extra)

    def render(self, sums, stamps, novel: np.ndarray, sparse: np.ndarray):
        """-> pinned uint8 tensor with the whole augmented GFA (tests, bench parity check)."""
        torch = self.torch
        out, total, extra = self.render_device(sums, stamps, novel, sparse)
        n = total + len(extra)
        host = torch.empty(n, dtype=torch.uint8, pin_memory=True) if n else torch.empty(0, dtype=torch.uint8)
        if total:
            host[:total].copy_(out[:total], non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        if extra:
            host[total:] = torch.frombuffer(bytearray(extra), dtype=torch.uint8)
        return host

    PIECE = 32 << 20

    def write(self, sums, stamps, novel, sparse, out=None) -> None:
        """Render on the device and stream to `out` through two small pinned buffers: the copy of piece i + 1 runs while
        piece i is written (no pinned buffer of the whole output, whose allocation alone costs more than the copy)."""
        torch = self.torch
        dev_out, total, extra = self.render_device(sums, stamps, novel, sparse)
        out = out or sys.stdout
        raw = getattr(out, "buffer", None)
        if raw is None:                                    # a text-only stream (tests): one copy
            host = dev_out[:total].cpu().numpy().tobytes() + extra
            out.write(host.decode("ascii"))
            return
        out.flush()
        with torch.cuda.device(self.dev):
            piece = self.PIECE
            bufs = [torch.empty(min(piece, max(total, 1)), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            evs = [torch.cuda.Event(), torch.cuda.Event()]
            n_pieces = (total + piece - 1) // piece

            def issue(i):
                a, b = i * piece, min((i + 1) * piece, total)
                bufs[i & 1][: b - a].copy_(dev_out[a:b], non_blocking=True)
                evs[i & 1].record()

            if n_pieces:
                issue(0)
            for i in range(n_pieces):
                evs[i & 1].synchronize()
                if i + 1 < n_pieces:
                    issue(i + 1)                           # (buffer (i + 1) & 1 was written out in the previous iteration)
                a, b = i * piece, min((i + 1) * piece, total)
                raw.write(memoryview(bufs[i & 1].numpy())[: b - a])
        if extra:
            raw.write(extra)
        raw.flush()


class _LazyHost:
    """numpy view of a device tensor, copied on first use (tests and the host writer look at the graph; the product does not)."""

    def __init__(self, t, dtype, n):
        self._t, self._dtype, self._n, self._a = t, dtype, n, None

    def _arr(self):
        if self._a is None:
            a = self._t[: self._n].cpu().numpy()
            self._a = a.view(self._dtype) if a.dtype.itemsize == np.dtype(self._dtype).itemsize else a.astype(self._dtype)
        return self._a

    @property
    def shape(self):
        return (self._n,)

    def __len__(self):
        return self._n

    def __array__(self, dtype=None, copy=None):
        a = self._arr()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, k):
        return self._arr()[k]
