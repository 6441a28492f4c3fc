"""Device engine: a thin object over the C ABI (include/pantas_aug.h).

PyTorch is used for device / pinned buffers, the current CUDA stream and, in
multi-GPU runs, torch.distributed; every computation is in libpantas_aug.so.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .counts import FlatResult
from .errors import NativeLibraryError, PantasDataError, UnsupportedInput
from .gfa import Graph


class AugmentEngine:
    """One context = one GPU.  Mirrors the reference's loop state
    (weights / nodes_weights / nodes_info / rej, REF:115-119) on the device."""

    def __init__(self, device: int = 0, use_torch_stream: bool = True):
        import torch

        if not torch.cuda.is_available():
            raise NativeLibraryError("no CUDA device: pantas_b200 runs only on a B200 (sm_100a); there is no CPU path")
        self.torch = torch
        self.lib = _lib.load_library()
        self.device = int(device)
        self._ctx = ctypes.c_void_p()
        rc = self.lib.pt_create(self.device, ctypes.byref(self._ctx))
        if rc != 0:
            raise NativeLibraryError(f"pt_create(device={device}) failed: {self.lib.pt_strerror(rc).decode()}")
        self.graph: Graph | None = None
        self._pinned = {}
        self.tdev = torch.device("cuda", self.device)
        self._torch_stream = bool(use_torch_stream)
        if use_torch_stream:
            with torch.cuda.device(self.tdev):
                self._check(self.lib.pt_set_stream(self._ctx, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))

    # -- plumbing
    def _check(self, rc):
        if rc < 0:
            msg = self.lib.pt_last_error(self._ctx).decode("utf-8", "replace")
            raise NativeLibraryError(f"{self.lib.pt_strerror(int(rc)).decode()}: {msg}")
        return rc

    def close(self):
        if self._ctx:
            self.lib.pt_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- graph (REF:121-126)
    def set_graph(self, graph: Graph, novel_cap: int = 0, sparse_cap: int = 0):
        node_len = np.ascontiguousarray(graph.node_len, dtype=np.uint32)
        keys = np.ascontiguousarray(graph.edge_keys, dtype=np.uint64)
        self._check(self.lib.pt_set_graph(
            self._ctx, node_len.ctypes.data, node_len.shape[0], graph.min_id,
            keys.ctypes.data if keys.shape[0] else None, keys.shape[0], novel_cap, sparse_cap))
        self.graph = graph

    def reset(self):
        self._check(self.lib.pt_reset_counts(self._ctx))

    # -- the GAF loop (REF:138-371)
    def process_device(self, gaf, nbytes: int | None = None, file_offset: int = 0, thr: int = 20):
        """gaf: uint8 CUDA tensor whose storage is readable up to nbytes rounded up to 16."""
        n = int(gaf.numel() if nbytes is None else nbytes)
        self._check(self.lib.pt_process_chunk(self._ctx, ctypes.c_void_p(gaf.data_ptr()), n, file_offset, thr))

    def process_host(self, host_ptr: int, nbytes: int, file_offset: int = 0, thr: int = 20) -> int:
        return self._check(self.lib.pt_process_host(self._ctx, ctypes.c_void_p(host_ptr), nbytes, file_offset, thr))

    def wait_copy(self, ticket: int):
        self._check(self.lib.pt_wait_copy(self._ctx, ticket))

    @property
    def stage_bytes(self) -> int:
        return int(self.lib.pt_stage_bytes(self._ctx))

    def set_stage_bytes(self, n: int):
        self._check(self.lib.pt_set_stage_bytes(self._ctx, n))

    def sync(self):
        self._check(self.lib.pt_sync(self._ctx))

    def check_data_error(self):
        """Raise like the reference would have crashed."""
        off = ctypes.c_uint64()
        code = ctypes.c_int()
        self._check(self.lib.pt_error(self._ctx, ctypes.byref(off), ctypes.byref(code)))
        if code.value:
            text = self.lib.pt_strerror(code.value).decode()
            msg = f"GAF byte offset {off.value}: {text}"
            if code.value < 20:
                raise PantasDataError(msg, code.value, off.value)
            raise UnsupportedInput(msg, code.value, off.value)

    # -- results
    def export_device(self):
        """(sums, stamps, novel, sparse) as CUDA tensors in the pantas_aug.h layout.  Stream-ordered when the context runs
        on torch's current stream (the default): no host synchronisation after the last kernel is enqueued."""
        torch = self.torch
        g = self.graph
        n, e = g.n_nodes, g.n_edges
        sums = torch.empty(3 * n + e + 4, dtype=torch.int64, device=self.tdev)
        stamps = torch.empty(2 * n, dtype=torch.int64, device=self.tdev)
        # the dense export needs no host-side number: enqueue it first, the row counts' round trip overlaps it
        self._check(self.lib.pt_export_dense(self._ctx, ctypes.c_void_p(sums.data_ptr()), sums.numel(),
                                             ctypes.c_void_p(stamps.data_ptr()), stamps.numel()))
        n_novel, n_sparse = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self.lib.pt_finalize(self._ctx, ctypes.byref(n_novel), ctypes.byref(n_sparse)))
        novel = torch.empty((max(n_novel.value, 1), 3), dtype=torch.int64, device=self.tdev)
        sparse = torch.empty((max(n_sparse.value, 1), 3), dtype=torch.int64, device=self.tdev)
        self._check(self.lib.pt_export_side(self._ctx, ctypes.c_void_p(novel.data_ptr()), n_novel.value,
                                            ctypes.c_void_p(sparse.data_ptr()), n_sparse.value))
        if not self._torch_stream:
            self.sync()
        return sums, stamps, novel[: n_novel.value], sparse[: n_sparse.value]

    def export_host(self):
        """(sums, stamps, novel, sparse) copied into pinned host buffers (kept between calls): D2H at PCIe speed."""
        torch = self.torch
        dev = self.export_device()
        host = []
        for k, t in enumerate(dev):
            h = self._pinned.get(k)
            if h is None or h.numel() < t.numel():
                h = torch.empty(max(t.numel(), 1), dtype=t.dtype, pin_memory=True)
                self._pinned[k] = h
            v = h[: t.numel()].view(t.shape)
            v.copy_(t, non_blocking=True)
            host.append(v)
        torch.cuda.current_stream(self.tdev).synchronize()
        return tuple(host)

    def export(self) -> FlatResult:
        sums, stamps, novel, sparse = self.export_host()
        g = self.graph
        # (copies: the pinned buffers are reused by the next export)
        return FlatResult(g.n_nodes, g.n_edges, sums.numpy().copy(), stamps.numpy().copy(),
                          novel.numpy().copy().view(np.uint64).reshape(-1, 3),
                          sparse.numpy().copy().view(np.uint64).reshape(-1, 3))

    # -- measurement helpers (bench.py)
    def timer_start(self):
        self._check(self.lib.pt_timer_start(self._ctx))

    def timer_stop(self) -> float:
        ms = ctypes.c_float()
        self._check(self.lib.pt_timer_stop(self._ctx, ctypes.byref(ms)))
        return float(ms.value)

    def profile(self, on: bool = True):
        self._check(self.lib.pt_profile_enable(self._ctx, 1 if on else 0))

    def kernel_time(self):
        """(summed ms of augment_team_kernel, launches) since the last call."""
        ms, n = ctypes.c_float(), ctypes.c_uint64()
        self._check(self.lib.pt_kernel_time(self._ctx, ctypes.byref(ms), ctypes.byref(n)))
        return float(ms.value), int(n.value)

    WHY = ("long", "columns", "ints", "tags", "cs", "path", "steps_full", "walk", "lines_full", "other")

    def handover_reasons(self) -> dict:
        """Why records left the fast path (diagnostics)."""
        out = (ctypes.c_uint64 * 16)()
        self._check(self.lib.pt_debug_counters(self._ctx, out, 16))
        return {k: int(out[i]) for i, k in enumerate(self.WHY)}

    def kernel_time_split(self):
        """(ms fast-path kernel, ms per-record kernel, launches) since the last call."""
        a, b, n = ctypes.c_float(), ctypes.c_float(), ctypes.c_uint64()
        self._check(self.lib.pt_kernel_time_split(self._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)))
        return float(a.value), float(b.value), int(n.value)

    def stats(self) -> dict:
        a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self.lib.pt_stats(self._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"kernel_launches": a.value, "deferred_lines": b.value, "tiles": c.value}
