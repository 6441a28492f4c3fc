"""Test-only driver for tests/hostsim/hostsim.cpp (line_core.cuh compiled for the CPU).

Gives CPU-only tests a stand-in for the device library that runs the SAME
per-record code and exports the SAME flat result layout.  Never imported by
pantas_b200/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from pantas_b200.counts import FlatResult

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "hostsim.cpp")
CORE = os.path.join(os.path.dirname(HERE), "pantas_b200", "csrc", "line_core.cuh")
SO = os.path.join(HERE, "hostsim", "libhostsim.so")


class _Res(ctypes.Structure):
    _fields_ = [("sums", ctypes.POINTER(ctypes.c_int64)), ("stamps", ctypes.POINTER(ctypes.c_int64)),
                ("novel", ctypes.POINTER(ctypes.c_uint64)), ("n_novel", ctypes.c_uint64),
                ("sparse", ctypes.POINTER(ctypes.c_uint64)), ("n_sparse", ctypes.c_uint64),
                ("err_offset", ctypes.c_uint64), ("err_code", ctypes.c_int), ("n_deferred", ctypes.c_uint64)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if (not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(CORE))):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", SO, SRC], check=True)
        lib = ctypes.CDLL(SO)
        lib.hostsim_run.restype = ctypes.c_int
        lib.hostsim_run.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p,
                                    ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(_Res)]
        lib.hostsim_free.argtypes = [ctypes.POINTER(_Res)]
        _lib = lib
    return _lib


def run_hostsim(graph, gaf: bytes, thr: int = 20, file_off: int = 0, tile: int = 0, over: int = 64):
    """-> (FlatResult, err_code, err_offset, n_deferred)"""
    lib = _load()
    node_len = np.ascontiguousarray(graph.node_len, dtype=np.uint32)
    keys = np.ascontiguousarray(graph.edge_keys, dtype=np.uint64)
    res = _Res()
    buf = np.frombuffer(gaf, dtype=np.uint8) if len(gaf) else np.zeros(1, dtype=np.uint8)
    lib.hostsim_run(buf.ctypes.data, len(gaf), file_off, thr, node_len.ctypes.data, node_len.shape[0],
                    graph.min_id, keys.ctypes.data if keys.shape[0] else None, keys.shape[0], tile, over,
                    ctypes.byref(res))
    n, e = node_len.shape[0], keys.shape[0]
    sums = np.ctypeslib.as_array(res.sums, shape=(3 * n + e + 4,)).copy()
    stamps = np.ctypeslib.as_array(res.stamps, shape=(2 * n,)).copy()
    novel = (np.ctypeslib.as_array(res.novel, shape=(3 * res.n_novel,)).copy().reshape(-1, 3)
             if res.n_novel else np.zeros((0, 3), dtype=np.uint64))
    sparse = (np.ctypeslib.as_array(res.sparse, shape=(3 * res.n_sparse,)).copy().reshape(-1, 3)
              if res.n_sparse else np.zeros((0, 3), dtype=np.uint64))
    out = (FlatResult(n, e, sums, stamps, novel, sparse), res.err_code, res.err_offset, res.n_deferred)
    lib.hostsim_free(ctypes.byref(res))
    return out


# ---------------------------------------------------------------- fast path under the CUDA emulator

FSRC = os.path.join(HERE, "hostsim", "fastsim.cpp")
FSO = os.path.join(HERE, "hostsim", "libfastsim.so")
_CSRC = os.path.join(os.path.dirname(HERE), "pantas_b200", "csrc")
FDEPS = [FSRC, os.path.join(HERE, "hostsim", "cuda_emu.h")] + [os.path.join(_CSRC, f) for f in
                                                                ("aug_kernels.cuh", "team_tiles.cuh", "tables.cuh", "line_core.cuh")]


class _FRes(ctypes.Structure):
    _fields_ = _Res._fields_ + [("why", ctypes.c_uint64 * 16)]


_flib = None


def _fload():
    global _flib
    if _flib is None:
        if not os.path.exists(FSO) or os.path.getmtime(FSO) < max(os.path.getmtime(d) for d in FDEPS):
            subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused", "-Wno-unknown-pragmas", "-o", FSO, FSRC],
                           check=True)
        lib = ctypes.CDLL(FSO)
        lib.fastsim_run.restype = ctypes.c_int
        lib.fastsim_run.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p,
                                    ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.POINTER(_FRes)]
        lib.fastsim_free.argtypes = [ctypes.POINTER(_FRes)]
        _flib = lib
    return _flib


def run_fastsim(graph, gaf: bytes, thr: int = 20, file_off: int = 0, geo: int = 0, grid: int = 2):
    """The real kernels (fast path + per-record kernel) under tests/hostsim/cuda_emu.h.
    -> (FlatResult, err_code, err_offset, n_deferred, why[16])"""
    lib = _fload()
    node_len = np.ascontiguousarray(graph.node_len, dtype=np.uint32)
    keys = np.ascontiguousarray(graph.edge_keys, dtype=np.uint64)
    res = _FRes()
    buf = np.frombuffer(gaf, dtype=np.uint8) if len(gaf) else np.zeros(1, dtype=np.uint8)
    lib.fastsim_run(buf.ctypes.data, len(gaf), file_off, thr, node_len.ctypes.data, node_len.shape[0],
                    graph.min_id, keys.ctypes.data if keys.shape[0] else None, keys.shape[0], geo, grid,
                    ctypes.byref(res))
    n, e = node_len.shape[0], keys.shape[0]
    sums = np.ctypeslib.as_array(res.sums, shape=(3 * n + e + 4,)).copy()
    stamps = np.ctypeslib.as_array(res.stamps, shape=(2 * n,)).copy()
    novel = (np.ctypeslib.as_array(res.novel, shape=(3 * res.n_novel,)).copy().reshape(-1, 3)
             if res.n_novel else np.zeros((0, 3), dtype=np.uint64))
    sparse = (np.ctypeslib.as_array(res.sparse, shape=(3 * res.n_sparse,)).copy().reshape(-1, 3)
              if res.n_sparse else np.zeros((0, 3), dtype=np.uint64))
    out = (FlatResult(n, e, sums, stamps, novel, sparse), res.err_code, res.err_offset, res.n_deferred, list(res.why))
    lib.fastsim_free(ctypes.byref(res))
    return out
