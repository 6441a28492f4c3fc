"""Synthetic annotated spliced pangenomes and vg-mpmap-shaped GAFs (bench / test input).

ctypes wrapper over synth.c (gcc, built in-tree).  Not on the augment path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "synth.c")
SO = os.path.join(HERE, "libpantas_synth.so")

# named graph scales of BASELINE.json configs; node counts are approximate
PRESETS = {
    #                 genes, mean sites/gene, bubble prob after a plain site, optional-node fraction, zipf s
    "tiny":           dict(n_genes=40, mean_sites=120, bubble=0.8, opt=0.15, zipf=0.0),
    "dm-chr4":        dict(n_genes=110, mean_sites=300, bubble=0.8, opt=0.15, zipf=0.0),       # ~6e4 nodes
    "dm-full":        dict(n_genes=14000, mean_sites=240, bubble=0.8, opt=0.15, zipf=0.0),     # ~6e6 nodes
    "hs-chr1":        dict(n_genes=5200, mean_sites=1300, bubble=0.8, opt=0.15, zipf=0.0),     # ~1.2e7 nodes
    "hs-wg":          dict(n_genes=60000, mean_sites=1850, bubble=0.8, opt=0.15, zipf=0.0),    # ~2e8 nodes
    "gene-panel":     dict(n_genes=123, mean_sites=450, bubble=0.8, opt=0.15, zipf=1.2),       # ~1e5 nodes, skewed
}


def build_synth(force: bool = False) -> str:
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-pthread", "-o", SO, SRC, "-lm"], check=True)
    return SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_synth())
        lib.synth_create.restype = ctypes.c_void_p
        lib.synth_create.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_double, ctypes.c_uint32]
        lib.synth_destroy.argtypes = [ctypes.c_void_p]
        for f in ("synth_n_nodes", "synth_n_links", "synth_n_sites", "synth_n_transcripts"):
            getattr(lib, f).restype = ctypes.c_uint64
            getattr(lib, f).argtypes = [ctypes.c_void_p]
        lib.synth_fill_tables.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.synth_write_gfa.restype = ctypes.c_int
        lib.synth_write_gfa.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        lib.synth_gaf.restype = ctypes.c_uint64
        lib.synth_gaf.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int,
                                  ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64)]
        lib.synth_free_buf.argtypes = [ctypes.c_void_p]
        _lib = lib
    return _lib


class SynthGraph:
    def __init__(self, preset: str = "tiny", seed: int = 1001, read_len: int = 150, **over):
        p = dict(PRESETS[preset])
        p.update(over)
        self.preset = preset
        self.seed = seed
        self.params = p
        self.lib = _load()
        self.h = self.lib.synth_create(seed, p["n_genes"], p["mean_sites"], p["bubble"], p["opt"], p["zipf"], read_len)
        self.n_nodes = int(self.lib.synth_n_nodes(self.h))
        self.n_links = int(self.lib.synth_n_links(self.h))

    def close(self):
        if self.h:
            self.lib.synth_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tables(self):
        """(node_len uint32[N], edge_keys uint64[E]) -- what load_graph() would return for write_gfa()'s file."""
        node_len = np.empty(self.n_nodes, dtype=np.uint32)
        keys = np.empty(self.n_links, dtype=np.uint64)
        self.lib.synth_fill_tables(self.h, node_len.ctypes.data, keys.ctypes.data)
        return node_len, keys

    def graph(self):
        from ..gfa import Graph

        node_len, keys = self.tables()
        return Graph(min_id=1, node_len=node_len, edge_keys=keys,
                     link_edge=np.arange(keys.shape[0], dtype=np.int64), n_s_lines=self.n_nodes, path="<synthetic>")

    def write_gfa(self, path: str):
        if self.lib.synth_write_gfa(self.h, path.encode()) != 0:
            raise OSError(f"cannot write {path}")

    def gaf(self, n_pairs: int, first_pair: int = 0, threads: int | None = None, out: np.ndarray | None = None):
        """Records for pairs [first_pair, first_pair + n_pairs): 2 lines per pair.
        Returns (uint8 array, n_lines).  If `out` is given the bytes are copied into it."""
        threads = threads or min(32, os.cpu_count() or 1)
        ptr = ctypes.c_void_p()
        nl = ctypes.c_uint64()
        n = self.lib.synth_gaf(self.h, first_pair, n_pairs, threads, ctypes.byref(ptr), ctypes.byref(nl))
        src = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(n,))
        if out is not None:
            out[:n] = src
            res = out[:n]
        else:
            res = src.copy()
        self.lib.synth_free_buf(ptr)
        return res, int(nl.value)
