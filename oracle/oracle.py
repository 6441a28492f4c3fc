"""Python access to the CPU oracle and (in the build container only) the reference.

TEST INFRASTRUCTURE.  Only tests/, ``__graft_entry__.smoke()`` and bench.py's
cpu_baseline / ``--impl reference`` legs may import this module; nothing under
``pantas_b200/`` does.

* ``run_oracle`` calls ``oracle/liboracle.so`` (built from augment_oracle.c, the
  C restatement of /root/reference/scripts/alignments_augmentation_from_gaf.py).
* ``run_reference`` executes the reference script itself as a subprocess
  (that is how the experiments call it); it exists only
  where ``/root/reference`` is mounted, i.e. never on the GPU box.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from dataclasses import dataclass

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SCRIPT = "/root/reference/scripts/alignments_augmentation_from_gaf.py"


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with gcc (idempotent)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "augment_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "all"], check=True)
    return so


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.oracle_augment_mem.restype = ctypes.c_int
        lib.oracle_augment_mem.argtypes = [
            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
            ctypes.c_longlong, ctypes.c_int,
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
            ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong),
            ctypes.POINTER(ctypes.c_double), ctypes.c_char_p, ctypes.c_size_t,
        ]
        lib.oracle_free_buf.argtypes = [ctypes.c_void_p]
        lib.oracle_free_buf.restype = None
        lib.oracle_graph_load.restype = ctypes.c_void_p
        lib.oracle_graph_load.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_size_t]
        lib.oracle_graph_free.argtypes = [ctypes.c_void_p]
        lib.oracle_graph_free.restype = None
        lib.oracle_graph_nodes.argtypes = [ctypes.c_void_p]
        lib.oracle_graph_nodes.restype = ctypes.c_longlong
        lib.oracle_augment_graph.restype = ctypes.c_int
        lib.oracle_augment_graph.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_longlong, ctypes.c_int,
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
            ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong),
            ctypes.POINTER(ctypes.c_double), ctypes.c_char_p, ctypes.c_size_t,
        ]
        _lib = lib
    return _lib


@dataclass
class OracleResult:
    rc: int            # 0 ok, 1 reference would raise, 2 not modelled
    out: bytes         # augmented GFA (empty unless rc == 0 and write_output)
    rej: int
    n_lines: int
    gaf_seconds: float  # time spent in the GAF loop only
    err: str


def _as_ptr(b):
    """bytes / bytearray / numpy uint8 array / (ptr, n) -> (void*, n, keepalive)."""
    if isinstance(b, tuple):
        return ctypes.c_void_p(b[0]), b[1], None
    if isinstance(b, bytes):
        return ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p), len(b), b
    if isinstance(b, bytearray):
        arr = (ctypes.c_char * len(b)).from_buffer(b)
        return ctypes.cast(arr, ctypes.c_void_p), len(b), arr
    # numpy uint8 array
    return ctypes.c_void_p(b.ctypes.data), b.nbytes, b


def run_oracle(gaf, gfa, thr: int = 20, write_output: bool = True) -> OracleResult:
    lib = _load()
    gp, gn, k1 = _as_ptr(gaf)
    fp, fn, k2 = _as_ptr(gfa)
    out = ctypes.c_void_p()
    out_n = ctypes.c_size_t()
    rej = ctypes.c_longlong()
    nl = ctypes.c_longlong()
    secs = ctypes.c_double()
    err = ctypes.create_string_buffer(256)
    rc = lib.oracle_augment_mem(gp, gn, fp, fn, thr, 1 if write_output else 0,
                                ctypes.byref(out), ctypes.byref(out_n), ctypes.byref(rej),
                                ctypes.byref(nl), ctypes.byref(secs), err, 256)
    data = b""
    if rc == 0 and out.value:
        data = ctypes.string_at(out.value, out_n.value)
        lib.oracle_free_buf(out)
    del k1, k2
    return OracleResult(rc, data, rej.value, nl.value, secs.value, err.value.decode("ascii", "replace"))


class OracleGraph:
    """A GFA parsed once by the oracle (REF:121-126); ``run`` is the GAF loop (+ the writer) against it with private
    counters, so several host threads can time the loop on their own byte ranges without re-reading the GFA."""

    def __init__(self, gfa):
        lib = _load()
        fp, fn, keep = _as_ptr(gfa)
        rc = ctypes.c_int()
        err = ctypes.create_string_buffer(256)
        self._h = lib.oracle_graph_load(fp, fn, ctypes.byref(rc), err, 256)
        del keep
        self.rc = rc.value
        self.err = err.value.decode("ascii", "replace")
        self.n_nodes = int(lib.oracle_graph_nodes(self._h))

    def run(self, gaf, thr: int = 20, write_output: bool = True) -> OracleResult:
        lib = _load()
        gp, gn, keep = _as_ptr(gaf)
        out = ctypes.c_void_p()
        out_n = ctypes.c_size_t()
        rej = ctypes.c_longlong()
        nl = ctypes.c_longlong()
        secs = ctypes.c_double()
        err = ctypes.create_string_buffer(256)
        rc = lib.oracle_augment_graph(self._h, gp, gn, thr, 1 if write_output else 0, ctypes.byref(out), ctypes.byref(out_n),
                                      ctypes.byref(rej), ctypes.byref(nl), ctypes.byref(secs), err, 256)
        data = b""
        if rc == 0 and out.value:
            data = ctypes.string_at(out.value, out_n.value)
            lib.oracle_free_buf(out)
        del keep
        return OracleResult(rc, data, rej.value, nl.value, secs.value, err.value.decode("ascii", "replace"))

    def close(self):
        if self._h:
            _load().oracle_graph_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


REF_COPY = os.path.join(_HERE, "_ref", "alignments_augmentation_from_gaf.py")


def reference_script():
    """The reference script itself: /root/reference in the build container, the copy `make -C oracle ref` put under
    oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot) elsewhere.  None if neither exists."""
    for p in (REFERENCE_SCRIPT, REF_COPY):
        if os.path.exists(p):
            return p
    return None


def reference_available() -> bool:
    return os.path.exists(REFERENCE_SCRIPT)


@dataclass
class ReferenceResult:
    returncode: int
    stdout: bytes
    stderr: bytes

    @property
    def rej(self):
        for line in self.stderr.decode("utf-8", "replace").splitlines():
            if line.startswith("Rejected alignments:"):
                return int(line.split(":")[1])
        return None


def run_reference(gaf_path: str, gfa_path: str, thr=None, timeout: float = 600.0) -> ReferenceResult:
    """python3 <reference script> gaf gfa [thr]  (pantas:132)."""
    if not reference_available():
        raise RuntimeError("reference not mounted at /root/reference")
    cmd = [sys.executable, "-W", "ignore", REFERENCE_SCRIPT, gaf_path, gfa_path]
    if thr is not None:
        cmd.append(str(thr))
    p = subprocess.run(cmd, capture_output=True, timeout=timeout)
    return ReferenceResult(p.returncode, p.stdout, p.stderr)
