"""`pantas augment` -- host driver with the reference's argv / stdout / stderr.

Mirrors main() of /root/reference/scripts/alignments_augmentation_from_gaf.py
(REF:110-427): same positional arguments ``gaf gfa [thr=20]``, the augmented GFA
on stdout, the same four progress lines on stderr, and on malformed input a
non-zero exit with nothing on stdout.  The per-line loop REF:138-371 runs on the
GPU through include/pantas_aug.h; there is no CPU fallback.

Multi-GPU: launch under torchrun (one process per GPU); each rank parses its
byte range of the GAF, rank 0 writes the GFA.  Or set PANTAS_GPUS=N and the
driver re-launches itself that way.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from .counts import Counts, FlatResult
from .errors import PantasDataError, UnsupportedInput
from .gfa import Graph, load_graph, write_augmented
from .shard import shard_bounds


def _last_newline(arr: np.ndarray, n: int) -> int:
    """Index of the last '\\n' in arr[:n], -1 if none."""
    w = 1 << 12
    hi = n
    while hi > 0:
        lo = max(0, hi - w)
        nz = np.flatnonzero(arr[lo:hi] == 10)
        if nz.size:
            return lo + int(nz[-1])
        hi = lo
        w *= 4
    return -1


_READERS = 4


def _pread_into(fd: int, view, offset: int) -> int:
    """pread() until `view` is full or the file ends; -> bytes read (the GIL is released during the system calls)."""
    got, want = 0, len(view)
    while got < want:
        r = os.preadv(fd, [view[got:]], offset + got)
        if r <= 0:
            break
        got += r
    return got


def _read_parallel(pool, fd: int, view, offset: int) -> int:
    """Fill `view` (a memoryview of pinned memory) from file offset `offset` with a few concurrent readers: one thread copies
    out of the page cache at 3-4 GB/s, less than a tenth of what the H2D stream behind it moves.  -> contiguous bytes read."""
    want = len(view)
    if want < (8 << 20) or pool is None:
        return _pread_into(fd, view, offset)
    part = -(-want // _READERS)
    part = (part + 4095) & ~4095
    jobs = []
    for a in range(0, want, part):
        b = min(a + part, want)
        jobs.append((b - a, pool.submit(_pread_into, fd, view[a:b], offset + a)))
    got, short = 0, False
    for n, j in jobs:
        r = j.result()
        if not short:
            got += r
            short = r < n
    return got


def stream_gaf_range(engine, gaf_file: str, lo: int, hi: int, thr: int) -> None:
    """Feed records starting in [lo, hi) of the file to the engine, double-buffered
    through two pinned host buffers (file read || H2D || kernels)."""
    from concurrent.futures import ThreadPoolExecutor

    import torch

    stage = engine.stage_bytes
    bufs = [torch.empty(stage, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    views = [b.numpy() for b in bufs]
    tickets = [None, None]
    carry = np.zeros(0, dtype=np.uint8)
    k = 0
    pos = lo                      # file offset of the first byte not yet handed to the engine
    fpos = lo                     # file offset of the next byte to read
    with open(gaf_file, "rb", buffering=0) as f, ThreadPoolExecutor(_READERS) as pool:
        fd = f.fileno()
        remaining = hi - lo
        while remaining > 0 or carry.size:
            if tickets[k] is not None:
                engine.wait_copy(tickets[k])
            v = views[k]
            c = carry.size
            if c:
                v[:c] = carry
            want = min(stage - c, remaining)
            got = _read_parallel(pool, fd, memoryview(v)[c:c + want], fpos) if want else 0
            fpos += got
            if got < want:
                remaining = got           # file shorter than expected
            remaining -= got
            n = c + got
            if n == 0:
                break
            if remaining > 0:
                cut = _last_newline(v, n) + 1
                if cut <= 0:
                    raise UnsupportedInput(f"GAF record longer than the {stage >> 20} MiB staging buffer "
                                           "(raise PANTAS_STAGE_MB)")
            else:
                cut = n
            carry = v[cut:n].copy()
            tickets[k] = engine.process_host(bufs[k].data_ptr(), cut, pos, thr)
            pos += cut
            k ^= 1
    engine.sync()


def is_stream(gaf_file: str) -> bool:
    """gzip / bgzip files and stdin ("-") cannot be split by byte range: they are read front to back."""
    return gaf_file == "-" or gaf_file.endswith((".gz", ".bgz"))


def stream_gaf_sequential(engine, gaf_file: str, thr: int, rank: int = 0, world: int = 1) -> None:
    """Compressed / piped GAF (`vg mpmap ... | gzip`): the stream is inflated on the host into
    the two pinned staging buffers, cut at line ends, and chunk k goes to rank k % world (every rank reads the whole
    stream, which has no byte ranges to seek to, and skips the chunks of the others; file offsets are offsets of the
    INFLATED text, so stamps -- and the output -- are the same as for the plain file)."""
    import gzip

    import torch

    stage = engine.stage_bytes
    bufs = [torch.empty(stage, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    views = [b.numpy() for b in bufs]
    tickets = [None, None]
    if gaf_file == "-":
        f = sys.stdin.buffer
        if f.peek(2)[:2] == b"\x1f\x8b":
            f = gzip.GzipFile(fileobj=f, mode="rb")
    else:
        f = gzip.open(gaf_file, "rb")
    carry = np.zeros(0, dtype=np.uint8)
    k = 0
    chunk_no = 0
    pos = 0
    eof = False
    while not eof or carry.size:
        if tickets[k] is not None:
            engine.wait_copy(tickets[k])
            tickets[k] = None
        v = views[k]
        c = carry.size
        if c:
            v[:c] = carry
        got = 0
        while c + got < stage and not eof:
            r = f.readinto(memoryview(v[c + got:stage]))
            if not r:
                eof = True
                break
            got += r
        n = c + got
        if n == 0:
            break
        if not eof:
            cut = _last_newline(v, n) + 1
            if cut <= 0:
                raise UnsupportedInput(f"GAF record longer than the {stage >> 20} MiB staging buffer (raise PANTAS_STAGE_MB)")
        else:
            cut = n
        carry = v[cut:n].copy()
        if chunk_no % world == rank:
            tickets[k] = engine.process_host(bufs[k].data_ptr(), cut, pos, thr)
            k ^= 1
        pos += cut
        chunk_no += 1
    engine.sync()


def augment_file(graph: Graph, gaf_file: str, thr: int = 20, device: int = 0, lo: int | None = None,
                 hi: int | None = None, engine=None, rank: int = 0, world: int = 1):
    """GAF loop on one GPU over [lo, hi) of the file (plain files) or over this rank's chunks (streams) -> engine."""
    from .engine import AugmentEngine

    eng = engine or AugmentEngine(device)
    if eng.graph is not graph:
        eng.set_graph(graph)
    if is_stream(gaf_file):
        stream_gaf_sequential(eng, gaf_file, thr, rank, world)
        return eng
    size = os.path.getsize(gaf_file)
    stream_gaf_range(eng, gaf_file, 0 if lo is None else lo, size if hi is None else hi, thr)
    return eng


def _distributed_env():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    return ws, int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def main(argv, out=None, err=None) -> int:
    out = out or sys.stdout
    err = err or sys.stderr
    gaf_file = argv[0]
    gfa_file = argv[1]
    thr = int(argv[2]) if len(argv) > 2 else 20           # REF:113
    world, rank, local_rank = _distributed_env()
    host_passes = os.environ.get("PANTAS_GFA_PASSES", "device") == "host"     # the Python GFA passes of gfa.py (cross-check)
    import time

    marks = [("start", time.perf_counter())]

    def mark(what):                                        # PANTAS_TIMING=1: seconds per stage on stderr when the run ends
        marks.append((what, time.perf_counter()))

    from .engine import AugmentEngine
    from .gfa_device import DeviceGfa

    mark("imports")

    if world > 1:
        import torch
        import torch.distributed as dist

        # stdout carries the GFA and nothing else: NCCL prints its version banner on fd 1,
        # so park the real stdout and point fd 1 at stderr until the writer runs
        if out is sys.stdout:
            sys.stdout.flush()
            saved_fd = os.dup(1)
            os.dup2(2, 1)
            out = os.fdopen(saved_fd, "w")
        torch.cuda.set_device(local_rank)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        device = local_rank
    else:
        device = int(os.environ.get("PANTAS_DEVICE", "0"))
    if os.environ.get("PANTAS_NUMA", "1") != "0":
        from .numa import bind_to_gpu_node

        bind_to_gpu_node(device)                           # before the first pinned allocation (first touch places the pages)
    eng = AugmentEngine(device)
    mark("context")

    print("Read GFA", file=err) if rank == 0 else None    # REF:120
    dg = None
    if host_passes:
        graph = load_graph(gfa_file)
        eng.set_graph(graph)
    else:
        dg = DeviceGfa.load(eng, gfa_file)                 # REF:121-126 on the device
        dg.set_graph()
        graph = dg.graph
    mark("GFA pass 1")
    print("Augmentation by GAF alignments", file=err) if rank == 0 else None   # REF:134

    if world == 1:
        augment_file(graph, gaf_file, thr, engine=eng)
        eng.check_data_error()
        sums, stamps, novel, sparse = eng.export_device()
        novel_h = novel.cpu().numpy().view(np.uint64).reshape(-1, 3)
        sparse_h = sparse.cpu().numpy().view(np.uint64).reshape(-1, 3)
    else:
        from .dist import ERR_HOST, ERR_NONE, reduce_error, reduce_results, rows_to_host

        # every rank reaches the collectives below, whatever happens to its own share of the GAF
        word = ERR_NONE
        try:
            if is_stream(gaf_file):
                augment_file(graph, gaf_file, thr, engine=eng, rank=rank, world=world)
            else:
                b = shard_bounds(gaf_file, world)
                augment_file(graph, gaf_file, thr, engine=eng, lo=b[rank], hi=b[rank + 1])
            eng.check_data_error()
        except (PantasDataError, UnsupportedInput) as e:
            word = ((e.offset or 0) << 8) | (e.code or ERR_HOST)
        except Exception as e:                             # host-side failure on this rank only
            print(f"rank {rank}: {type(e).__name__}: {e}", file=sys.stderr)
            word = ERR_HOST
        word = reduce_error(word, eng.tdev)
        if word != ERR_NONE:
            code, off = word & 0xFF, word >> 8
            if code == ERR_HOST:
                raise RuntimeError("a rank failed on the host side (see its message above)")
            text = eng.lib.pt_strerror(code).decode()
            exc = PantasDataError if code < 20 else UnsupportedInput
            raise exc(f"GAF byte offset {off}: {text}", code, off)
        sums, stamps, novel, sparse = eng.export_device()
        sums, stamps, novel, sparse = reduce_results(sums, stamps, novel, sparse, graph.n_nodes, dst=0)
        if rank != 0:
            return 0
        novel_h, sparse_h = rows_to_host(novel), rows_to_host(sparse)

    mark("GAF loop + export")
    n, e = graph.n_nodes, graph.n_edges
    rej = int(sums[3 * n + e].item())
    print(f"Rejected alignments: {rej}", file=err)         # REF:375
    print("Annotating GFA", file=err)                      # REF:376
    if dg is not None:
        dg.write(sums, stamps, novel_h, sparse_h, out)     # REF:377-427 on the device
    else:
        flat = FlatResult(n, e, sums.cpu().numpy(), stamps.cpu().numpy(), novel_h, sparse_h)
        write_augmented(gfa_file, graph, Counts.from_flat(flat), out)
    out.flush()
    mark("GFA pass 2 + write")
    if os.environ.get("PANTAS_TIMING"):
        print("timing: " + ", ".join(f"{b[0]} {b[1] - a[1]:.2f} s" for a, b in zip(marks, marks[1:])), file=sys.stderr)
    return 0


def cli(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    gpus = int(os.environ.get("PANTAS_GPUS", "1"))
    if gpus > 1 and "WORLD_SIZE" not in os.environ:
        import subprocess

        cmd = [sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
               "--nnodes=1", f"--nproc-per-node={gpus}", "-m", "pantas_b200.augment"] + list(argv)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # the children import the package from any cwd
        env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        return subprocess.call(cmd, env=env)
    rc = main(argv)
    if rc == 0 and "WORLD_SIZE" not in os.environ and os.environ.get("PANTAS_FAST_EXIT", "1") != "0":
        # everything is written: skip the interpreter's teardown (unpinning ~1 GB of host buffers and destroying the CUDA
        # context costs about a second of a run whose GPU work is two)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return rc


if __name__ == "__main__":
    sys.exit(cli())
