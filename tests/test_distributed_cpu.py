"""world_size-2 gloo tests of the N>1 host path: byte-range sharding + the one-shot reduction.

Per-shard results come from tests/hostsim (the kernels' per-record code compiled for
the CPU) in the exact layout the device library exports; the reduction code under test
is pantas_b200/dist.py, the same code the NCCL path runs on the GPU box.
"""
import io
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fuzzgen
from oracle.oracle import run_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, gfa_path, gaf_path, out_path, to_root=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hostsim_util import run_hostsim
    from pantas_b200.counts import Counts
    from pantas_b200.counts import FlatResult
    from pantas_b200.dist import ERR_NONE, allreduce_results, reduce_error, reduce_results, rows_to_host
    from pantas_b200.gfa import load_graph, write_augmented
    from pantas_b200.shard import shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    graph = load_graph(gfa_path)
    b = shard_bounds(gaf_path, world)
    with open(gaf_path, "rb") as f:
        f.seek(b[rank])
        data = f.read(b[rank + 1] - b[rank])
    flat, code, off, _ = run_hostsim(graph, data, 20, file_off=b[rank], tile=4096, over=512)
    word = ((off << 8) | code) if code else ERR_NONE
    word = reduce_error(word, torch.device("cpu"))
    if word != ERR_NONE:
        if rank == 0:
            with open(out_path, "wb") as f:
                f.write(b"ERROR %d" % (word & 0xFF))
        dist.destroy_process_group()
        return
    args = (torch.from_numpy(flat.sums), torch.from_numpy(flat.stamps), torch.from_numpy(flat.novel.view(np.int64)),
            torch.from_numpy(flat.sparse.view(np.int64)))
    if to_root:        # the product's reduction: to rank 0 only
        if to_root == "subset":        # big-graph variant: stamps only where the writer reads them
            import pantas_b200.dist as pdist
            pdist.FULL_STAMPS_BYTES = 0
        sums, stamps, novel, sparse = reduce_results(*args, graph.n_nodes, dst=0)
        res = FlatResult(graph.n_nodes, graph.n_edges, sums.numpy(), stamps.numpy(), rows_to_host(novel), rows_to_host(sparse))
    else:
        res = allreduce_results(*args, graph.n_nodes, graph.n_edges)
    if rank == 0:
        out = io.StringIO()
        c = Counts.from_flat(res)
        write_augmented(gfa_path, graph, c, out)
        with open(out_path, "wb") as f:
            f.write(out.getvalue().encode() + b"\nREJ %d" % c.rej)
    dist.destroy_process_group()


@pytest.mark.parametrize("seed,world,to_root", [(8101, 2, False), (8102, 2, True), (8103, 3, "subset"), (8104, 2, "subset")])
def test_sharded_reduction_matches_single_run(seed, world, to_root, tmp_path):
    gfa, gaf = fuzzgen.make_case(seed, n_nodes=25, n_reads=300, weird=True)
    want = run_oracle(gaf.encode(), gfa.encode())
    assert want.rc == 0
    gp, ap, op = tmp_path / "g.gfa", tmp_path / "a.gaf", tmp_path / "out.gfa"
    gp.write_bytes(gfa.encode())
    ap.write_bytes(gaf.encode())
    port = 29500 + (os.getpid() + seed) % 2000
    mp.spawn(_worker, args=(world, port, str(gp), str(ap), str(op), to_root), nprocs=world, join=True)
    got = op.read_bytes()
    assert got == want.out + b"\nREJ %d" % want.rej


def test_error_on_one_rank_is_raised_by_all(tmp_path):
    gfa, gaf = fuzzgen.make_case(8200, n_nodes=12, n_reads=60)
    gaf = gaf + "bad\tline\n"                       # lands in the last shard only
    gp, ap, op = tmp_path / "g.gfa", tmp_path / "a.gaf", tmp_path / "out.gfa"
    gp.write_bytes(gfa.encode())
    ap.write_bytes(gaf.encode())
    port = 29500 + (os.getpid() + 77) % 2000
    mp.spawn(_worker, args=(2, port, str(gp), str(ap), str(op)), nprocs=2, join=True)
    assert op.read_bytes() == b"ERROR 1"


def test_shard_bounds_are_line_starts(tmp_path):
    from pantas_b200.shard import shard_bounds, shard_bounds_bytes

    _, gaf = fuzzgen.make_case(8300, n_reads=200)
    data = gaf.encode()
    ap = tmp_path / "a.gaf"
    ap.write_bytes(data)
    for world in (1, 2, 3, 8, 64, 1000):
        b = shard_bounds(str(ap), world)
        assert b == shard_bounds_bytes(data, world)
        assert b[0] == 0 and b[-1] == len(data) and b == sorted(b)
        for x in b[1:-1]:
            assert x == len(data) or data[x - 1:x] == b"\n"
