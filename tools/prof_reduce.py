#!/usr/bin/env python3
"""Where the multi-GPU epilogue spends its time (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/prof_reduce.py

Every rank runs a short augment pass over its own reads, then the stages of the end-of-job reduction are timed one by
one (host clock around a synchronize: no overlap, so the sum is an upper bound of the real epilogue)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pantas_b200 import dist as pdist  # noqa: E402
from pantas_b200.engine import AugmentEngine  # noqa: E402
from pantas_b200.synth import SynthGraph  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=1_000_000)
ap.add_argument("--preset", default="dm-full")
ap.add_argument("--seed", type=int, default=1002)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

sg = SynthGraph(a.preset, seed=a.seed)
buf, n_lines = sg.gaf(a.pairs, first_pair=rank * a.pairs)
n = int(buf.shape[0])
eng = AugmentEngine(local)
g = sg.graph()
eng.set_graph(g)
gaf = torch.zeros(n + 32, dtype=torch.uint8, device=dev)
gaf[:n] = torch.from_numpy(buf).to(dev)
eng.reset()
eng.process_device(gaf, n, 0, 20)
torch.cuda.synchronize()


def timed(label, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
    return r


acc, acc2 = {}, {}
N = g.n_nodes
for it in range(a.iters + 2):
    A = {} if it < 2 else acc
    dist.barrier()
    sums, stamps, novel, sparse = timed("export_device", eng.export_device, A)
    timed("gather_side(novel)", lambda: pdist.gather_side(novel), A)
    sp = timed("gather_side(sparse)", lambda: pdist.gather_side(sparse), A)
    timed("reduce(sums) int64", lambda: dist.reduce(sums, dst=0, op=dist.ReduceOp.SUM), A)
    s2 = sums.clone()
    timed("all_reduce(sums) int64", lambda: dist.all_reduce(s2, op=dist.ReduceOp.SUM), A)
    s32 = timed("cast int32", lambda: sums.to(torch.int32), A)
    timed("reduce(sums) int32", lambda: dist.reduce(s32, dst=0, op=dist.ReduceOp.SUM), A)
    timed("all_reduce(sums) int32", lambda: dist.all_reduce(s32, op=dist.ReduceOp.SUM), A)

    def stamps_subset():
        if sp.shape[0]:
            nodes = torch.unique((sp[:, 0] >> 32) & 0xFFFFFFFF)
            idx = torch.cat([nodes, nodes + N])
            sub = stamps[idx].contiguous()
            dist.reduce(sub, dst=0, op=dist.ReduceOp.MIN)
            if rank == 0:
                stamps[idx] = sub
    timed("stamps subset", stamps_subset, A)
    timed("whole reduce_results", lambda: pdist.reduce_results(sums, stamps, novel, sparse, N), A)
if rank == 0:
    print(f"world {world}, {a.preset}, {N} nodes, sums {sums.numel() * 8 / 1e6:.1f} MB, novel rows {novel.shape[0]}, sparse rows {sparse.shape[0]}")
    for k, v in acc.items():
        print(f"  {k:28s} {v / a.iters:8.3f} ms")
dist.destroy_process_group()
