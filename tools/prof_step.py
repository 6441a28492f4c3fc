#!/usr/bin/env python3
"""A short device-resident augment run for ncu: python tools/prof_step.py [--pairs N] [--preset P] [--steps K]

Builds the synthetic graph + GAF (same generator as bench.py), uploads them and runs K passes of
pt_process_chunk.  Prints ms per pass (CUDA events); numbers printed under ncu are not bench values.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pantas_b200.engine import AugmentEngine  # noqa: E402
from pantas_b200.synth import SynthGraph  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=1_000_000)
ap.add_argument("--preset", default="dm-full")
ap.add_argument("--seed", type=int, default=1002)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--ladder", action="store_true", help="ablation ladder: time the pass stopped after every phase (PANTAS_ABLATE)")
a = ap.parse_args()

sg = SynthGraph(a.preset, seed=a.seed)
buf, n_lines = sg.gaf(a.pairs, first_pair=0)
n = int(buf.shape[0])
eng = AugmentEngine(0)
eng.set_graph(sg.graph())
eng.profile(True)
dev = torch.zeros(n + 32, dtype=torch.uint8, device="cuda")
dev[:n] = torch.from_numpy(buf).cuda()
torch.cuda.synchronize()
for i in range(a.steps):
    eng.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.process_device(dev, n, 0, 20)
    e1.record()
    e1.synchronize()
    print(f"pass {i}: {e0.elapsed_time(e1):.3f} ms, {n / e0.elapsed_time(e1) / 1e6:.1f} GB/s, {n_lines} records, {n} bytes")
f, sl, nn = eng.kernel_time_split()
print(f"fast kernel {f / nn:.3f} ms, per-record kernel {sl / nn:.3f} ms (avg of {nn})")
eng.check_data_error()
print(eng.stats(), eng.handover_reasons())
if a.ladder:
    # every tile stops after phase k: TMA only, + scan, + records, + ids, + walk 1, + walk 2, everything (results are wrong on purpose
    # for k != 0; timing only).  A fresh context per rung: PANTAS_ABLATE is read by pt_create.
    names = {1: "tma only", 2: "+ scan", 3: "+ records", 4: "+ ids", 5: "+ walk", 0: "+ fold + count (all)"}
    g = sg.graph()
    for k in (1, 2, 3, 4, 5, 0):
        os.environ["PANTAS_ABLATE"] = str(k)
        e2 = AugmentEngine(0)
        e2.set_graph(g)
        e2.profile(True)
        for i in range(a.steps):
            e2.reset()
            e2.process_device(dev, n, 0, 20)
        e2.sync()
        f, sl, nn = e2.kernel_time_split()
        ms = f / nn
        print(f"ladder {names[k]:22s} {ms:8.3f} ms  {n / ms / 1e6:8.1f} GB/s")
        e2.close()
    os.environ.pop("PANTAS_ABLATE", None)
