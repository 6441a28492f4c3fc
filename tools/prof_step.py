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
if os.environ.get("PANTAS_PHASE_CLOCKS"):
    pc = eng.phase_cycles()
    tot = sum(pc.values()) or 1
    print("phase share of CTA time:", {k: round(100.0 * v / tot, 1) for k, v in pc.items()})
