#!/usr/bin/env python3
"""Where the end-to-end time of bench.py's host step goes (GPU box): python tools/e2e_breakdown.py [--pairs N]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from pantas_b200.engine import AugmentEngine  # noqa: E402
from pantas_b200.synth import SynthGraph  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=5_000_000)
a = ap.parse_args()
sg = SynthGraph("dm-full", seed=1002)
buf, n_lines = sg.gaf(a.pairs, first_pair=0)
n = int(buf.shape[0])
pinned = torch.empty(n + 32, dtype=torch.uint8).pin_memory()
pinned[:n] = torch.from_numpy(buf)
eng = AugmentEngine(0)
eng.set_graph(sg.graph())
dev = torch.empty(n + 32, dtype=torch.uint8, device="cuda")


def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


print(f"GAF {n / 1e9:.3f} GB, {n_lines} records")
ms = t(lambda: dev[:n].copy_(pinned[:n], non_blocking=True))
print(f"plain H2D (pinned -> device)      {ms:8.2f} ms  {n / ms / 1e6:6.1f} GB/s")


def host_pass():
    stage = eng.stage_bytes
    view = pinned.numpy()
    pos, base = 0, pinned.data_ptr()
    while pos < n:
        end = min(pos + stage, n)
        if end < n:
            w = view[max(pos, end - (1 << 16)):end]
            end = max(pos, end - (1 << 16)) + int(np.flatnonzero(w == 10)[-1]) + 1
        eng.process_host(base + pos, end - pos, pos, 20)
        pos = end


eng.reset()
host_pass()
eng.reset()
ms = t(lambda: (eng.reset(), host_pass()), reps=2)
print(f"reset + process_host chunks       {ms:8.2f} ms")
ms_r = t(lambda: eng.reset(), reps=2)
print(f"  reset alone                     {ms_r:8.2f} ms")
ms = t(lambda: eng.export_device(), reps=2)
print(f"export_device (fold + export)     {ms:8.2f} ms")
sums, stamps, novel, sparse = eng.export_device()
nb = sum(x.numel() * x.element_size() for x in (sums, stamps, novel, sparse))
ms = t(lambda: [x.cpu() for x in (sums, stamps, novel, sparse)], reps=2)
print(f"D2H .cpu() of {nb / 1e6:.0f} MB (pageable)   {ms:8.2f} ms  {nb / ms / 1e6:6.1f} GB/s")
hp = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in (sums, stamps, novel, sparse)]
ms = t(lambda: [h.copy_(x, non_blocking=True) for h, x in zip(hp, (sums, stamps, novel, sparse))], reps=2)
print(f"D2H into pinned buffers           {ms:8.2f} ms  {nb / ms / 1e6:6.1f} GB/s")
