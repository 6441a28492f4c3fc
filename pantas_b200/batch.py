"""Several samples against one graph.

The experiments run `pantas augment` once per replicate on the SAME annotated pangenome
(/root/reference/exps/1-dm-sim/workflow/rules/pantas.smk:107-125: one `pantas_weight` job per sample).  Here the GFA
is read, parsed and put into the device tables once; every GAF then only costs its own pass:

    python -m pantas_b200.batch GFA OUT_DIR [--thr 20] GAF [GAF ...]

writes OUT_DIR/<basename of GAF>.gfa for each sample, byte-identical to what the one-sample script prints.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from .augment import augment_file


def augment_many(gfa_file: str, gaf_files, out_paths, thr: int = 20, device: int = 0, err=None):
    """-> [rejected alignments per sample].  One context, one GFA parse, one graph upload."""
    from .engine import AugmentEngine
    from .gfa_device import DeviceGfa

    err = err or sys.stderr
    eng = AugmentEngine(device)
    print("Read GFA", file=err)
    dg = DeviceGfa.load(eng, gfa_file)
    dg.set_graph()
    g = dg.graph
    rejected = []
    for gaf, outp in zip(gaf_files, out_paths):
        print(f"Augmentation by GAF alignments: {gaf}", file=err)
        eng.reset()                                            # counters, stamps, side tables; the graph stays resident
        augment_file(g, gaf, thr, engine=eng)
        eng.check_data_error()
        sums, stamps, novel, sparse = eng.export_device()
        rej = int(sums[3 * g.n_nodes + g.n_edges].item())
        print(f"Rejected alignments: {rej}", file=err)
        host = dg.render(sums, stamps, novel.cpu().numpy().view(np.uint64).reshape(-1, 3),
                         sparse.cpu().numpy().view(np.uint64).reshape(-1, 3))
        with open(outp, "wb") as f:
            f.write(memoryview(host.numpy()))
        rejected.append(rej)
    eng.close()
    return rejected


def cli(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m pantas_b200.batch")
    ap.add_argument("gfa")
    ap.add_argument("out_dir")
    ap.add_argument("gafs", nargs="+")
    ap.add_argument("--thr", type=int, default=20)
    a = ap.parse_args(argv)
    os.makedirs(a.out_dir, exist_ok=True)
    outs = [os.path.join(a.out_dir, os.path.basename(g).split(".gaf")[0] + ".gfa") for g in a.gafs]
    augment_many(a.gfa, a.gafs, outs, a.thr, int(os.environ.get("PANTAS_DEVICE", "0")))
    return 0


if __name__ == "__main__":
    sys.exit(cli())
