#!/usr/bin/env python3
"""A longer differential fuzz than the test suite runs: the emulated fast path (tests/hostsim) against the CPU oracle.

    python tools/fuzz_campaign.py 30000 31000          # seeds; safe cases, quality tags, far links, several shards
    python tools/fuzz_campaign.py 40000 42000 risky    # cases that may make the reference raise

No GPU needed (the same kernels run thread by thread on the CPU); run several ranges in parallel, one process each.
Round 2: 1000 safe + 2000 risky seeds, no mismatch."""
import pathlib
import sys
import tempfile

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import fuzzgen  # noqa: E402
from oracle.oracle import run_oracle  # noqa: E402
from test_fastpath_emulated import pipeline  # noqa: E402

lo, hi = int(sys.argv[1]), int(sys.argv[2])
risky = len(sys.argv) > 3 and sys.argv[3] == "risky"
bad = 0
tmp = pathlib.Path(tempfile.mkdtemp())
for seed in range(lo, hi):
    if risky:
        gfa, gaf = fuzzgen.make_risky_case(seed)
        if seed % 4 == 0:
            gaf = fuzzgen.add_quality_tags(gaf, seed, share=0.6)
        shards = 1 + seed % 2
    else:
        gfa, gaf = fuzzgen.make_case(seed, n_nodes=8 + seed % 60, n_reads=100 + (seed * 7) % 500, weird=(seed % 2 == 0),
                                     crlf=(seed % 5 == 0), trailing_newline=(seed % 4 != 0))
        if seed % 3 == 0:
            gaf = fuzzgen.add_quality_tags(gaf, seed, share=0.5)
        if seed % 7 == 0:
            gfa, gaf = fuzzgen.spread_ids(gfa, gaf, pivot=5 + seed % 10, shift=40000 + seed % 1000)
        shards = 1 + seed % 3
    orc = run_oracle(gaf.encode(), gfa.encode())
    res = pipeline(tmp, gfa, gaf, geo=seed % 3, grid=1 + seed % 4, shards=shards)
    if orc.rc != 0:
        ok = res[0] in ("raise", "unsupported")
    else:
        ok = (res[0] == "ok" and res[1] == orc.out and res[2] == orc.rej) or (risky and res[0] == "unsupported")
    if not ok:
        bad += 1
        print("MISMATCH seed", seed, res[0], orc.rc, flush=True)
print("done", lo, hi, "bad", bad, flush=True)
sys.exit(1 if bad else 0)
