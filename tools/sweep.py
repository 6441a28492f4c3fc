#!/usr/bin/env python3
"""Sweep kernel launch parameters on the GPU box: python tools/sweep.py 'threads,tile_bytes,over_bytes,list_cap' ..."""
import json
import os
import subprocess
import sys

for cfg in sys.argv[1:]:
    th, tile, over, cap = cfg.split(",")
    env = dict(os.environ, PANTAS_THREADS=th, PANTAS_TILE_BYTES=tile, PANTAS_OVER_BYTES=over, PANTAS_LIST_CAP=cap)
    p = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--no-e2e", "--steps", "3"], env=env,
                       capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print(cfg, "ms/step %.2f" % d["ms_per_step"], "GB/s %.1f" % d["gaf_gb_per_s"], "deferred", d["deferred_records"],
              flush=True)
    except Exception as e:
        print(cfg, "FAILED", e, p.stderr[-400:], flush=True)
