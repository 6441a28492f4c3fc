#!/usr/bin/env python3
"""Sweep tile lengths of the fast path on the GPU box: python tools/sweep.py 8192 8704 8992 9216 ...  (PANTAS_TILE_BYTES values)"""
import json
import os
import subprocess
import sys

for cfg in sys.argv[1:]:
    env = dict(os.environ, PANTAS_TILE_BYTES=cfg)
    p = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--no-e2e", "--no-cli", "--steps", "4"], env=env,
                       capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print(cfg, "ms/step %.3f" % d["ms_per_step"], "GB/s %.1f" % d["gaf_gb_per_s"], "deferred", d["deferred_records"],
              flush=True)
    except Exception as e:
        print(cfg, "FAILED", e, p.stderr[-400:], flush=True)
