#!/usr/bin/env python3
# python alignments_augmentation_from_gaf.py alignment.gaf input.gfa 20 > output.gfa
"""Drop-in for pantas' scripts/alignments_augmentation_from_gaf.py (same file
name, argv, stdout and stderr), so `pantas:132`, and the snakemake rules that
call the script directly, work unchanged.  The per-line loop runs on a B200
through libpantas_aug.so; see INTEGRATION.md."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pantas_b200.augment import cli  # noqa: E402

if __name__ == "__main__":
    sys.exit(cli(sys.argv[1:]))
