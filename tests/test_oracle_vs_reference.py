"""Live differential test: oracle vs the reference script (build container only)."""
import os

import pytest

import fuzzgen
from oracle.oracle import reference_available, run_oracle, run_reference

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("seed", range(5000, 5040))
def test_fuzz_against_reference(seed, tmp_path):
    for kind in ("safe", "risky"):
        if kind == "safe":
            gfa, gaf = fuzzgen.make_case(seed, weird=(seed % 2 == 0), crlf=(seed % 5 == 0),
                                         trailing_newline=(seed % 3 != 0))
        else:
            gfa, gaf = fuzzgen.make_risky_case(seed)
        if seed % 2 == 1:
            gaf = fuzzgen.add_quality_tags(gaf, seed)      # bq:Z: tags (FASTQ reads): inert for both regexes of the reference
        gp, ap = tmp_path / f"{kind}.gfa", tmp_path / f"{kind}.gaf"
        gp.write_bytes(gfa.encode())
        ap.write_bytes(gaf.encode())
        ref = run_reference(str(ap), str(gp))
        orc = run_oracle(gaf.encode(), gfa.encode())
        if ref.returncode != 0:
            assert orc.rc == 1, (kind, orc.err)
        else:
            assert orc.rc == 0, (kind, orc.err)
            assert orc.out == ref.stdout
            assert orc.rej == ref.rej
