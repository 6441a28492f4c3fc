"""Host pipeline on the CPU: GFA loader -> per-record logic -> reduction -> GFA writer.

The per-record logic is pantas_b200/csrc/line_core.cuh, the code the CUDA
kernels run, compiled for the CPU by tests/hostsim (test harness).  Expected
outputs come from the reference (tests/golden) and from the oracle.
"""
import io

import pytest

import fuzzgen
from conftest import GOLDEN, UNSUPPORTED_BY_DESIGN
from hostsim_util import run_hostsim
from oracle.oracle import run_oracle
from pantas_b200.counts import Counts, merge_flat
from pantas_b200.errors import PantasDataError, UnsupportedInput
from pantas_b200.gfa import load_graph, write_augmented
from pantas_b200.shard import shard_bounds_bytes


def pipeline(tmp_path, gfa: str, gaf: str, thr=20, tile=0, over=64, shards=1):
    """-> ('ok', stdout_bytes, rej) | ('raise', code) | ('unsupported', code)"""
    gp = tmp_path / "g.gfa"
    gp.write_bytes(gfa.encode())
    try:
        graph = load_graph(str(gp))
    except PantasDataError:
        return ("raise", 0)
    data = gaf.encode()
    bounds = shard_bounds_bytes(data, shards)
    parts = []
    worst = None
    for r in range(shards):
        lo, hi = bounds[r], bounds[r + 1]
        flat, code, off, _ = run_hostsim(graph, data[lo:hi], thr, file_off=lo, tile=tile, over=over)
        if code and (worst is None or (off, code) < worst):
            worst = (off, code)
        parts.append(flat)
    if worst:
        return ("raise" if worst[1] < 20 else "unsupported", worst[1])
    counts = Counts.from_flat(merge_flat(parts))
    out = io.StringIO()
    try:
        write_augmented(str(gp), graph, counts, out)
    except PantasDataError:
        return ("raise", 0)
    return ("ok", out.getvalue().encode(), counts.rej)


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_golden(case, tmp_path):
    thr = 20 if case["thr"] is None else case["thr"]
    res = pipeline(tmp_path, case["gfa"], case["gaf"], thr)
    if case["name"] in UNSUPPORTED_BY_DESIGN:
        assert res[0] == "unsupported", res
    elif case["returncode"] != 0:
        assert res[0] == "raise", res
    else:
        assert res[0] == "ok", res
        assert res[1] == case["stdout"]
        assert res[2] == case["rej"]


@pytest.mark.parametrize("tile,over,shards", [(48, 32, 1), (256, 64, 1), (1024, 16, 3), (0, 64, 2), (64, 4096, 5)])
def test_golden_tiled_and_sharded(tile, over, shards, tmp_path):
    """Tiling, the deferral path and byte-range sharding must not change a byte."""
    for case in GOLDEN:
        if case["returncode"] != 0 or case["name"] in UNSUPPORTED_BY_DESIGN or "\r" in case["gaf"]:
            continue
        thr = 20 if case["thr"] is None else case["thr"]
        res = pipeline(tmp_path, case["gfa"], case["gaf"], thr, tile=tile, over=over, shards=shards)
        assert res[0] == "ok", (case["name"], res)
        assert res[1] == case["stdout"], case["name"]
        assert res[2] == case["rej"]


@pytest.mark.parametrize("seed", range(7000, 7060))
def test_fuzz_vs_oracle(seed, tmp_path):
    gfa, gaf = fuzzgen.make_case(seed, n_nodes=8 + seed % 20, n_reads=60, weird=(seed % 2 == 0),
                                 crlf=(seed % 6 == 0), trailing_newline=(seed % 4 != 0))
    orc = run_oracle(gaf.encode(), gfa.encode())
    res = pipeline(tmp_path, gfa, gaf, tile=[0, 64, 512][seed % 3], over=[64, 16, 128][seed % 3],
                   shards=1 + seed % 4)
    assert orc.rc == 0
    assert res[0] == "ok", res
    assert res[1] == orc.out
    assert res[2] == orc.rej


@pytest.mark.parametrize("seed", range(7100, 7160))
def test_fuzz_risky_vs_oracle(seed, tmp_path):
    gfa, gaf = fuzzgen.make_risky_case(seed)
    orc = run_oracle(gaf.encode(), gfa.encode())
    res = pipeline(tmp_path, gfa, gaf)
    if orc.rc == 0:
        assert res[0] == "ok" and res[1] == orc.out and res[2] == orc.rej
    else:
        assert res[0] == "raise", (res, orc.err)
