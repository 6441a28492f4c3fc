"""The fast path (pantas_b200/csrc/team_tiles.cuh) and the kernels around it, run thread by thread on
the CPU by tests/hostsim/cuda_emu.h (test harness) and compared with the reference's golden outputs and
the oracle.  Same code as the sm_100a build; the GPU parity tests (-m gpu) cover the real thing.
"""
import io

import pytest

import fuzzgen
from conftest import GOLDEN, UNSUPPORTED_BY_DESIGN
from hostsim_util import run_fastsim
from oracle.oracle import run_oracle
from pantas_b200.counts import Counts, merge_flat
from pantas_b200.errors import PantasDataError
from pantas_b200.gfa import load_graph, write_augmented
from pantas_b200.shard import shard_bounds_bytes


def pipeline(tmp_path, gfa: str, gaf: str, thr=20, geo=0, grid=2, shards=1, stats=None):
    gp = tmp_path / "g.gfa"
    gp.write_bytes(gfa.encode())
    try:
        graph = load_graph(str(gp))
    except PantasDataError:
        return ("raise", 0)
    data = gaf.encode()
    bounds = shard_bounds_bytes(data, shards)
    parts, worst = [], None
    for r in range(shards):
        lo, hi = bounds[r], bounds[r + 1]
        flat, code, off, ndef, why = run_fastsim(graph, data[lo:hi], thr, file_off=lo, geo=geo, grid=grid)
        if stats is not None:
            stats["deferred"] = stats.get("deferred", 0) + ndef
            stats["lines"] = stats.get("lines", 0) + int(flat.sums[-3])
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


@pytest.mark.parametrize("geo", [0, 1])
def test_golden(geo, tmp_path):
    for case in GOLDEN:
        thr = 20 if case["thr"] is None else case["thr"]
        res = pipeline(tmp_path, case["gfa"], case["gaf"], thr, geo=geo)
        if case["name"] in UNSUPPORTED_BY_DESIGN:
            assert res[0] == "unsupported", (case["name"], res)
        elif case["returncode"] != 0:
            assert res[0] == "raise", (case["name"], res)
        else:
            assert res[0] == "ok", (case["name"], res)
            assert res[1] == case["stdout"], case["name"]
            assert res[2] == case["rej"], case["name"]


@pytest.mark.parametrize("seed", range(7300, 7340))
def test_fuzz_vs_oracle(seed, tmp_path):
    gfa, gaf = fuzzgen.make_case(seed, n_nodes=8 + seed % 20, n_reads=60 + 40 * (seed % 3), weird=(seed % 2 == 0),
                                 crlf=(seed % 6 == 0), trailing_newline=(seed % 4 != 0))
    orc = run_oracle(gaf.encode(), gfa.encode())
    res = pipeline(tmp_path, gfa, gaf, geo=seed % 2, grid=1 + seed % 3, shards=1 + seed % 3)
    assert orc.rc == 0
    assert res[0] == "ok", res
    assert res[1] == orc.out
    assert res[2] == orc.rej


@pytest.mark.parametrize("seed", range(7400, 7430))
def test_fuzz_risky_vs_oracle(seed, tmp_path):
    gfa, gaf = fuzzgen.make_risky_case(seed)
    orc = run_oracle(gaf.encode(), gfa.encode())
    res = pipeline(tmp_path, gfa, gaf, geo=seed % 2)
    if orc.rc == 0:
        assert res[0] == "ok" and res[1] == orc.out and res[2] == orc.rej
    else:
        assert res[0] == "raise", (res, orc.err)


def test_synthetic_reads_stay_on_the_fast_path(tmp_path):
    """Generator output (bench input): bit-exact, and (almost) nothing is handed to the per-record kernel."""
    from pantas_b200.synth import SynthGraph

    sg = SynthGraph("tiny", seed=11)
    gp = tmp_path / "g.gfa"
    sg.write_gfa(str(gp))
    buf, n = sg.gaf(300, first_pair=0)
    gaf = bytes(buf)
    orc = run_oracle(gaf, gp.read_bytes())
    assert orc.rc == 0
    for geo in (1, 2):
        st = {}
        res = pipeline(tmp_path, gp.read_text(), gaf.decode(), geo=geo, grid=3, stats=st)
        assert res[0] == "ok" and res[1] == orc.out and res[2] == orc.rej
        assert st["lines"] == n
        assert st["deferred"] <= n // 50, st


def test_records_with_base_qualities_stay_on_the_fast_path(tmp_path, monkeypatch):
    """What vg mpmap writes for FASTQ reads: a bq:Z: tag (one quality character per base, ':' '<' '>' among them) between
    AS and cs.  A long token like that must not send the record to the per-record kernel."""
    from pantas_b200.synth import SynthGraph

    monkeypatch.setenv("PANTAS_SYNTH_BQ", "1")
    sg = SynthGraph("tiny", seed=12)
    gp = tmp_path / "g.gfa"
    sg.write_gfa(str(gp))
    buf, n = sg.gaf(300, first_pair=0)
    gaf = bytes(buf)
    assert gaf.count(b"\tbq:Z:") == n - gaf.count(b"\t*\t0\t0\t0")      # (unmapped mates carry no tags)
    orc = run_oracle(gaf, gp.read_bytes())
    assert orc.rc == 0
    for geo in (1, 2):
        st = {}
        res = pipeline(tmp_path, gp.read_text(), gaf.decode(), geo=geo, grid=3, stats=st)
        assert res[0] == "ok" and res[1] == orc.out and res[2] == orc.rej
        assert st["lines"] == n
        assert st["deferred"] <= n // 50, st
    # a lower-case letter in the quality string (never from a sequencer): the exact path decides, same bytes
    weird = gaf.replace(b"\tbq:Z:#", b"\tbq:Z:s", 5).replace(b"\tbq:Z:$", b"\tbq:Z:cs:Z::3", 2)
    orc2 = run_oracle(weird, gp.read_bytes())
    res = pipeline(tmp_path, gp.read_text(), weird.decode(), geo=2, grid=3)
    if orc2.rc == 0:
        assert res[0] == "ok" and res[1] == orc2.out and res[2] == orc2.rej
    else:
        assert res[0] in ("raise", "unsupported"), res


def test_dense_short_records_overflow_the_tile_lists(tmp_path):
    """More records per tile than the record list holds: the whole tile takes the per-record kernel."""
    gfa = "H\tVN:Z:1.1\nS\t1\tACGTACGTAC\nS\t2\tGGGGG\nL\t1\t+\t2\t+\t*\n"
    rec = ["q\t10\t0\t10\t+\t>1\t10\t0\t10\t10\t10\t60\tcs:Z::10\tdv:f:0\n",
           "r\t12\t0\t12\t+\t>1>2\t15\t0\t12\t12\t12\t60\tcs:Z::12\tdv:f:0\n",
           "s\t12\t0\t12\t+\t<2<1\t15\t3\t15\t12\t12\t7\tcs:Z::12\tdv:f:0\n"]
    gaf = "".join(rec[i % 3] for i in range(240))
    orc = run_oracle(gaf.encode(), gfa.encode())
    assert orc.rc == 0
    for geo in (0, 1):
        st = {}
        res = pipeline(tmp_path, gfa, gaf, geo=geo, stats=st)
        assert res[0] == "ok" and res[1] == orc.out and res[2] == orc.rej
        if geo == 1:                   # 4 KiB tiles of 53-byte records: 77 > the 64 record slots of a tile
            assert st["deferred"] > 0


@pytest.mark.parametrize("seed", range(7600, 7606))
def test_fuzz_with_quality_tags(seed, tmp_path):
    """Long inert tags anywhere among the tags (before, between and after cs / dv), CRLF line ends included."""
    gfa, gaf = fuzzgen.make_case(seed, n_nodes=25, n_reads=300, weird=(seed % 2 == 0), crlf=(seed % 3 == 0))
    gaf = fuzzgen.add_quality_tags(gaf, seed)
    orc = run_oracle(gaf.encode(), gfa.encode())
    assert orc.rc == 0
    st = {}
    res = pipeline(tmp_path, gfa, gaf, geo=1 + seed % 2, grid=1 + seed % 3, stats=st)
    assert res[0] == "ok" and res[1] == orc.out and res[2] == orc.rej


@pytest.mark.parametrize("seed", range(7500, 7506))
def test_fuzz_production_geometry(seed, tmp_path):
    gfa, gaf = fuzzgen.make_case(seed, n_nodes=30, n_reads=400, weird=(seed % 2 == 0), crlf=(seed % 3 == 0))
    orc = run_oracle(gaf.encode(), gfa.encode())
    res = pipeline(tmp_path, gfa, gaf, geo=2, grid=1 + seed % 3)
    assert orc.rc == 0
    assert res[0] == "ok", res
    assert res[1] == orc.out
    assert res[2] == orc.rej


@pytest.mark.parametrize("seed", range(7600, 7612))
def test_far_links_are_counted_one_tile_later(seed, tmp_path):
    """Links too far apart for a node's inline deltas go through the hash table; the fast path lists them per tile and
    drains the list during the next tile's records phase (and after the last tile): many small tiles, several CTAs."""
    gfa, gaf = fuzzgen.make_case(seed, n_nodes=24, n_reads=150, weird=False)
    gfa, gaf = fuzzgen.spread_ids(gfa, gaf, pivot=12, shift=40000)
    orc = run_oracle(gaf.encode(), gfa.encode())
    assert orc.rc == 0
    res = pipeline(tmp_path, gfa, gaf, geo=seed % 3, grid=1 + seed % 4)
    assert res[0] == "ok", res
    assert res[1] == orc.out
    assert res[2] == orc.rej


def test_long_paths_duplicates_and_multi_op_records(tmp_path):
    """Paths of 40+ steps (a multi-op record spans more than one 32-lane batch of `walk`), consecutive duplicate ids
    (REF:188 collapses them) at every alignment relative to the 32-byte mask words, mismatches / indels on short nodes."""
    import random

    rng = random.Random(5)
    n = 70
    gfa = ["H\tVN:Z:1.1"] + [f"S\t{i}\t{'ACGT'[i % 4] * 2}" for i in range(1, n + 1)]
    gfa += [f"L\t{i}\t+\t{i + 1}\t+\t*" for i in range(1, n)] + [f"L\t{i}\t+\t{i + 2}\t+\t*" for i in range(1, n - 1, 7)]
    recs = []
    for r in range(160):
        a = rng.randrange(1, 20)
        k = rng.randrange(2, 48) if r % 5 == 0 else rng.randrange(2, 24)
        ids = list(range(a, a + k))
        if r % 3 == 0:                                   # a consecutive duplicate somewhere
            j = rng.randrange(0, k)
            ids.insert(j, ids[j])
        rev = r % 2 == 1
        path = "".join(("<" if rev else ">") + str(i) for i in (reversed(ids) if rev else ids))
        nuniq = k
        plen = 2 * len(ids)
        qlen = 2 * nuniq
        kind = r % 4
        if kind == 0:
            cs = f":{qlen}"
        elif kind == 1:
            x = rng.randrange(1, qlen - 1)
            cs = f":{x}*ag:{qlen - x - 1}"
        elif kind == 2:
            x = rng.randrange(1, qlen - 2)
            cs = f":{x}-ac:{qlen - x - 2}"
        else:
            x = rng.randrange(1, qlen - 1)
            cs = f":{x}+t:{qlen - x - 1}"
        name = "r" * (200 + r % 37)                 # long names: few records per tile, so that most of them stay on the fast path
        recs.append(f"{name}\t{qlen}\t0\t{qlen}\t+\t{path}\t{plen}\t0\t{plen}\t{qlen}\t{qlen}\t60\tAS:i:1\tcs:Z:{cs}\tdv:f:0.01")
    gfa_s, gaf_s = "\n".join(gfa) + "\n", "\n".join(recs) + "\n"
    orc = run_oracle(gaf_s.encode(), gfa_s.encode())
    assert orc.rc == 0, orc.err
    for geo in (0, 1, 2):
        st = {}
        res = pipeline(tmp_path, gfa_s, gaf_s, geo=geo, grid=2, stats=st)
        assert res[0] == "ok", res
        assert res[1] == orc.out
        assert res[2] == orc.rej
        if geo:
            assert st["deferred"] < 90, st                 # the 54 records with a duplicate id + a few list overflows
