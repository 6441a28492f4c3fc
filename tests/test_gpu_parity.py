"""Parity of the CUDA path (through the C ABI) with the reference / oracle.  B200 only."""
import io
import os

import numpy as np
import pytest

import fuzzgen
from conftest import GOLDEN, UNSUPPORTED_BY_DESIGN
from oracle.oracle import run_oracle

pytestmark = pytest.mark.gpu


def _engine(**env):
    from pantas_b200.engine import AugmentEngine

    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return AugmentEngine(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def gpu_pipeline(tmp_path, gfa: str, gaf: str, thr=20, eng=None, via_host=False, chunks=1):
    """-> ('ok', stdout, rej) | ('raise', code) | ('unsupported', code)"""
    import torch

    from pantas_b200.counts import Counts
    from pantas_b200.errors import PantasDataError, UnsupportedInput
    from pantas_b200.gfa import load_graph, write_augmented
    from pantas_b200.shard import shard_bounds_bytes

    gp = tmp_path / "g.gfa"
    gp.write_bytes(gfa.encode())
    try:
        graph = load_graph(str(gp))
    except PantasDataError:
        return ("raise", 0)
    eng = eng or _engine()
    eng.set_graph(graph)
    data = np.frombuffer(gaf.encode(), dtype=np.uint8)
    bounds = shard_bounds_bytes(data, chunks) if data.size else [0, 0]
    keep = []
    for lo, hi in zip(bounds, bounds[1:]):
        n = hi - lo
        if via_host:
            h = torch.zeros(n + 16, dtype=torch.uint8).pin_memory()
            h[:n] = torch.from_numpy(data[lo:hi].copy())
            keep.append(h)
            t = eng.process_host(h.data_ptr(), n, lo, thr)
            eng.wait_copy(t)
        else:
            d = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
            d[:n] = torch.from_numpy(data[lo:hi].copy()).cuda()
            keep.append(d)
            eng.process_device(d, n, lo, thr)
    try:
        eng.check_data_error()
    except PantasDataError as e:
        return ("raise", e.code)
    except UnsupportedInput as e:
        return ("unsupported", e.code)
    counts = Counts.from_flat(eng.export())
    out = io.StringIO()
    try:
        write_augmented(str(gp), graph, counts, out)
    except PantasDataError:
        return ("raise", 0)
    return ("ok", out.getvalue().encode(), counts.rej)


def gpu_pipeline_device_gfa(gfa: str, gaf: str, thr=20, eng=None):
    """The same with BOTH GFA passes on the device (pantas_b200/gfa_device.py): what the CLI runs."""
    import torch

    from pantas_b200.errors import PantasDataError, UnsupportedInput
    from pantas_b200.gfa_device import DeviceGfa

    eng = eng or _engine()
    try:
        dg = DeviceGfa.load_bytes(eng, gfa.encode())
    except PantasDataError:
        return ("raise", 0)
    except UnsupportedInput as e:
        return ("unsupported", e.code)
    dg.set_graph()
    data = np.frombuffer(gaf.encode(), dtype=np.uint8)
    n = int(data.size)
    d = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
    if n:
        d[:n] = torch.from_numpy(data.copy()).cuda()
    eng.process_device(d, n, 0, thr)
    try:
        eng.check_data_error()
    except PantasDataError as e:
        return ("raise", e.code)
    except UnsupportedInput as e:
        return ("unsupported", e.code)
    sums, stamps, novel, sparse = eng.export_device()
    g = dg.graph
    rej = int(sums[3 * g.n_nodes + g.n_edges].item())
    try:
        host = dg.render(sums, stamps, novel.cpu().numpy().view(np.uint64).reshape(-1, 3),
                         sparse.cpu().numpy().view(np.uint64).reshape(-1, 3))
    except PantasDataError:
        return ("raise", 0)
    return ("ok", host.numpy().tobytes(), rej)


def check_golden(case, res):
    if case["name"] in UNSUPPORTED_BY_DESIGN:
        assert res[0] == "unsupported", (case["name"], res)
    elif case["returncode"] != 0:
        assert res[0] == "raise", (case["name"], res)
    else:
        assert res[0] == "ok", (case["name"], res)
        assert res[1] == case["stdout"], case["name"]
        assert res[2] == case["rej"], case["name"]


def test_library_is_native():
    from pantas_b200 import _lib

    lib = _lib.load_library()
    assert lib.pt_abi_version() == 1


def test_golden_cases_device_buffers(tmp_path):
    eng = _engine()
    for case in GOLDEN:
        thr = 20 if case["thr"] is None else case["thr"]
        check_golden(case, gpu_pipeline(tmp_path, case["gfa"], case["gaf"], thr, eng=eng))


def test_golden_cases_device_gfa_passes():
    """GFA pass 1 and the writer on the device (REF:121-126, 377-427): same bytes, same failures."""
    eng = _engine()
    for case in GOLDEN:
        thr = 20 if case["thr"] is None else case["thr"]
        check_golden(case, gpu_pipeline_device_gfa(case["gfa"], case["gaf"], thr, eng=eng))


def test_device_gfa_passes_long_path_lines():
    """The P line of a reference path is one token of up to 100 MB: pass 1 must not tokenise it, the writer copies it with
    a whole block.  Lengths around the 4096-byte switch, leading / trailing blanks (the reference echoes the stripped line)."""
    gfa, gaf = fuzzgen.make_case(9300, n_nodes=40, n_reads=300, weird=False)
    lines = gfa.splitlines(keepends=True)
    extra = []
    for k, n in enumerate((4090, 4096, 4097, 5000, 70001, 300000)):
        body = ",".join(f"{1 + (i % 40)}+" for i in range(n // 3))[:n - 10]
        extra.append(f"{'  ' if k % 2 else ''}P\tref{k}\t{body}\t*{'  ' if k % 3 == 0 else ''}\n")
    extra.append("#" + "x" * 9000 + " \t \n")
    gfa2 = "".join(lines[:3] + extra[:3] + lines[3:] + extra[3:])
    orc = run_oracle(gaf.encode(), gfa2.encode())
    assert orc.rc == 0
    res = gpu_pipeline_device_gfa(gfa2, gaf)
    assert res[0] == "ok", res
    assert res[1] == orc.out
    assert res[2] == orc.rej


@pytest.mark.parametrize("seed", range(9200, 9212))
def test_fuzz_device_gfa_passes_vs_oracle(seed):
    gfa, gaf = fuzzgen.make_case(seed, n_nodes=10 + seed % 40, n_reads=300, weird=(seed % 2 == 0), crlf=(seed % 6 == 0),
                                 trailing_newline=(seed % 4 != 0))
    orc = run_oracle(gaf.encode(), gfa.encode())
    assert orc.rc == 0
    res = gpu_pipeline_device_gfa(gfa, gaf, eng=_engine(PANTAS_TEAM_TILE=[8192, 1024, 8193][seed % 3]))
    assert res[0] == "ok", res
    assert res[1] == orc.out
    assert res[2] == orc.rej


@pytest.mark.parametrize("preset,pairs,seed", [("dm-full", 1_000_000, 1002), ("hs-chr1", 500_000, 1003), ("gene-panel", 1_000_000, 1005)])
def test_bench_scale_graphs_byte_identical_to_the_oracle(preset, pairs, seed, tmp_path):
    """The graphs the bench and the scaling runs use, at 1-2 M alignments: the augmented GFA (every NC / IL / OL / RC tag
    of millions of S and L lines, novel links in first-seen order) is byte-identical to the CPU oracle's."""
    import torch

    from pantas_b200.gfa_device import DeviceGfa
    from pantas_b200.synth import SynthGraph

    sg = SynthGraph(preset, seed=seed)
    gp = tmp_path / "g.gfa"
    sg.write_gfa(str(gp))
    gaf, n_lines = sg.gaf(pairs, first_pair=0)
    gfa_bytes = gp.read_bytes()
    want = run_oracle(gaf, gfa_bytes)
    assert want.rc == 0, want.err
    eng = _engine()
    dg = DeviceGfa.load(eng, str(gp))
    dg.set_graph()
    n = int(gaf.shape[0])
    d = torch.zeros(n + 32, dtype=torch.uint8, device="cuda")
    d[:n] = torch.from_numpy(gaf).cuda()
    eng.process_device(d, n, 0, 20)
    eng.check_data_error()
    sums, stamps, novel, sparse = eng.export_device()
    g = dg.graph
    assert int(sums[3 * g.n_nodes + g.n_edges].item()) == want.rej
    assert int(sums[3 * g.n_nodes + g.n_edges + 1].item()) == n_lines == want.n_lines
    host = dg.render(sums, stamps, novel.cpu().numpy().view(np.uint64).reshape(-1, 3), sparse.cpu().numpy().view(np.uint64).reshape(-1, 3))
    got = host.numpy()
    ref = np.frombuffer(want.out, dtype=np.uint8)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)
    assert eng.stats()["deferred_lines"] < 0.02 * n_lines


def test_golden_cases_host_buffers(tmp_path):
    eng = _engine()
    for case in GOLDEN:
        thr = 20 if case["thr"] is None else case["thr"]
        check_golden(case, gpu_pipeline(tmp_path, case["gfa"], case["gaf"], thr, eng=eng, via_host=True))


GEOS = [dict(PANTAS_TEAM_TILE=8192), dict(PANTAS_TEAM_TILE=1024), dict(PANTAS_TEAM_TILE=8193),
        dict(PANTAS_TEAM_TILE=8192, PANTAS_TILE_BYTES=6144), dict(PANTAS_TEAM_TILE=8192, PANTAS_TILE_BYTES=2000),
        dict(PANTAS_TEAM_TILE=8193, PANTAS_TILE_BYTES=7168), dict(PANTAS_TEAM_TILE=8192, PANTAS_TILE_BYTES=9216)]


@pytest.mark.parametrize("geo", range(len(GEOS)))
def test_golden_team_geometries(geo, tmp_path):
    """The fast path gives the same bytes for every layout and tile length (1 KiB tiles: many tile boundaries, records
    longer than the look-ahead, full lists -- all of which hand records to the exact per-record kernel)."""
    eng = _engine(**GEOS[geo])
    for case in GOLDEN:
        thr = 20 if case["thr"] is None else case["thr"]
        check_golden(case, gpu_pipeline(tmp_path, case["gfa"], case["gaf"], thr, eng=eng))


@pytest.mark.parametrize("seed", range(9000, 9030))
def test_fuzz_vs_oracle(seed, tmp_path):
    gfa, gaf = fuzzgen.make_case(seed, n_nodes=10 + seed % 40, n_reads=400, weird=(seed % 2 == 0),
                                 crlf=(seed % 6 == 0), trailing_newline=(seed % 4 != 0))
    if seed % 3 == 1:
        gaf = fuzzgen.add_quality_tags(gaf, seed)      # bq:Z: tags of FASTQ reads: long inert tokens among the tags
    orc = run_oracle(gaf.encode(), gfa.encode())
    assert orc.rc == 0
    eng = _engine(**GEOS[seed % len(GEOS)])
    res = gpu_pipeline(tmp_path, gfa, gaf, eng=eng, chunks=1 + seed % 3, via_host=(seed % 5 == 0))
    assert res[0] == "ok", res
    assert res[1] == orc.out
    assert res[2] == orc.rej


@pytest.mark.parametrize("preset,pairs,geo", [("tiny", 20000, 0), ("dm-chr4", 100000, 0), ("dm-chr4", 50000, 5),
                                              ("tiny", 20000, 1), ("gene-panel", 50000, 6), ("tiny", 20000, 3)])
def test_synthetic_workload_matches_oracle(preset, pairs, geo, tmp_path):
    """The bench workload's generator (vg-mpmap-shaped records, mostly fast-path) at a size the CPU oracle
    finishes in seconds: the augmented GFA is byte-identical."""
    import torch

    from pantas_b200.counts import Counts
    from pantas_b200.gfa import load_graph, write_augmented
    from pantas_b200.synth import SynthGraph

    sg = SynthGraph(preset, seed=77)
    gp = tmp_path / "g.gfa"
    sg.write_gfa(str(gp))
    gaf, n_lines = sg.gaf(pairs, first_pair=0, threads=4)
    want = run_oracle(gaf, gp.read_bytes())
    assert want.rc == 0, want.err
    graph = load_graph(str(gp))
    eng = _engine(**GEOS[geo])
    eng.set_graph(graph)
    n = int(gaf.shape[0])
    d = torch.zeros(n + 32, dtype=torch.uint8, device="cuda")
    d[:n] = torch.from_numpy(gaf).cuda()
    eng.process_device(d, n, 0, 20)
    eng.check_data_error()
    counts = Counts.from_flat(eng.export())
    assert counts.n_lines == n_lines == want.n_lines
    assert counts.rej == want.rej
    out = io.StringIO()
    write_augmented(str(gp), graph, counts, out)
    assert out.getvalue().encode() == want.out
    st = eng.stats()
    if GEOS[geo]["PANTAS_TEAM_TILE"] >= 4096:
        assert st["deferred_lines"] < 0.02 * n_lines, st      # the fast path really is the path taken


def test_rerun_after_reset_is_identical(tmp_path):
    gfa, gaf = fuzzgen.make_case(4242, n_nodes=30, n_reads=500)
    eng = _engine()
    a = gpu_pipeline(tmp_path, gfa, gaf, eng=eng)
    eng.reset()
    b = gpu_pipeline(tmp_path, gfa, gaf, eng=eng)
    assert a == b and a[0] == "ok"


def test_no_graph_is_an_error():
    from pantas_b200.errors import NativeLibraryError

    eng = _engine()
    with pytest.raises(NativeLibraryError):
        eng.reset()


@pytest.mark.parametrize("preset,pairs", [("dm-full", 1_000_000), ("hs-chr1", 250_000), ("gene-panel", 500_000)])
def test_full_size_graphs_are_linear_in_the_gaf(preset, pairs):
    """Size-independent property at the bench's graph sizes (the CPU oracle would take minutes there):
    counting a GAF in one piece == counting its halves separately and adding up (sums add, first-touch
    stamps take the minimum, side tables merge), and every record is accounted for."""
    import torch

    from pantas_b200.counts import merge_flat
    from pantas_b200.shard import shard_bounds_bytes
    from pantas_b200.synth import SynthGraph

    sg = SynthGraph(preset, seed=1003)
    graph = sg.graph()
    gaf, n_lines = sg.gaf(pairs, first_pair=0)
    n = int(gaf.shape[0])
    d = torch.zeros(n + 32, dtype=torch.uint8, device="cuda")
    d[:n] = torch.from_numpy(gaf).cuda()

    eng = _engine()
    eng.set_graph(graph)
    eng.process_device(d, n, 0, 20)
    eng.check_data_error()
    whole = eng.export()
    assert int(whole.sums[-3]) == n_lines
    assert eng.stats()["deferred_lines"] < 0.05 * n_lines

    bounds = shard_bounds_bytes(gaf, 3)
    parts = []
    for k, (lo, hi) in enumerate(zip(bounds, bounds[1:])):
        e2 = _engine(**GEOS[[0, 5, 3][k]])
        e2.set_graph(graph)
        piece = torch.zeros(hi - lo + 32, dtype=torch.uint8, device="cuda")
        piece[: hi - lo] = d[lo:hi]
        e2.process_device(piece, hi - lo, lo, 20)
        e2.check_data_error()
        parts.append(e2.export())
        e2.close()
    merged = merge_flat(parts)
    assert np.array_equal(merged.sums, whole.sums)
    assert np.array_equal(merged.stamps, whole.stamps)
    for a, b in ((merged.novel, whole.novel), (merged.sparse, whole.sparse)):
        assert sorted(map(tuple, a.tolist())) == sorted(map(tuple, b.tolist()))


def test_pieces_out_of_order_in_one_context_fold_epochs():
    """Three pieces of one GAF fed to ONE context, last piece first: a chunk that starts before the end of what the
    open epoch has seen folds the epoch, so the export has to combine folded 64-bit totals with the open epoch's 32-bit
    state -- and it must not change anything (a second export gives the same arrays)."""
    import torch
    from pantas_b200.shard import shard_bounds_bytes
    from pantas_b200.synth import SynthGraph

    sg = SynthGraph("dm-chr4", seed=1001)
    graph = sg.graph()
    gaf, n_lines = sg.gaf(100_000, first_pair=0)
    n = int(gaf.shape[0])
    d = torch.zeros(n + 32, dtype=torch.uint8, device="cuda")
    d[:n] = torch.from_numpy(gaf).cuda()
    eng = _engine()
    eng.set_graph(graph)
    eng.process_device(d, n, 0, 20)
    eng.check_data_error()
    whole = eng.export()

    eng.reset()
    bounds = shard_bounds_bytes(gaf, 3)
    for lo, hi in reversed(list(zip(bounds, bounds[1:]))):
        piece = torch.zeros(hi - lo + 32, dtype=torch.uint8, device="cuda")
        piece[: hi - lo] = d[lo:hi]
        eng.process_device(piece, hi - lo, lo, 20)
        eng.sync()
    eng.check_data_error()
    for _ in range(2):
        got = eng.export()
        assert np.array_equal(got.sums, whole.sums)
        assert np.array_equal(got.stamps, whole.stamps)
        for a, b in ((got.novel, whole.novel), (got.sparse, whole.sparse)):
            assert sorted(map(tuple, a.tolist())) == sorted(map(tuple, b.tolist()))
    eng.close()


@pytest.mark.parametrize("seed", range(9100, 9108))
def test_far_links_vs_oracle(seed, tmp_path):
    """Links farther apart than an inline delta: hash-table probes, listed per tile and drained one tile later."""
    gfa, gaf = fuzzgen.make_case(seed, n_nodes=30, n_reads=3000, weird=False)
    gfa, gaf = fuzzgen.spread_ids(gfa, gaf, pivot=15, shift=50000)
    orc = run_oracle(gaf.encode(), gfa.encode())
    assert orc.rc == 0
    eng = _engine(**GEOS[seed % 4])
    res = gpu_pipeline(tmp_path, gfa, gaf, eng=eng, chunks=1 + seed % 2)
    assert res[0] == "ok", res
    assert res[1] == orc.out
    assert res[2] == orc.rej
