"""The drop-in script end to end on the GPU box: argv / stdout / stderr / exit status (pantas:132)."""
import os
import subprocess
import sys

import pytest

import fuzzgen
from conftest import GOLDEN
from oracle.oracle import run_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "scripts", "alignments_augmentation_from_gaf.py")


def run_cli(gaf_path, gfa_path, *extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, SCRIPT, str(gaf_path), str(gfa_path), *extra], capture_output=True, env=e,
                          timeout=600)


def test_cli_matches_reference_output(tmp_path):
    case = next(c for c in GOLDEN if c["name"] == "B1_mixed")
    (tmp_path / "g.gfa").write_bytes(case["gfa"].encode())
    (tmp_path / "a.gaf").write_bytes(case["gaf"].encode())
    p = run_cli(tmp_path / "a.gaf", tmp_path / "g.gfa")
    assert p.returncode == 0, p.stderr.decode()
    assert p.stdout == case["stdout"]
    err = p.stderr.decode().splitlines()
    assert err == ["Read GFA", "Augmentation by GAF alignments", "Rejected alignments: 2", "Annotating GFA"]
    p = run_cli(tmp_path / "a.gaf", tmp_path / "g.gfa", "61")
    assert p.returncode == 0 and b"Rejected alignments: 9" in p.stderr


def test_cli_malformed_input_exits_nonzero_with_empty_stdout(tmp_path):
    case = next(c for c in GOLDEN if c["name"] == "C4_mixed_orient")
    (tmp_path / "g.gfa").write_bytes(case["gfa"].encode())
    (tmp_path / "a.gaf").write_bytes(case["gaf"].encode())
    p = run_cli(tmp_path / "a.gaf", tmp_path / "g.gfa")
    assert p.returncode != 0
    assert p.stdout == b""


def test_cli_streams_a_file_larger_than_the_staging_buffer(tmp_path):
    gfa, gaf = fuzzgen.make_case(6001, n_nodes=60, n_reads=6000, weird=True)
    want = run_oracle(gaf.encode(), gfa.encode())
    (tmp_path / "g.gfa").write_bytes(gfa.encode())
    (tmp_path / "a.gaf").write_bytes(gaf.encode())
    p = run_cli(tmp_path / "a.gaf", tmp_path / "g.gfa", env={"PANTAS_STAGE_MB": "1"})   # 1 MiB chunks, ~1 MB file
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    assert p.stdout == want.out
    assert f"Rejected alignments: {want.rej}".encode() in p.stderr


def test_cli_two_gpus_byte_range_shards(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    gfa, gaf = fuzzgen.make_case(6002, n_nodes=80, n_reads=5000, weird=True)
    want = run_oracle(gaf.encode(), gfa.encode())
    (tmp_path / "g.gfa").write_bytes(gfa.encode())
    (tmp_path / "a.gaf").write_bytes(gaf.encode())
    p = run_cli(tmp_path / "a.gaf", tmp_path / "g.gfa", env={"PANTAS_GPUS": "2"})
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    assert p.stdout == want.out
    assert f"Rejected alignments: {want.rej}".encode() in p.stderr


def test_cli_gzip_and_stdin_streams(tmp_path):
    """Compressed / piped GAF: same bytes as for the plain file."""
    import gzip

    gfa, gaf = fuzzgen.make_case(6003, n_nodes=60, n_reads=6000, weird=True)
    want = run_oracle(gaf.encode(), gfa.encode())
    (tmp_path / "g.gfa").write_bytes(gfa.encode())
    with gzip.open(tmp_path / "a.gaf.gz", "wb") as f:
        f.write(gaf.encode())
    p = run_cli(tmp_path / "a.gaf.gz", tmp_path / "g.gfa", env={"PANTAS_STAGE_MB": "1"})
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    assert p.stdout == want.out
    e = dict(os.environ, PANTAS_STAGE_MB="1")
    p = subprocess.run([sys.executable, SCRIPT, "-", str(tmp_path / "g.gfa")], input=gaf.encode(), capture_output=True, env=e, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    assert p.stdout == want.out


def test_batch_of_samples_on_one_graph(tmp_path):
    """Several GAFs against one resident graph: each output equals the one-sample result."""
    gfa, gaf1 = fuzzgen.make_case(6004, n_nodes=50, n_reads=3000, weird=True)
    gaf2 = "".join(gaf1.splitlines(keepends=True)[::2])               # another sample over the same graph
    (tmp_path / "g.gfa").write_bytes(gfa.encode())
    (tmp_path / "s1.gaf").write_bytes(gaf1.encode())
    (tmp_path / "s2.gaf").write_bytes(gaf2.encode())
    p = subprocess.run([sys.executable, "-m", "pantas_b200.batch", str(tmp_path / "g.gfa"), str(tmp_path / "out"),
                        str(tmp_path / "s1.gaf"), str(tmp_path / "s2.gaf"), str(tmp_path / "s1.gaf")],
                       capture_output=True, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    for name, gaf in (("s1", gaf1), ("s2", gaf2)):
        want = run_oracle(gaf.encode(), gfa.encode())
        assert want.rc == 0
        assert (tmp_path / "out" / f"{name}.gfa").read_bytes() == want.out
