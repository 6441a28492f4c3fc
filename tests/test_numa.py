"""Host-side NUMA placement helper (pantas_b200/numa.py): parsing and the do-nothing paths (no GPU needed)."""
import os

from pantas_b200 import numa


def test_parse_cpulist():
    assert numa.parse_cpulist("0-3,8,10-11") == {0, 1, 2, 3, 8, 10, 11}
    assert numa.parse_cpulist("") == set()
    assert numa.parse_cpulist("5") == {5}
    assert numa.parse_cpulist("0-1, 4-5\n") == {0, 1, 4, 5}


def test_single_node_box_is_left_alone(tmp_path):
    (tmp_path / "devices" / "system" / "node").mkdir(parents=True)
    (tmp_path / "devices" / "system" / "node" / "online").write_text("0\n")
    before = os.sched_getaffinity(0)
    info = numa.bind_to_gpu_node(0, sysfs=str(tmp_path))
    assert info["bound"] is False and info["why"] == "one NUMA node"
    assert os.sched_getaffinity(0) == before


def test_unknown_gpu_node_is_left_alone(tmp_path):
    (tmp_path / "devices" / "system" / "node").mkdir(parents=True)
    (tmp_path / "devices" / "system" / "node" / "online").write_text("0-1\n")
    before = os.sched_getaffinity(0)
    info = numa.bind_to_gpu_node(0, sysfs=str(tmp_path))      # no CUDA device here / no numa_node file in the fake tree
    assert info["bound"] is False
    assert os.sched_getaffinity(0) == before


def test_missing_sysfs_is_left_alone(tmp_path):
    info = numa.bind_to_gpu_node(0, sysfs=str(tmp_path / "nothing"))
    assert info["bound"] is False


def test_binds_to_the_cpus_of_the_node(tmp_path):
    """Two nodes in a fake sysfs tree; node 1 owns one of the CPUs this process may run on."""
    before = os.sched_getaffinity(0)
    mine = sorted(before)
    nodes = tmp_path / "devices" / "system" / "node"
    (nodes / "node0").mkdir(parents=True)
    (nodes / "node1").mkdir(parents=True)
    (nodes / "online").write_text("0-1\n")
    (nodes / "node0" / "cpulist").write_text("100000-100003\n")               # CPUs this process does not have
    (nodes / "node1" / "cpulist").write_text(f"{mine[0]},100004-100007\n")
    try:
        info = numa.bind_to_gpu_node(0, sysfs=str(tmp_path), node=1)
        assert info["bound"] is True and info["node"] == 1 and info["cpus"] == 1
        assert os.sched_getaffinity(0) == {mine[0]}
        info = numa.bind_to_gpu_node(0, sysfs=str(tmp_path), node=0)             # nothing allowed there: left alone
        assert info["bound"] is False
        assert os.sched_getaffinity(0) == {mine[0]}
    finally:
        os.sched_setaffinity(0, before)
