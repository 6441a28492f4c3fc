"""pantas_b200/tags.py against the consumer's own parser: call.py's build_attrs applied to the text the oracle prints
(needs /root/reference; skipped where it is not mounted) -- SURVEY.md section 8f row 3."""
import importlib.util
import os

import numpy as np
import pytest

import fuzzgen
from hostsim_util import run_hostsim
from oracle.oracle import run_oracle
from pantas_b200.counts import Counts
from pantas_b200.gfa import load_graph
from pantas_b200.tags import cluster, link_attrs, node_attrs

CALL = "/root/reference/scripts/call.py"


def _call_module():
    spec = importlib.util.spec_from_file_location("ref_call", CALL)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.skipif(not os.path.exists(CALL), reason="reference not mounted")
@pytest.mark.parametrize("seed", range(8400, 8412))
def test_node_and_link_attrs_match_build_attrs(seed, tmp_path):
    call = _call_module()
    gfa, gaf = fuzzgen.make_case(seed, n_nodes=20 + seed % 15, n_reads=500, weird=False)
    want = run_oracle(gaf.encode(), gfa.encode())
    assert want.rc == 0
    gp = tmp_path / "g.gfa"
    gp.write_bytes(gfa.encode())
    graph = load_graph(str(gp))
    flat, code, _, _ = run_hostsim(graph, gaf.encode(), 20)
    assert code == 0
    counts = Counts.from_flat(flat)
    na = node_attrs(graph, counts)
    la = link_attrs(graph, counts)
    seen_multi = 0
    novel = []
    key_to_edge = {int(k): e for e, k in enumerate(graph.edge_keys)}
    for line in want.out.decode().splitlines():
        if line.startswith("S"):
            _, nid, seq, *fields = line.split()
            a = call.build_attrs(fields)
            i = int(nid) - graph.min_id
            assert a["NC"] == na.nc[i]
            for tag, d, mx in (("IL", na.il, na.max_il), ("OL", na.ol, na.max_ol)):
                if tag in a:
                    assert d[i] == a[tag], (nid, tag)
                    assert mx[i] == a["MAX" + tag]
                    seen_multi += len(a[tag]) > 1
                else:
                    assert i not in d and mx[i] == -1
        elif line.startswith("L"):
            _, f, _, t, _, _, *fields = line.split()
            a = call.build_attrs(fields)
            if "ID" in a:
                novel.append((int(f) - graph.min_id, int(t) - graph.min_id, a["RC"]))
            else:
                e = key_to_edge.get(((int(f) - graph.min_id) << 32) | (int(t) - graph.min_id))
                if e is not None and a["RC"]:
                    assert la.rc[e] == a["RC"]
    assert [tuple(r) for r in la.novel.tolist()] == novel


def test_cluster_is_call_py_clustering():
    assert cluster([[0, 5]]) == [[0, 5]]
    assert cluster([[0, 5], [1, 3]]) == [[0, 8]]                      # closer than d = 3: one cluster, floor of the weighted mean
    assert cluster([[0, 5], [10, 3], [9, 1], [2, 2]]) == [[0, 7], [9, 4]]
