"""Generate tests/golden/cases.json.gz by RUNNING THE REFERENCE.

Run in the build container, where /root/reference is mounted:

    python tests/golden/make_golden.py

Every case records the inputs and what
/root/reference/scripts/alignments_augmentation_from_gaf.py printed for them
(stdout bytes, the "Rejected alignments" count, the exit status).  The file is
committed; the GPU box (no /root/reference) checks the oracle and the CUDA path
against it.  The hand-written cases are the known-answer vectors of SURVEY.md
Appendix B / C; the rest come from tests/fuzzgen.py with fixed seeds.
"""
from __future__ import annotations

import base64
import gzip
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fuzzgen  # noqa: E402
from oracle.oracle import run_reference  # noqa: E402

T = "\t"

G1 = "\n".join([
    "H\tVN:Z:1.1",
    "S\t1\tACGTACGTAC\tEX:Z:T1_R1.1",
    "S\t2\tGGGGG",
    "S\t3\tTTTTTTTT\tEX:Z:T1_R1.2",
    "S\t4\tCCCC",
    "L\t1\t+\t2\t+\t*",
    "L\t2\t+\t3\t+\t*",
    "L\t1\t+\t3\t+\t*\tJN:Z:T1_R1.1.2",
    "L\t3\t+\t4\t+\t*",
    "P\tT1_R1\t1+,3+\t*",
]) + "\n"

A1 = "\n".join([
    "r1\t15\t0\t15\t+\t>1>3\t18\t3\t11\t15\t15\t60\tAS:i:15\tcs:Z::15\tdv:f:0.000000",
    "r2\t15\t0\t15\t+\t<3<1\t18\t7\t15\t15\t15\t60\tAS:i:15\tcs:Z::15\tdv:f:0.000000",
    "r3\t10\t0\t10\t+\t>1>2>3\t23\t5\t16\t10\t11\t60\tAS:i:9\tcs:Z::5-g:5\tdv:f:0.050000",
    "r4\t10\t0\t10\t+\t>1>2>3\t23\t5\t20\t10\t15\t60\tAS:i:9\tcs:Z::5-ggggg:5\tdv:f:0.050000",
    "r5\t10\t0\t10\t+\t>2>4\t9\t0\t9\t9\t9\t60\tAS:i:9\tcs:Z::9\tdv:f:0.000000",
    "r6\t10\t0\t10\t+\t>1\t10\t0\t10\t10\t10\t5\tAS:i:9\tcs:Z::10\tdv:f:0.000000",
    "r7\t10\t0\t0\t+\t*\t0\t0\t0\t0\t0\t0",
    "r8\t10\t0\t10\t+\t>1\t10\t0\t10\t10\t10\t60\tAS:i:9\tcs:Z::10\tdv:f:0.200000",
    "r9\t10\t0\t10\t+\t>1>2\t15\t8\t15\t7\t7\t60\tAS:i:5\tcs:Z::2*ag:4\tdv:f:0.100000",
]) + "\n"

G2 = "\n".join([
    "S\t1\tAAAA", "S\t2\tCCCC", "S\t3\tGGGG", "S\t4\tTTTT",
    "L\t1\t+\t2\t+\t*", "L\t1\t+\t2\t+\t*", "L\t2\t-\t3\t-\t*", "", "L", "P\tx\t1+,2+\t*",
]) + "\n"

A2 = "\n".join([
    "a\t8\t0\t8\t+\t<4<2\t8\t0\t8\t8\t8\t60\tcs:Z::8\tdv:f:0",
    "b\t8\t0\t8\t+\t>1>3\t8\t0\t8\t8\t8\t60\tcs:Z::8\tdv:f:0",
    "c\t8\t0\t8\t+\t>2>4\t8\t0\t8\t8\t8\t60\tcs:Z::8\tdv:f:0",
    "d\t8\t0\t8\t+\t>1>2>3\t12\t0\t12\t12\t12\t60\tcs:Z::12\tdv:f:0",
    "e\t8\t0\t8\t+\t>3>2\t8\t0\t8\t8\t8\t60\tcs:Z::8\tdv:f:0",
]) + "\n"


def one(path, plen, pstart, pend, tags, mapq=60):
    return f"q\t10\t0\t10\t+\t{path}\t{plen}\t{pstart}\t{pend}\t10\t10\t{mapq}" + "".join(T + t for t in tags) + "\n"


HAND = [
    ("B1_mixed", G1, A1, None),
    ("B1_thr0", G1, A1, 0),
    ("B1_thr61", G1, A1, 61),
    ("B2_order_dups", G2, A2, None),
    ("B3_ins_mid_node", G1, one(">1>2>3", 23, 8, 17, ["cs:Z::3+tt:6", "dv:f:0.1"]), None),
    ("B3_lone_star", G1, one(">1>2>3", 23, 9, 16, ["cs:Z:*ag:5:1", "dv:f:0.1"]), None),
    ("B3_leading_star", G1, one(">1>2", 15, 5, 15, ["cs:Z:*ag:9", "dv:f:0.1"]), None),
    ("B3_clip_front", G1, one(">1>2", 15, 0, 7, ["cs:Z:+acg:7", "dv:f:0"]), None),
    ("B3_dup_collapse", G1, one(">1>1>2", 25, 0, 15, ["cs:Z::15", "dv:f:0"]), None),
    ("B3_no_cs_single", G1, one(">1", 10, 0, 10, ["dv:f:0"]), None),
    ("C1_no_dv", G1, one(">1", 10, 0, 10, ["cs:Z::10"]), None),
    ("C2_empty_line", G1, A1 + "\n", None),
    ("C2_short_line", G1, "x\t1\t2\n", None),
    ("C3_no_cs_multi", G1, one(">1>2", 15, 0, 15, ["dv:f:0"]), None),
    ("C4_mixed_orient", G1, one(">1<2", 15, 0, 15, ["cs:Z::15", "dv:f:0"]), None),
    ("C5_unknown_node", G1, one(">1>77", 15, 0, 15, ["cs:Z::15", "dv:f:0"]), None),
    ("C5_leading_zero_id", G1, one(">01", 10, 0, 10, ["cs:Z::10", "dv:f:0"]), None),
    ("C6_dv_exponent", G1, one(">1", 10, 0, 10, ["cs:Z::10", "dv:f:1e-05"]), None),
    ("C7_bad_mapq", G1, one(">1", 10, 0, 10, ["cs:Z::10", "dv:f:0"], mapq="6x"), None),
    ("C7_bad_coord", G1, one(">1", "1o", 0, 10, ["cs:Z::10", "dv:f:0"]), None),
    ("rejected_garbage_ok", G1, one(">1<<", "zz", "y", "w", ["nothing"], mapq=3), None),
    ("unmapped_garbage_ok", G1, one("*", "zz", "y", "w", ["nothing"]), None),
    ("signed_ints", G1, one(">1>2", "+15", "+0", "15", ["cs:Z::15", "dv:f:0"], mapq="+60"), None),
    ("neg_mapq", G1, one(">1>2", 15, 0, 15, ["cs:Z::15", "dv:f:0"], mapq="-1"), None),
    ("neg_start", G1, one(">1>2", 15, -2, 15, ["cs:Z::20", "dv:f:0"]), None),
    ("underscore_int", G1, one(">1>2", "1_5", 0, 15, ["cs:Z::15", "dv:f:0"]), None),
    ("path_prefix_junk_rev", G1, one("junk<3<1", 18, 7, 15, ["cs:Z::15", "dv:f:0"]), None),
    ("path_no_sep", G1, one("123", 18, 7, 15, ["cs:Z::15", "dv:f:0"]), None),
    ("path_empty_piece", G1, one(">>1", 10, 0, 10, ["cs:Z::10", "dv:f:0"]), None),
    ("path_trailing_sep", G1, one(">1>", 10, 0, 10, ["cs:Z::10", "dv:f:0"]), None),
    ("cs_inside_other_tag", G1, one(">1", 10, 0, 10, ["xx:Z:abcs:Z::4", "cs:Z::10", "dv:f:0"]), None),
    ("cs_not_Z", G1, one(">1>2", 15, 8, 14, ["cs:i:5", "dv:f:0"]), None),
    ("cs_double_prefix", G1, one(">1>2", 15, 0, 15, ["cs:Z::5cs:Z::10", "dv:f:0"]), None),
    ("cs_zero_len_ops", G1, one(">1>2>3", 23, 0, 23, ["cs:Z::10-:0:5+:8", "dv:f:0"]), None),
    ("cs_trailing_zero_op", G1, one(">1>2", 15, 0, 14, ["cs:Z::10-", "dv:f:0"]), None),
    ("cs_eq_ops", G1, one(">1>2", 15, 0, 15, ["cs:Z:=ACGTACGTAC=GGGGG", "dv:f:0"]), None),
    ("cs_colon_text", G1, one(">1>2", 15, 7, 15, ["cs:Z::abc:5", "dv:f:0"]), None),
    ("cs_two_leading_stars", G1, one(">1>2", 15, 0, 15, ["cs:Z:*ag*ct:8:5", "dv:f:0"]), None),
    ("cs_del_at_node_start", G1, one(">1>2>3", 23, 0, 23, ["cs:Z::10-gg:3:8", "dv:f:0"]), None),
    ("cs_del_at_node_end", G1, one(">1>2>3", 23, 0, 23, ["cs:Z::12-ggg:8", "dv:f:0"]), None),
    ("cs_del_at_node_end_rev", G1, one("<1<2<3", 23, 0, 23, ["cs:Z::12-ggg:8", "dv:f:0"]), None),
    ("cs_del_at_node_start_rev", G1, one("<1<2<3", 23, 0, 23, ["cs:Z::10-gg:3:8", "dv:f:0"]), None),
    ("cs_whole_node_ins", G1, one(">1>2>3", 23, 0, 23, ["cs:Z::10+ggggg:8", "dv:f:0"]), None),
    ("cs_tilde", G1, one(">1>2", 15, 0, 15, ["cs:Z::5~gt12ag:5", "dv:f:0"]), None),
    ("dv_second_match", G1, one(">1", 10, 0, 10, ["cs:Z::10", "dv:f:x", "zdv:f:0.5"]), None),
    ("dv_trailing_dot", G1, one(">1", 10, 0, 10, ["cs:Z::10", "dv:f:0."]), None),
    ("dv_half_ulp_below", G1, one(">1", 10, 0, 10, ["cs:Z::10", "dv:f:0.100000000000000012490009027033011079765856266021728515625"]), None),
    ("dv_half_ulp_above", G1, one(">1", 10, 0, 10, ["cs:Z::10", "dv:f:0.1000000000000000124900090270330110797658562660217285156250000001"]), None),
    ("spaces_as_separators", G1, one(">1>2", 15, 0, 15, ["cs:Z::15", "dv:f:0"]).replace("\t", "  "), None),
    ("leading_ws_line", G1, "  \t" + one(">1>2", 15, 0, 15, ["cs:Z::15", "dv:f:0"]), None),
    ("no_trailing_newline", G1, one(">1>2", 15, 0, 15, ["cs:Z::15", "dv:f:0"]).rstrip("\n"), None),
    ("empty_gaf", G1, "", None),
    ("crlf_gaf", G1, A1.replace("\n", "\r\n"), None),
    ("bare_cr_gaf", G1, A1.replace("\n", "\r"), None),
    ("revisit_node", G1, one(">1>2>1>2", 30, 0, 30, ["cs:Z::30", "dv:f:0"]), None),
    ("self_loop_dedupe", G1, one(">2>2>2", 15, 1, 14, ["cs:Z::13", "dv:f:0"]), None),
]


def run_case(d, name, gfa, gaf, thr):
    gp, ap = os.path.join(d, "g.gfa"), os.path.join(d, "a.gaf")
    with open(gp, "w", newline="") as f:
        f.write(gfa)
    with open(ap, "w", newline="") as f:
        f.write(gaf)
    ref = run_reference(ap, gp, thr)
    return {
        "name": name, "gfa": gfa, "gaf": gaf, "thr": thr,
        "returncode": 0 if ref.returncode == 0 else 1,
        "stdout_b64": base64.b64encode(ref.stdout).decode("ascii"),
        "rej": ref.rej,
    }


def main():
    cases = []
    with tempfile.TemporaryDirectory() as d:
        for name, gfa, gaf, thr in HAND:
            cases.append(run_case(d, name, gfa, gaf, thr))
        for s in range(120):
            gfa, gaf = fuzzgen.make_case(1000 + s, n_nodes=10 + s % 9, n_reads=30, weird=(s % 3 == 0),
                                         crlf=(s % 7 == 0), trailing_newline=(s % 5 != 0))
            cases.append(run_case(d, f"fuzz_{1000 + s}", gfa, gaf, None if s % 4 else 20))
        for s in range(80):
            gfa, gaf = fuzzgen.make_risky_case(2000 + s)
            cases.append(run_case(d, f"risky_{2000 + s}", gfa, gaf, None))
        # records with a bq:Z: quality tag (what vg mpmap writes for FASTQ reads) somewhere among the tags
        for s in range(24):
            if s < 18:
                gfa, gaf = fuzzgen.make_case(3000 + s, n_nodes=10 + s % 9, n_reads=30, weird=(s % 3 == 0),
                                             crlf=(s % 7 == 0), trailing_newline=(s % 5 != 0))
            else:
                gfa, gaf = fuzzgen.make_risky_case(3000 + s)
            gaf = fuzzgen.add_quality_tags(gaf, 3000 + s, share=0.8)
            cases.append(run_case(d, f"quals_{3000 + s}", gfa, gaf, None))
    out = os.path.join(HERE, "cases.json.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        f.write(json.dumps(cases, indent=0).encode("utf-8"))
    ok = sum(1 for c in cases if c["returncode"] == 0)
    print(f"wrote {out}: {len(cases)} cases, {ok} succeed, {len(cases) - ok} raise")
    for c in cases[: len(HAND)]:
        print(f"  {c['name']:28s} rc={c['returncode']} rej={c['rej']}")


if __name__ == "__main__":
    main()
