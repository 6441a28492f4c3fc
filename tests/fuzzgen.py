"""Seeded random GFA + GAF generator for differential tests.

Produces small graphs in the build/annotate.cpp GFA dialect and GAF records
that exercise every branch of the reference's augment loop
(/root/reference/scripts/alignments_augmentation_from_gaf.py:142-363): both
orientations, start/end offsets, all cs op kinds, leading '*', whole-node
deletions/insertions (node dropping -> novel links), 2-op clipping, duplicate
and revisited nodes, MAPQ / '*' / dv filters with values around 0.1, missing cs
tag, shuffled tag order, CRLF endings, no trailing newline.

``safe=True`` records never make the reference raise (cs always covers the
walk); ``safe=False`` may (used one record per case, expecting a crash or not).
"""
from __future__ import annotations

import random

DV_CHOICES = [
    "0", "0.0", "0.000000", "0.050000", "0.1", "0.100000", "0.10", "0.1000000000000000055",
    "0.10000000000000000555111512312578270211815834045410156250", "0.100000000000000012",
    "0.1000000000000000124900090270330110797658562660217285156250",
    "0.1000000000000000124900090270330110797658562660217285156251",
    "0.10000000000000002", "0.100001", "0.11", "0.2", "0.099999", "0.09", "1", "00.05", "000.2", "3.5",
    "0.100000000000000005551115123125782702118158340454101562500000000000000000000000000000001",
]


def make_graph(rng: random.Random, n_nodes: int, max_len: int = 9, id0: int = 1,
               extra_links: float = 0.3, crlf: bool = False, weird: bool = False):
    """Returns (gfa_text, ids(list[int]), lens(dict id->len), links(list[(a,b)]))."""
    ids = list(range(id0, id0 + n_nodes))
    if weird and n_nodes > 4:
        # a hole in the id space
        ids = [i for i in ids if i != id0 + 2]
    lens = {i: rng.randint(1, max_len) for i in ids}
    nl = "\r\n" if crlf else "\n"
    out = ["H\tVN:Z:1.1"]
    for i in ids:
        seq = "".join(rng.choice("ACGT") for _ in range(lens[i]))
        if rng.random() < 0.3:
            out.append(f"S\t{i}\t{seq}\tEX:Z:T{rng.randint(1, 3)}_R1.{rng.randint(1, 5)}")
        else:
            out.append(f"S\t{i}\t{seq}")
    links = []
    for a, b in zip(ids, ids[1:]):
        if rng.random() < 0.9:
            links.append((a, b))
    for _ in range(int(extra_links * n_nodes)):
        a, b = rng.choice(ids), rng.choice(ids)
        links.append((a, b))           # may duplicate, may be a self loop, may go backwards
    for a, b in links:
        if weird and rng.random() < 0.1:
            out.append(f"L\t{a}\t-\t{b}\t-\t*")
        elif rng.random() < 0.2:
            out.append(f"L\t{a}\t+\t{b}\t+\t*\tJN:Z:T1_R1.{rng.randint(1, 4)}.{rng.randint(2, 5)}")
        else:
            out.append(f"L\t{a}\t+\t{b}\t+\t*")
    if weird:
        out.insert(rng.randint(1, len(out)), "")
        out.insert(rng.randint(1, len(out)), "L")
        out.insert(rng.randint(1, len(out)), "W\tsample\t1\tchr\t0\t10\t>1>2")
        out.append(f"L\t{ids[0]}\t+\t99999\t+\t*")      # dangling link: never counted
        out.append("  # trailing comment with spaces  ")
    out.append("P\tT1_R1\t" + ",".join(f"{i}+" for i in ids[: min(4, len(ids))]) + "\t*")
    text = nl.join(out) + nl
    return text, ids, lens, links


def _rand_bases(rng, n, upper=False):
    s = "".join(rng.choice("acgt") for _ in range(n))
    return s.upper() if upper else s


def _ops_covering(rng: random.Random, demand: int, style: str):
    """A cs string whose op lengths sum to >= demand (so the walk never starves)."""
    if demand <= 0:
        demand = 1
    if style == "perfect":
        return f":{demand + rng.randint(0, 2)}"
    parts = []
    total = 0
    first = True
    target = demand + rng.randint(0, 3)
    while total < target:
        r = rng.random()
        if style == "mismatch":
            kinds = ":*" if not first else ":*"
            k = rng.choice(kinds) if r < 0.5 else ":"
        else:
            k = rng.choice(":::**--++=")
        if k == ":":
            n = rng.randint(1, 12)
            parts.append(f":{n}")
        elif k == "*":
            n = 1
            parts.append("*" + _rand_bases(rng, 2))
        elif k == "-":
            n = rng.randint(1, 9)
            parts.append("-" + _rand_bases(rng, n))
        elif k == "+":
            n = rng.randint(1, 5)
            parts.append("+" + _rand_bases(rng, n))
        else:
            n = rng.randint(1, 6)
            parts.append("=" + _rand_bases(rng, n, upper=True))
        total += n
        first = False
    if len(parts) == 2 and parts[0][0] == ":" and parts[1][0] == "+":
        parts.append(":3")      # else cigar_clipping would drop the '+' and the walk could starve
    if len(parts) == 2 and parts[0][0] == "+" and parts[1][0] == ":":
        parts[1] = ":" + str(int(parts[1][1:]) + len(parts[0]) - 1)
    return "".join(parts)


def make_record(rng: random.Random, ids, lens, name: str, safe: bool = True):
    """One GAF record (without newline)."""
    r = rng.random()
    mapq = rng.choice([60, 60, 60, 60, 40, 20, 19, 5, 0, 255])
    if r < 0.04:
        # unmapped
        return f"{name}\t100\t0\t0\t+\t*\t0\t0\t0\t0\t0\t{rng.choice([0, 0, 0, 60])}"
    n = rng.choice([1, 1, 2, 2, 3, 3, 4, 5, 6, 8])
    # a walk: mostly increasing ids, sometimes jumps / repeats
    start_i = rng.randrange(len(ids))
    nodes = [ids[start_i]]
    while len(nodes) < n:
        q = rng.random()
        cur = ids.index(nodes[-1])
        if q < 0.7 and cur + 1 < len(ids):
            nodes.append(ids[cur + 1])
        elif q < 0.85 and cur + 2 < len(ids):
            nodes.append(ids[cur + 2])
        elif q < 0.9:
            nodes.append(nodes[-1])             # consecutive duplicate (collapsed)
        elif q < 0.95:
            nodes.append(rng.choice(ids))       # arbitrary jump / revisit
        else:
            break
    # deduped list as the reference sees it
    ded = []
    for x in nodes:
        if not ded or ded[-1] != x:
            ded.append(x)
    rev = rng.random() < 0.5
    sep = "<" if rev else ">"
    path = "".join(f"{sep}{x}" for x in nodes)
    plen = sum(lens[x] for x in nodes)
    first_len, last_len = lens[ded[0]], lens[ded[-1]]
    q = rng.random()
    if q < 0.85:
        start = rng.randint(0, max(0, first_len - 1))
    else:
        start = rng.randint(0, first_len + 2)      # may empty the first node
    q = rng.random()
    if q < 0.85:
        end_rel = rng.randint(0, last_len)
    else:
        end_rel = rng.randint(0, last_len + 3)     # may empty the last node
    pend = plen - end_rel
    # demand as the reference computes it
    demand = 0
    for i, x in enumerate(ded):
        L = lens[x]
        if i == 0:
            L -= start
        if i == len(ded) - 1:
            L = L - end_rel + 1
        if L > 0:
            demand += L
    style = rng.choice(["perfect", "perfect", "perfect", "mismatch", "wild", "wild"])
    q = rng.random()
    tags = []
    cs = None
    if q < 0.06:
        # 2-op clipping forms (cigar_clipping); +k at the front shifts start by k
        k = rng.randint(1, 4)
        if rng.random() < 0.5:
            cs = f"+{_rand_bases(rng, k)}:{demand + 2}"
        else:
            cs = f":{demand + 2}+{_rand_bases(rng, k)}"
    elif q < 0.09 and len(ded) == 1 and lens[ded[0]] - start - end_rel + 1 <= 1:
        cs = None                                   # absent tag -> [('*', 1)]
    else:
        cs = _ops_covering(rng, demand, style)
        if not safe and rng.random() < 0.5:
            cs = _ops_covering(rng, max(1, demand // 2), style)   # may starve -> crash
    if cs is not None:
        tags.append("cs:Z:" + cs)
    dv = rng.choice(DV_CHOICES) if rng.random() < 0.5 else "0.%06d" % rng.randint(0, 120000)
    if safe or rng.random() < 0.8:
        tags.append("dv:f:" + dv)
    tags.append(f"AS:i:{rng.randint(0, 150)}")
    if rng.random() < 0.5:
        tags.append("fn:Z:" + name + "_mate")
    if rng.random() < 0.3:
        tags.append("pd:b:1")
    if rng.random() < 0.7:
        tags.sort()
    else:
        rng.shuffle(tags)
    cols = [name, "100", "0", "100", rng.choice("+-"), path, str(plen), str(start), str(pend),
            str(rng.randint(0, 100)), str(rng.randint(0, 100)), str(mapq)] + tags
    return "\t".join(cols)


def make_case(seed: int, n_nodes: int = 12, n_reads: int = 40, weird: bool = False, crlf: bool = False,
              trailing_newline: bool = True):
    """Returns (gfa_text, gaf_text) with only safe records."""
    rng = random.Random(seed)
    gfa, ids, lens, _ = make_graph(rng, n_nodes, crlf=crlf, weird=weird, id0=rng.choice([1, 1, 7, 1000]))
    nl = "\r\n" if crlf else "\n"
    recs = [make_record(rng, ids, lens, f"r{seed}_{k}") for k in range(n_reads)]
    gaf = nl.join(recs)
    if trailing_newline and recs:
        gaf += nl
    return gfa, gaf


def add_quality_tags(gaf: str, seed: int, share: float = 0.7) -> str:
    """Insert a bq:Z: tag (Phred+33 characters '!'..'K', ':' '<' '>' among them, 20-300 of them) somewhere among the tags
    of `share` of the records: what vg mpmap writes for FASTQ reads, inert for both regexes of the reference."""
    rng = random.Random(seed ^ 0xB9B9)
    out = []
    for line in gaf.split("\n"):
        end = "\r" if line.endswith("\r") else ""
        t = (line[:-1] if end else line).split("\t")
        if len(t) >= 12 and rng.random() < share:
            q = "".join(chr(rng.randint(33, 75)) for _ in range(rng.randint(20, 300)))
            t.insert(rng.randint(12, len(t)), "bq:Z:" + q)
        out.append("\t".join(t) + end)
    return "\n".join(out)


def make_risky_case(seed: int):
    """Small graph, one safe prefix and one possibly-crashing record at the end."""
    rng = random.Random(seed ^ 0x5EED)
    gfa, ids, lens, _ = make_graph(rng, 8)
    recs = [make_record(rng, ids, lens, f"s{seed}_{k}") for k in range(3)]
    q = rng.random()
    if q < 0.5:
        recs.append(make_record(rng, ids, lens, f"x{seed}", safe=False))
    elif q < 0.6:
        recs.append(f"x{seed}\t10\t0\t10\t+\t>{ids[0]}<{ids[1]}\t20\t0\t5\t5\t5\t60\tcs:Z::5\tdv:f:0")
    elif q < 0.7:
        recs.append(f"x{seed}\t10\t0\t10\t+\t>424242\t20\t0\t5\t5\t5\t60\tcs:Z::5\tdv:f:0")
    elif q < 0.8:
        recs.append(f"x{seed}\t10\t0\t10\t+\t>{ids[0]}\t20\t0\t5\t5\t5")
    elif q < 0.9:
        recs.append(f"x{seed}\t10\t0\t10\t+\t>{ids[0]}\t20\tzz\t5\t5\t5\t60\tcs:Z::5\tdv:f:0")
    else:
        recs.append(f"x{seed}\t10\t0\t10\t+\t>{ids[0]}\t20\t0\t5\t5\t5\t6x\tcs:Z::5\tdv:f:0")
    return gfa, "\n".join(recs) + "\n"


def spread_ids(gfa: str, gaf: str, pivot: int, shift: int):
    """Move every node id > pivot up by `shift`: links across the gap are farther than an inline delta (2^15)."""
    import re

    def mv(tok):
        return str(int(tok) + shift) if tok.isdigit() and int(tok) > pivot else tok

    g = []
    for line in gfa.split("\n"):
        t = line.split("\t")
        if t[0] == "S" and len(t) >= 3:
            t[1] = mv(t[1])
        elif t[0] == "L" and len(t) >= 5:
            t[1], t[3] = mv(t[1]), mv(t[3])
        elif t[0] == "P" and len(t) >= 3:
            t[2] = ",".join(mv(x[:-1]) + x[-1] if x[:-1].isdigit() else x for x in t[2].split(","))
        g.append("\t".join(t))
    a = []
    for line in gaf.split("\n"):
        t = line.split("\t")
        if len(t) > 5:
            t[5] = re.sub(r"\d+", lambda m: mv(m.group(0)), t[5])
        a.append("\t".join(t))
    return "\n".join(g), "\n".join(a)
