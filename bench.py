#!/usr/bin/env python3
"""bench.py -- pantas `augment` hot path on B200: alignments/s and GAF GB/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (GAF records -> NC/RC/IL/OL counters, exported in the C-ABI layout) over one synthetic
GAF.  At N=1 the workload is BASELINE.json configs[1]: a 10 M-alignment GAF over the dm-full synthetic annotated spliced
pangenome (no `vg` exists here, so graph and reads come from pantas_b200/synth).  At N>1 each rank
parses its own 10 M-alignment byte range of an N x 10 M GAF (weak scaling) and the step ends with the one-shot NCCL
reduction of the counters to rank 0.

  value         alignments/s, whole job, GAF already resident in HBM (kernels + fold + export [+ reduction])
  e2e           same through the host-buffer C-ABI call: pinned host GAF -> H2D -> kernels -> export [-> reduction] -> D2H
  roofline      augment_team_kernel: GAF bytes parsed / kernel time (CUDA events around the kernel) vs measured HBM copy peak
  parity_checked  the augmented GFA of a prefix of the benchmarked GAF (both GFA passes on the device) is byte-identical
                  to the CPU oracle's; at N>1: every rank's prefix, reduced, against one oracle run
  cpu_baseline  the reference script itself (oracle/_ref, 1 core: it is single-threaded) on a bounded sample, with the
                C port of it (oracle/augment_oracle.c, 1 core) beside it
  cli           wall time of the drop-in script, file -> stdout, with the GFA passes on the device and in Python
`--impl reference` times the C port of the reference's loop on all host cores (one shared parse of the GFA, every thread
its own byte range of a bounded sample) and, when oracle/_ref holds it, the reference script on one core.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "augment GAF alignments/sec"
WORKLOADS = {
    # name: (preset, pairs per GPU, seed)
    "dm-full-10M": ("dm-full", 5_000_000, 1002),
    "dm-chr4-1M": ("dm-chr4", 500_000, 1001),
    "hs-chr1-10M": ("hs-chr1", 5_000_000, 1003),
    "hs-chr1-100M-strong": ("hs-chr1", 50_000_000, 1003),       # BASELINE.json configs[2]: 100 M alignments in total, split over the ranks
    "gene-panel-10M": ("gene-panel", 5_000_000, 1005),
    "gene-panel-100M": ("gene-panel", 50_000_000, 1005),         # BASELINE.json configs[4] at a fifth of its 500 M reads: Zipf-skewed coverage
    "hs-wg-25M": ("hs-wg", 12_500_000, 1004),                    # BASELINE.json configs[3] graph (2e8 nodes), 25 M alignments per GPU
    "hs-wg-125M": ("hs-wg", 62_500_000, 1004),                   # BASELINE.json configs[3] itself at 8 GPUs: 8 x 125 M = 1 G alignments
    "gene-panel-500M": ("gene-panel", 250_000_000, 1005),        # BASELINE.json configs[4] at its full 500 M reads (153 GB of GAF in HBM)
    "dm-full-10M-bq": ("dm-full", 5_000_000, 1002),               # the default workload with a bq:Z: quality tag per record (FASTQ reads)
    "tiny-20k": ("tiny", 10_000, 7),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PANTAS_BENCH_WORKLOAD", "dm-full-10M"), choices=list(WORKLOADS))
    ap.add_argument("--cpu-sample-lines", type=int, default=1_500_000, help="records of the C-port baseline / parity prefix")
    ap.add_argument("--ref-sample-lines", type=int, default=120_000, help="records the reference script itself is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the drop-in script wall-time measurement")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--tables", action="store_true", help="graph tables straight from the generator (no GFA text: for graphs whose "
                    "GFA would be tens of GB); implies --no-parity --no-cli")
    return ap.parse_args()


def config_of(workload: str, world: int, n_nodes=None, n_links=None, n_lines=None, nbytes=None):
    """The `config` object: the same keys in both arms."""
    preset, pairs, seed = WORKLOADS[workload]
    strong = workload.endswith("-strong")
    per_gpu = 2 * (pairs // world if strong else pairs)
    c = {"workload": workload, "graph": preset, "seed": seed, "alignments_per_gpu": per_gpu, "partition": f"byte-range x{world}",
         "l2": "input (GAF bytes per GPU) is larger than L2; no flush needed",
         "timing": "per-step CUDA events on the launching stream; counter reset outside the events"}
    if n_nodes is not None:
        c.update({"graph_nodes": n_nodes, "graph_links": n_links})
    if nbytes is not None:
        c.update({"gaf_bytes_per_gpu": nbytes, "bytes_per_alignment": nbytes / max(n_lines, 1)})
    return c


def measured_traffic(workload: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")) as f:
            d = json.load(f)
        if d.get("workload") == workload:
            return int(d["dram_bytes_read_per_launch"]) + int(d["dram_bytes_write_per_launch"])
    except (OSError, ValueError, KeyError):
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.3 or ts > t1 + 0.3:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


def make_inputs(workload: str, rank: int, world: int, pinned: bool = True, device=None):
    """-> (SynthGraph, uint8 GAF of this rank's shard (pinned torch tensor or numpy), nbytes, n_lines, chunks)

    Inputs beyond 6 GB (configs 3 and 4) are streamed when `device` is given: every generator piece goes straight into its
    own device buffer and only the first piece stays on the host (parity prefix), so host memory stays bounded;
    chunks = [(first byte, end byte, device tensor)], else None."""
    from pantas_b200.synth import SynthGraph

    preset, pairs, seed = WORKLOADS[workload]
    if workload.endswith("-strong"):
        pairs //= world
    if workload.endswith("-bq"):
        os.environ["PANTAS_SYNTH_BQ"] = "1"                  # (read when the generator is created)
    sg = SynthGraph(preset, seed=seed)
    threads = max(1, (os.cpu_count() or 8) // max(world, 1))
    piece = 2_500_000                                       # pairs per generator call: bounds the generator's own buffers
    stream = device is not None and pairs * 2 * 330 > (6 << 30)
    parts, chunks, n_lines, n = [], [], 0, 0
    t_gen = time.perf_counter()
    for p0 in range(0, pairs, piece):
        b, nl = sg.gaf(min(piece, pairs - p0), first_pair=rank * pairs + p0, threads=threads)
        n_lines += nl
        if stream:
            import torch

            m = int(b.shape[0])
            t = torch.empty(((m + 15) // 16) * 16 + 16, dtype=torch.uint8, device=device)
            t[:m].copy_(torch.from_numpy(b))
            chunks.append((n, n + m, t))
            if not parts:
                parts.append(b)
        else:
            parts.append(b)
        n += int(b.shape[0])
    if rank == 0:
        print(f"[bench] generator: {n_lines} records, {n / 1e9:.2f} GB per rank in {time.perf_counter() - t_gen:.1f} s "
              f"({threads} threads per rank)", file=sys.stderr)
    if not pinned:
        return sg, (np.concatenate(parts) if len(parts) > 1 else parts[0]), n, n_lines, None
    import torch

    if stream:
        return sg, torch.from_numpy(parts[0]), n, n_lines, chunks
    t = torch.empty(n + 64, dtype=torch.uint8)
    if n <= (6 << 30):                                      # (larger inputs stay pageable: the host-buffer leg is skipped for them)
        t = t.pin_memory()
    pos = 0
    for b in parts:
        t[pos:pos + b.shape[0]] = torch.from_numpy(b)
        pos += b.shape[0]
    return sg, t, n, n_lines, None


def cut_lines(gaf_np: np.ndarray, n_lines_target: int):
    """(end byte, lines) of the longest prefix with at most n_lines_target records."""
    approx = min(gaf_np.shape[0], int(n_lines_target * 400) + 4096)
    nl = np.flatnonzero(gaf_np[:approx] == 10)
    if nl.size == 0:
        return 0, 0
    k = min(n_lines_target, int(nl.size))
    return int(nl[k - 1]) + 1, k


def port_loop(og, sample: np.ndarray, threads: int):
    """The C port's GAF loop (REF:138-371) over `sample`, sharded over `threads` host threads at line boundaries, against one
    shared parse of the GFA.  -> (max loop seconds over the threads, rejected)"""
    from pantas_b200.shard import shard_bounds_bytes

    bounds = shard_bounds_bytes(sample, threads)
    results = [None] * threads

    def work(k):
        results[k] = og.run(sample[bounds[k]:bounds[k + 1]], 20, write_output=False)

    ths = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert all(r.rc == 0 for r in results), [r.err for r in results]
    return max(r.gaf_seconds for r in results), sum(r.rej for r in results)


def reference_script_time(gfa_path: str, gaf_np: np.ndarray, n_lines: int):
    """The reference script itself (1 core: it has no parallelism) on the first n_lines records; the GFA passes are
    timed apart with an empty GAF, so that the figure is the GAF loop like everything else here."""
    from oracle.oracle import reference_script

    script = reference_script()
    if script is None:
        return None
    end, lines = cut_lines(gaf_np, n_lines)
    if lines == 0:
        return None
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        ap, ep = os.path.join(d, "a.gaf"), os.path.join(d, "empty.gaf")
        gaf_np[:end].tofile(ap)
        open(ep, "wb").close()

        def run(gaf):
            t0 = time.time()
            p = subprocess.run([sys.executable, "-W", "ignore", script, gaf, gfa_path], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
            assert p.returncode == 0, p.stderr.decode()[-500:]
            return time.time() - t0

        t_empty = run(ep)
        t_full = run(ap)
    loop_s = max(t_full - t_empty, 1e-9)
    return {"lines": lines, "bytes": end, "gaf_loop_s": loop_s, "gfa_passes_s": t_empty, "wall_s": t_full}


def run_reference_arm(args):
    """The reference's CPU implementation of the path on the box's host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle.oracle import OracleGraph

    cores = os.cpu_count() or 1
    sg, buf, nbytes, n_lines, _ = make_inputs(args.workload, 0, world, pinned=False)
    with tempfile.TemporaryDirectory() as d:
        gp = os.path.join(d, "g.gfa")
        sg.write_gfa(gp)
        with open(gp, "rb") as f:
            gfa = f.read()
        ref = None if args.no_cpu_baseline else reference_script_time(gp, buf, args.ref_sample_lines)
    og = OracleGraph(gfa)                                     # ONE parse of the GFA, shared by every thread and step
    assert og.rc == 0, og.err
    per_step = max(20_000, min(n_lines, 100_000 * cores))     # a bounded sample: ~0.3 s of loop per step and core
    end, lines = cut_lines(buf, per_step)
    sample = buf[:end]
    times = []
    for i in range(args.warmup + args.steps):
        secs, _ = port_loop(og, sample, cores)
        if i >= args.warmup:
            times.append(secs)
    og.close()
    t = float(np.mean(times))
    value = lines / t
    cfg = config_of(args.workload, world, sg.n_nodes, sg.n_links, n_lines, nbytes)
    sample_txt = (f"first {lines} records of the {args.workload} GAF per step, byte-range sharded over {cores} host threads "
                  "(C port of the reference's loop, one shared GFA parse; GAF loop only, cross-shard merge excluded)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "alignments/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "strong" if args.workload.endswith("-strong") else "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": cfg, "gaf_gb_per_s": end / t / 1e9,
        "cpu_baseline": {"value": value, "unit": "alignments/s", "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if ref:
        line["reference_script"] = {"value": ref["lines"] / ref["gaf_loop_s"], "unit": "alignments/s", "cores": 1, "kind": "reference",
                                    "sample": f"the unmodified reference script (oracle/_ref) on the first {ref['lines']} records, 1 core; "
                                              f"GAF loop {ref['gaf_loop_s']:.1f} s, its two GFA passes {ref['gfa_passes_s']:.1f} s"}
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    numa = None
    if world > 1 and os.environ.get("PANTAS_NUMA", "1") != "0":
        # a rank's pinned buffers belong on its GPU's socket (first touch): bind before anything is allocated.  Not at N = 1:
        # the CPU baseline of that run wants every core.
        from pantas_b200.numa import bind_to_gpu_node

        numa = bind_to_gpu_node(local_rank)

    from pantas_b200.dist import reduce_results, rows_to_host
    from pantas_b200.engine import AugmentEngine
    from pantas_b200.gfa_device import DeviceGfa

    K, W = args.steps, max(args.warmup, 3)
    sg, pinned, nbytes, n_lines, chunks = make_inputs(args.workload, rank, world, device=dev)
    if args.tables:
        args.no_parity = args.no_cli = True
    if not pinned.is_pinned():
        args.no_e2e = args.no_cli = True
    tmp = tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    gfa_path = os.path.join(tmp.name, f"g{rank}.gfa")
    eng = AugmentEngine(local_rank)
    t0 = time.perf_counter()
    if args.tables:
        dg = None
        graph = sg.graph()
        eng.set_graph(graph)
    else:
        sg.write_gfa(gfa_path)
        t0 = time.perf_counter()
        dg = DeviceGfa.load(eng, gfa_path)                    # GFA pass 1 on the device (REF:121-126)
        dg.set_graph()
        graph = dg.graph
    torch.cuda.synchronize()
    gfa_load_ms = 1e3 * (time.perf_counter() - t0)
    eng.profile(True)
    n, e = graph.n_nodes, graph.n_edges

    # global file offset of this rank's shard
    file_off = 0
    if world > 1:
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([nbytes], dtype=torch.int64, device=dev))
        file_off = int(sum(int(s.item()) for s in sizes[:rank]))
        tot = torch.tensor([n_lines, nbytes], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        total_lines, total_bytes = int(tot[0].item()), int(tot[1].item())
    else:
        total_lines, total_bytes = n_lines, nbytes

    # chunks of at most 3 GiB (one launch each; pt_process_chunk takes < 3.75 GiB), cut at line ends; every chunk in its own
    # (16-byte aligned) device buffer
    view_all = pinned.numpy()
    pos = 0
    if chunks is None:
        chunks = []
    else:
        pos = nbytes                                          # (streamed input: the chunks are on the device already)
    while pos < nbytes:
        end = min(pos + (3 << 30), nbytes)
        if end < nbytes:
            w = view_all[end - (1 << 16):end]
            end = end - (1 << 16) + int(np.flatnonzero(w == 10)[-1]) + 1
        t = torch.empty(((end - pos + 15) // 16) * 16 + 16, dtype=torch.uint8, device=dev)
        for a in range(pos, end, 1 << 30):
            b = min(a + (1 << 30), end)
            t[a - pos:b - pos].copy_(pinned[a:b], non_blocking=True)
        chunks.append((pos, end, t))
        pos = end
    torch.cuda.synchronize()
    gaf_dev = chunks[0][2]

    def finish_step():
        """fold + export in the C-ABI layout, then (N > 1) the one-shot reduction to rank 0"""
        sums, stamps, novel, sparse = eng.export_device()
        if world > 1:
            return reduce_results(sums, stamps, novel, sparse, n, dst=0)
        return sums, stamps, novel, sparse

    def device_step():
        for a, b, t in chunks:
            eng.process_device(t, b - a, file_off + a, 20)
        return finish_step()

    pinned_out = {}

    def pin_out(k, t):
        h = pinned_out.get(k)
        if h is None or h.numel() < t.numel():
            h = torch.empty(max(t.numel(), 1), dtype=t.dtype, pin_memory=True)
            pinned_out[k] = h
        v = h[: t.numel()].view(t.shape)
        v.copy_(t, non_blocking=True)
        return v

    def host_step():
        stage = eng.stage_bytes
        view = pinned.numpy()
        pos = 0
        base = pinned.data_ptr()
        while pos < nbytes:
            end = min(pos + stage, nbytes)
            if end < nbytes:
                w = view[max(pos, end - (1 << 16)):end]
                nlp = np.flatnonzero(w == 10)
                end = max(pos, end - (1 << 16)) + int(nlp[-1]) + 1
            eng.process_host(base + pos, end - pos, file_off + pos, 20)
            pos = end
        out = finish_step()
        d2h = 0
        if rank == 0:                                          # D2H of the step's (reduced) result into pinned host buffers
            for k, t in enumerate(out):
                pin_out(k, t)
                d2h += t.numel() * t.element_size()
        torch.cuda.synchronize()
        return d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """per-step CUDA events on the launching stream; counters are reset outside the timed part"""
        ms = []
        for _ in range(steps):
            eng.reset()
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            fn()
            ev1.record()
            ev1.synchronize()
            ms.append(ev0.elapsed_time(ev1))
        return ms

    # ---- device-resident
    for _ in range(W):
        eng.reset()
        device_step()
    torch.cuda.synchronize()
    eng.check_data_error()
    eng.kernel_time_split()
    launches0 = eng.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3)
    barrier()
    t_wall0 = time.time()
    ms = timed(device_step, K)
    barrier()
    t_wall1 = time.time()
    kern_ms, slow_ms, kern_n = eng.kernel_time_split()                # the fast-path kernel alone / the per-record kernel
    launches = eng.stats()["kernel_launches"] - launches0 - 5 * K      # minus the reset kernels
    step_ms = float(np.mean(ms))
    if world > 1:
        t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t.item())
    eng.check_data_error()
    st = eng.stats()

    # ---- end to end through the host-buffer entry point
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            eng.reset()
            host_step()
        d2h = 0
        ms2 = []
        for _ in range(K):
            eng.reset()
            barrier()
            t0 = time.perf_counter()
            d2h = host_step()
            ms2.append(1e3 * (time.perf_counter() - t0))
        e2e_ms = float(np.mean(ms2))
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e = {"value": total_lines / (e2e_ms / 1e3), "unit": "alignments/s", "h2d_bytes_per_step": int(nbytes),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "gaf_gb_per_s": total_bytes / e2e_ms / 1e6,
               "h2d_gb_per_s_aggregate": total_bytes / e2e_ms / 1e6}
        eng.check_data_error()
    clocks = sampler.window(t_wall0, t_wall1) if sampler else None
    if sampler:
        sampler.stop()

    # ---- parity of what was just timed: a prefix of this rank's GAF, reduced over the ranks, against the CPU oracle
    parity = None
    if not args.no_parity:
        from oracle.oracle import OracleGraph

        pre_lines = args.cpu_sample_lines if world == 1 else 200_000
        end, lines = cut_lines(pinned.numpy()[:nbytes], pre_lines)
        pre_off = 0
        if world > 1:
            sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([end], dtype=torch.int64, device=dev))
            sizes = [int(s.item()) for s in sizes]
            pre_off = sum(sizes[:rank])
        eng.reset()
        eng.process_device(gaf_dev, end, pre_off, 20)
        eng.check_data_error()
        sums, stamps, novel, sparse = finish_step()
        if world > 1:
            m = max(sizes)
            mine = torch.zeros(m, dtype=torch.uint8, device=dev)
            mine[:end] = gaf_dev[:end]
            parts = [torch.empty(m, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
            dist.gather(mine, parts, dst=0)
            novel_h, sparse_h = rows_to_host(novel), rows_to_host(sparse)
            if rank == 0:
                prefix = np.concatenate([p[:s].cpu().numpy() for p, s in zip(parts, sizes)])
        else:
            prefix = pinned.numpy()[:end]
            novel_h = novel.cpu().numpy().view(np.uint64).reshape(-1, 3)
            sparse_h = sparse.cpu().numpy().view(np.uint64).reshape(-1, 3)
        if rank == 0:
            got = dg.render(sums, stamps, novel_h, sparse_h).numpy()
            with open(gfa_path, "rb") as f:
                og = OracleGraph(f.read())
            want = og.run(prefix, 20, write_output=True)
            og.close()
            assert want.rc == 0, want.err
            ref = np.frombuffer(want.out, dtype=np.uint8)
            same = got.shape == ref.shape and bool(np.array_equal(got, ref))
            parity = {"checked": same, "records": int(want.n_lines), "gfa_bytes": int(ref.shape[0]), "rejected": int(want.rej),
                      "what": "augmented GFA (GFA passes on the device) of a prefix of every rank's GAF, reduced to rank 0, "
                              "byte-compared with the CPU oracle (oracle/augment_oracle.c)"}
            assert same, "PARITY FAILURE: the benchmarked path does not reproduce the oracle's GFA"
            port_secs = want.gaf_seconds

    if rank == 0:
        peak, peak_src = measured_peaks()
        kern_avg_ms = kern_ms / max(kern_n, 1)                 # per launch (one launch per <= 3 GiB chunk)
        achieved = nbytes * K / (kern_ms / 1e3) / 1e9          # GAF bytes of the K timed passes / time inside the kernel
        cfg = config_of(args.workload, world, n, e, n_lines, nbytes)
        line = {
            "metric": METRIC, "value": total_lines / (step_ms / 1e3), "unit": "alignments/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong" if args.workload.endswith("-strong") else "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": cfg,
            "gaf_gb_per_s": total_bytes / step_ms / 1e6,
            "wall_ms_per_step_incl_reset": 1e3 * (t_wall1 - t_wall0) / K,
            "gpu_launches": int(launches),
            "numa": numa,
            "deferred_records": st["deferred_lines"],
            "gfa_pass1_device_ms": gfa_load_ms,
            "clocks": clocks,
            "e2e": e2e,
            "parity_checked": bool(parity and parity["checked"]),
            "parity": parity,
            "roofline": {"bound": "hbm", "kernel": "augment_team_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload) if world == 1 else None,
                         "peak_source": peak_src, "kernel_ms": kern_avg_ms,
                         "algorithmic_bytes_per_launch": int(nbytes // len(chunks)), "launches_per_step": len(chunks)},
        }
        if world == 1 and not args.no_cpu_baseline:
            gaf_np = pinned.numpy()[:nbytes]
            port = None
            if parity:
                port = {"value": parity["records"] / port_secs, "unit": "alignments/s", "cores": 1, "kind": "port",
                        "sample": f"first {parity['records']} records of the same GAF, GAF loop only "
                                  f"(oracle/augment_oracle.c, {port_secs:.2f} s)"}
            ref = reference_script_time(gfa_path, gaf_np, args.ref_sample_lines)
            if ref:
                line["cpu_baseline"] = {
                    "value": ref["lines"] / ref["gaf_loop_s"], "unit": "alignments/s", "cores": 1, "kind": "reference",
                    "sample": f"the unmodified reference script (oracle/_ref) on the first {ref['lines']} records of the same GAF, "
                              f"1 core (it is single-threaded): GAF loop {ref['gaf_loop_s']:.1f} s; its two GFA passes over the "
                              f"{n}-node graph {ref['gfa_passes_s']:.1f} s; host has {os.cpu_count()} cores",
                    "gfa_passes_s": ref["gfa_passes_s"], "port": port}
            elif port:
                line["cpu_baseline"] = port
        if world == 1 and not args.no_cli:
            line["cli"] = cli_wall_times(gfa_path, pinned.numpy()[:nbytes], tmp.name, n_lines)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    tmp.cleanup()


def cli_wall_times(gfa_path: str, gaf_np: np.ndarray, tmpdir: str, n_lines: int):
    """Wall time of the drop-in script (pantas:132 runs it), GAF + GFA files -> augmented GFA on stdout: with both GFA passes
    on the device (the product) and with the host passes of pantas_b200/gfa.py (what round 1 shipped)."""
    script = os.path.join(ROOT, "scripts", "alignments_augmentation_from_gaf.py")
    ap = os.path.join(tmpdir, "bench.gaf")
    gaf_np.tofile(ap)
    out = {"alignments": n_lines, "gaf_bytes": int(gaf_np.shape[0]), "gfa_bytes": os.path.getsize(gfa_path)}
    for name, env in (("device_gfa_passes_s", {"PANTAS_TIMING": "1"}), ("python_gfa_passes_s", {"PANTAS_GFA_PASSES": "host"})):
        t0 = time.time()
        with open(os.path.join(tmpdir, name + ".gfa"), "wb") as fo:
            p = subprocess.run([sys.executable, script, ap, gfa_path], stdout=fo, stderr=subprocess.PIPE, env=dict(os.environ, **env))
        out[name] = time.time() - t0
        for ln in p.stderr.decode(errors="replace").splitlines():
            if ln.startswith("timing: "):
                out["device_stages"] = ln[8:]              # the rest of the wall clock is interpreter start-up and exit
        if p.returncode != 0:
            out[name] = None
            out["error"] = p.stderr.decode()[-300:]
    try:
        a = open(os.path.join(tmpdir, "device_gfa_passes_s.gfa"), "rb").read()
        b = open(os.path.join(tmpdir, "python_gfa_passes_s.gfa"), "rb").read()
        out["outputs_identical"] = a == b
        out["out_bytes"] = len(a)
    except OSError:
        pass
    os.remove(ap)
    return out


if __name__ == "__main__":
    main()
