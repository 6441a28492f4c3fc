#!/usr/bin/env python3
"""bench.py -- pantas `augment` hot path on B200: alignments/s and GAF GB/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (GAF records -> NC/RC/IL/OL counters) over one
synthetic GAF.  At N=1 the workload is BASELINE.json configs[1]: a 10 M-alignment
GAF over the dm-full synthetic annotated spliced pangenome (no `vg` exists here,
so graph and reads come from pantas_b200/synth).  At N>1
each rank parses its own 10 M-alignment byte range of an N x 10 M GAF (weak
scaling) and the step ends with the one-shot NCCL reduction of the counters.

  value     alignments/s, whole job, GAF already resident in HBM
  e2e       same through the host-buffer C-ABI call: pinned host GAF -> H2D ->
            kernels -> export -> D2H of the reduced counters
  roofline  augment_team_kernel: GAF bytes parsed / kernel time vs measured HBM copy peak
  cpu_baseline  the CPU oracle port (oracle/augment_oracle.c) on a bounded sample, 1 core
`--impl reference` times that CPU port on all host cores instead (the reference
itself is pure Python and is not present on the GPU box; its measured speed in
the build container is in BASELINE.md).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
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
    "gene-panel-10M": ("gene-panel", 5_000_000, 1005),
    "tiny-20k": ("tiny", 10_000, 7),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PANTAS_BENCH_WORKLOAD", "dm-full-10M"), choices=list(WORKLOADS))
    ap.add_argument("--cpu-sample-lines", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_traffic(workload: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_roofline_traffic.json")) as f:
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
        self.gpu = gpu_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
            if ts < t0 - 0.15 or ts > t1 + 0.15:
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


def make_inputs(workload: str, rank: int, world: int):
    """-> (SynthGraph, pinned uint8 torch tensor with this rank's shard, nbytes, n_lines)"""
    import torch

    from pantas_b200.synth import SynthGraph

    preset, pairs, seed = WORKLOADS[workload]
    sg = SynthGraph(preset, seed=seed)
    buf, n_lines = sg.gaf(pairs, first_pair=rank * pairs, threads=max(1, (os.cpu_count() or 8) // max(world, 1)))
    n = int(buf.shape[0])
    pinned = torch.empty(n + 64, dtype=torch.uint8).pin_memory()
    pinned[:n] = torch.from_numpy(buf)
    return sg, pinned, n, n_lines


def cpu_baseline(sg, gaf_np: np.ndarray, n_lines_target: int, threads: int):
    """Time the CPU oracle's GAF loop (REF:138-371 restated in C) on the first ~n_lines_target lines."""
    import tempfile

    from oracle.oracle import run_oracle
    from pantas_b200.shard import shard_bounds_bytes

    # cut the sample at a line boundary
    approx = min(gaf_np.shape[0], int(n_lines_target * 270))
    nl = np.flatnonzero(gaf_np[:approx] == 10)
    if nl.size == 0:
        return None
    if nl.size > n_lines_target:
        end = int(nl[n_lines_target - 1]) + 1
        lines = n_lines_target
    else:
        end = int(nl[-1]) + 1
        lines = int(nl.size)
    sample = gaf_np[:end]
    with tempfile.TemporaryDirectory() as d:
        gp = os.path.join(d, "g.gfa")
        sg.write_gfa(gp)
        with open(gp, "rb") as f:
            gfa = f.read()
    bounds = shard_bounds_bytes(sample, threads)
    results = [None] * threads

    def work(k):
        results[k] = run_oracle(sample[bounds[k]:bounds[k + 1]], gfa, 20, write_output=False)

    t0 = time.time()
    ths = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = time.time() - t0
    assert all(r.rc == 0 for r in results), [r.err for r in results]
    secs = max(r.gaf_seconds for r in results)          # GAF loop only; GFA load excluded
    return {"lines": lines, "bytes": end, "gaf_loop_s": secs, "wall_s": wall,
            "rej": sum(r.rej for r in results)}


def run_reference_arm(args):
    """CPU port of the reference's augment loop on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pantas_b200.synth import SynthGraph

    preset, pairs, seed = WORKLOADS[args.workload]
    sg = SynthGraph(preset, seed=seed)
    cores = os.cpu_count() or 1
    per_step_pairs = max(10_000, min(pairs, args.cpu_sample_lines * cores // 2 // 2))
    buf, n_lines = sg.gaf(per_step_pairs, first_pair=0, threads=cores)
    times = []
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_baseline(sg, buf, n_lines, cores)
        if i >= args.warmup:
            times.append(res["gaf_loop_s"])
    t = float(np.mean(times))
    value = res["lines"] / t
    sample = (f"first {res['lines']} records of the {args.workload} GAF per step, sharded over {cores} threads at line "
              "boundaries (GAF loop only; GFA load/write and the cross-shard merge excluded)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "alignments/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": args.workload, "note": "CPU oracle port of the Python reference (reference itself is "
                   "not on the GPU box; BASELINE.md has its measured speed)"},
        "gaf_gb_per_s": res["bytes"] / t / 1e9,
        "cpu_baseline": {"value": value, "unit": "alignments/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
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

    from pantas_b200.engine import AugmentEngine

    K, W = args.steps, max(args.warmup, 3)
    sg, pinned, nbytes, n_lines = make_inputs(args.workload, rank, world)
    graph = sg.graph()
    eng = AugmentEngine(local_rank)
    eng.set_graph(graph)
    eng.profile(True)

    # global file offset of this rank's shard
    file_off = 0
    if world > 1:
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([nbytes], dtype=torch.int64, device=dev))
        file_off = int(sum(int(s.item()) for s in sizes[:rank]))
        total_lines_t = torch.tensor([n_lines, nbytes], dtype=torch.int64, device=dev)
        dist.all_reduce(total_lines_t)
        total_lines, total_bytes = int(total_lines_t[0].item()), int(total_lines_t[1].item())
    else:
        total_lines, total_bytes = n_lines, nbytes

    gaf_dev = torch.empty(((nbytes + 15) // 16) * 16 + 16, dtype=torch.uint8, device=dev)
    gaf_dev[:nbytes].copy_(pinned[:nbytes], non_blocking=True)
    torch.cuda.synchronize()
    n, e = graph.n_nodes, graph.n_edges

    def reduce_step():
        """the one-shot counter reduction that ends a multi-GPU job"""
        sums, stamps, novel, sparse = eng.export_device()
        if world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            dist.all_reduce(stamps, op=dist.ReduceOp.MIN)
        return sums, stamps, novel, sparse

    def device_step():
        eng.process_device(gaf_dev, nbytes, file_off, 20)
        if world > 1:
            reduce_step()

    pinned_out = {}

    def pin_out(k, t):
        h = pinned_out.get(k)
        if h is None or h.numel() < t.numel():
            h = torch.empty(max(t.numel(), 1), dtype=t.dtype).pin_memory()
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
        if world > 1:
            out = reduce_step()
            out = tuple(pin_out(k, t) for k, t in enumerate(out))           # D2H of the step's (reduced) result
            torch.cuda.synchronize()
        else:
            out = eng.export_host()                                         # D2H of the step's result, pinned host buffers
        return sum(t.numel() * t.element_size() for t in out)

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
    launches = eng.stats()["kernel_launches"] - launches0 - 4 * K      # minus the reset kernels
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
            torch.cuda.synchronize()
            ms2.append(1e3 * (time.perf_counter() - t0))
        e2e_ms = float(np.mean(ms2))
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e = {"value": total_lines / (e2e_ms / 1e3), "unit": "alignments/s", "h2d_bytes_per_step": int(nbytes),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "gaf_gb_per_s": total_bytes / e2e_ms / 1e6}
        eng.check_data_error()
    clocks = sampler.window(t_wall0, t_wall1) if sampler else None
    if sampler:
        sampler.stop()

    if rank == 0:
        peak, peak_src = measured_peaks()
        kern_avg_ms = kern_ms / max(kern_n, 1)
        achieved = nbytes / (kern_avg_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": total_lines / (step_ms / 1e3), "unit": "alignments/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": args.workload, "graph_nodes": n, "graph_links": e,
                       "alignments_per_gpu": n_lines, "gaf_bytes_per_gpu": nbytes,
                       "bytes_per_alignment": nbytes / n_lines, "partition": f"byte-range x{world}",
                       "l2": "input (GAF bytes per GPU) is larger than L2; no flush needed",
                       "timing": "per-step CUDA events on the launching stream; counter reset outside the events"},
            "gaf_gb_per_s": total_bytes / step_ms / 1e6,
            "wall_ms_per_step_incl_reset": 1e3 * (t_wall1 - t_wall0) / K,
            "gpu_launches": int(launches),
            "deferred_records": st["deferred_lines"],
            "clocks": clocks,
            "e2e": e2e,
            "roofline": {"bound": "hbm", "kernel": "augment_team_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload) if world == 1 else None,
                         "peak_source": peak_src, "kernel_ms": kern_avg_ms,
                         "algorithmic_bytes_per_launch": int(nbytes)},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(sg, pinned.numpy()[:nbytes], args.cpu_sample_lines, 1)
            if cb:
                line["cpu_baseline"] = {
                    "value": cb["lines"] / cb["gaf_loop_s"], "unit": "alignments/s", "cores": 1, "kind": "port",
                    "sample": f"first {cb['lines']} records of the same GAF, GAF loop only (oracle/augment_oracle.c, "
                              f"{cb['gaf_loop_s']:.2f} s); host has {os.cpu_count()} cores",
                    "gaf_gb_per_s": cb["bytes"] / cb["gaf_loop_s"] / 1e9}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
