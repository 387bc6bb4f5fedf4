#!/usr/bin/env python3
"""Benchmark of the SPR placement-cost hot path on B200 (contract: see the task's bench.py section).

Workload (config.workload): synthetic 29,903-bp reference, N_SEQ sequences (~10 differences each),
UNREST + per-site rate variation (BASELINE.json configs[2] shape); the tree's four genome-list
families are built on the device; one STEP scores, for every non-root node of the tree (the subtree
an SPR move would prune), every branch within RADIUS hops of its parent against the stored
mid-branch lists -- appendProbNode(probVectTotUp[t], probVect[s], isTip[s], dist[s]), the phase-1 call of
findBestParentTopology (MAPLEv0.7.5.4.py:7011/7223) -- i.e. one SPR candidate placement per pair.

value : candidate placements / s, arguments and lists resident in HBM when the timed region starts
e2e   : the same through the host-buffer C-ABI call (arguments in pinned host memory, copies and the
        read-back of the scores inside the timed region)
--impl reference : the CPU restatement of the reference algorithm (oracle/, OpenMP over all host cores)
        on a bounded sample of the same pairs.  The reference itself is a pure-Python script that
        cannot travel to the GPU box; its CPython/pypy3 rates measured while surveying are in BASELINE.md.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "spr_candidate_placements_per_sec"
UNIT = "placements/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nseq", type=int, default=int(os.environ.get("MAPLE_BENCH_NSEQ", 100000)))
    ap.add_argument("--radius", type=int, default=int(os.environ.get("MAPLE_BENCH_RADIUS", 13)))
    ap.add_argument("--cpu-sample", type=int, default=4_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def build_problem(args, device_index, seed):
    import torch
    from maple_b200.engine import MapleEngine
    from maple_b200.genome_list import pack_lists
    from maple_b200.synthetic import generate
    from maple_b200.tree import DeviceTree
    from maple_b200.workloads import neighbourhood_pairs, algorithmic_bytes
    t0 = time.time()
    d = generate(args.nseq, lRef=29903, mean_diffs=10.0, rate_variation=True, seed=seed)
    eng = MapleEngine(d.model, device_index)
    tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
    tips = pack_lists(d.tip_lists, d.model.lRef, d.model.usingErrorRate)
    tree.recalculate_all_lists(d.tip_nodes, tips)
    s, p, c, tip, bl = neighbourhood_pairs(tree, args.radius)
    torch.cuda.synchronize()
    info = {"setup_s": round(time.time() - t0, 1), "nodes": tree.n, "pairs": int(p.numel()), "searches": int(torch.unique(s).numel()),
            "arena_bytes": tree.arena.used_bytes(), "alg_bytes": algorithmic_bytes(tree, s, p, c),
            "mean_parent_entries": float(tree.arena.nkeys[p.long()].float().mean().item()),
            "mean_child_entries": float(tree.arena.nkeys[c.long()].float().mean().item())}
    return d, eng, tree, (s, p, c, tip, bl), info


def run_reference(args):
    """CPU arm: the oracle port of the path over all host cores, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle.oracle import Oracle
    dev = 0
    d, eng, tree, (s, p, c, tip, bl), info = build_problem(args, dev, seed=1)
    host = tree.arena.to_host()
    n = min(args.cpu_sample, int(p.numel()))
    # the sample is a contiguous run of whole searches from the middle of the batch
    start = (int(p.numel()) - n) // 2
    sl = slice(start, start + n)
    pa, ca, ta, ba = (x[sl].cpu().numpy() for x in (p, c, tip, bl))
    orc = Oracle(d.model)
    for _ in range(max(1, min(args.warmup, 2))):
        orc.append_batch(host, pa[: n // 8], ca[: n // 8], ta[: n // 8], ba[: n // 8])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.append_batch(host, pa, ca, ta, ba)
    dt = (time.perf_counter() - t0) / args.steps
    cores = orc.num_threads()
    val = n / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "nseq": args.nseq, "radius": args.radius, "pairs_per_step": n},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d consecutive candidate pairs (whole searches) of the %d-pair step, oracle/maple_oracle.c "
                                       "with OpenMP on %d threads" % (n, int(p.numel()), cores)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(args):
    return ("synthetic 29903-bp, %d seqs ~10 diffs, UNREST+rateVariation; phase-1 SPR candidate scoring over stored "
            "mid-branch lists, radius %d" % (args.nseq, args.radius))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: maple_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d, eng, tree, (s, p, c, tip, bl), info = build_problem(args, local, seed=1)
    dev = eng.device
    # strong scaling: searches (pruned nodes) are dealt round-robin to ranks like coreNum[node]==corNum (:9619)
    if world > 1:
        mine = (s.long() % world) == rank
        s, p, c, tip, bl = s[mine], p[mine].contiguous(), c[mine].contiguous(), tip[mine].contiguous(), bl[mine].contiguous()
    n = int(p.numel())
    out = torch.empty(n, dtype=torch.float64, device=dev)
    n_total = torch.tensor([n], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(n_total)
    n_total = int(n_total.item())
    # per-search best candidate (what a search reports); dense [nNodes] so that ranks can all-reduce it
    best = torch.full((tree.n,), float("-inf"), dtype=torch.float64, device=dev)
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)  # > 126 MB L2

    def step():
        eng.append_prob_batch(p, c, tip, bl, out=out)
        best.fill_(float("-inf"))
        best.scatter_reduce_(0, s.long(), out, reduce="amax")
        if world > 1:
            dist.all_reduce(best, op=dist.ReduceOp.MAX)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 5.0:  # nvidia-smi takes a moment to emit its first line
            time.sleep(0.05)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    l0 = eng.launches
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations
        ev[k][0].record()
        eng.append_prob_batch(p, c, tip, bl, out=out)
        ev[k][1].record()
        best.fill_(float("-inf"))
        best.scatter_reduce_(0, s.long(), out, reduce="amax")
        if world > 1:
            dist.all_reduce(best, op=dist.ReduceOp.MAX)
        ev[k][2].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launches - l0
    step_ms = sum(a.elapsed_time(z) for a, _, z in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in ev) / args.steps
    t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n_total * args.steps / (total_ms / 1e3)

    # ---- e2e: host buffers through the C-ABI host call, copies inside the timed region
    hp = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True) for x in (p, c, tip, bl)]
    for h, x in zip(hp, (p, c, tip, bl)):
        h.copy_(x)
    hout = torch.empty(n, dtype=torch.float64, pin_memory=True)
    hnp = [h.numpy() for h in hp]
    for _ in range(2):
        eng.append_prob_batch_host(*hnp, out=hout.numpy())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.append_prob_batch_host(*hnp, out=hout.numpy())
        _ = float(hout[0])
    torch.cuda.synchronize()
    e2e_dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_val = n_total / float(e2e_dt.item())
    if rank == 0:
        # keep the GPU busy with the timed kernel a little longer so that the 100 ms sampler sees it under load
        t_busy = time.time()
        while time.time() - t_busy < 1.0:
            eng.append_prob_batch(p, c, tip, bl, out=out)
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = info["alg_bytes"] / world / (kern_ms / 1e3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "nseq": args.nseq, "radius": args.radius, "pairs_per_step": n_total,
                   "searches_per_step": info["searches"], "nodes": info["nodes"], "arena_MB": round(info["arena_bytes"] / 1e6, 1),
                   "mean_parent_entries": round(info["mean_parent_entries"], 2), "mean_child_entries": round(info["mean_child_entries"], 2),
                   "l2": "flushed between timed iterations (160 MB memset)", "parallelism": "searches dealt round-robin to %d GPU(s); "
                   "one all-reduce(max) of the dense per-node best score" % world},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": n * 17, "d2h_bytes_per_step": n * 8,
                "note": "maple_append_prob_batch_host: pinned host argument arrays in, scores out, per rank"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "k_append", "kernel_ms": kern_ms,
                     "alg_bytes_per_launch": info["alg_bytes"] // world,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle.oracle import Oracle
        host = tree.arena.to_host()
        ns = min(args.cpu_sample, n)
        start = (n - ns) // 2
        sl = slice(start, start + ns)
        pa, ca, ta, ba = (x[sl].cpu().numpy() for x in (p, c, tip, bl))
        orc = Oracle(d.model)
        orc.append_batch(host, pa[: ns // 8], ca[: ns // 8], ta[: ns // 8], ba[: ns // 8])
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or time.perf_counter() - t0 < 10.0:
            ref = orc.append_batch(host, pa, ca, ta, ba)
            reps += 1
            if time.perf_counter() - t0 > 30.0:
                break
        cdt = (time.perf_counter() - t0) / reps
        got = out[sl].cpu().numpy()
        fin = np.isfinite(ref)
        line["cpu_baseline"] = {"value": ns / cdt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                "sample": "%d consecutive candidate pairs of the step x %d repeats, oracle/maple_oracle.c (C, OpenMP)"
                                          % (ns, reps),
                                "max_abs_diff_vs_gpu": float(np.max(np.abs(got[fin] - ref[fin]))) if fin.any() else 0.0,
                                "inf_pattern_equal": bool(np.array_equal(np.isfinite(got), fin))}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
