#!/usr/bin/env python3
"""Benchmark of the SPR search hot path on B200 (contract: the task's bench.py section).

Workload (config.workload): synthetic 29,903-bp reference, NSEQ (default 100,000) sequences ~10 differences each,
UNREST + per-site rate variation (BASELINE.json configs[2] shape).  The tree is the simulated one with
MAPLE-like branch lengths (substitutions/lRef, zero-length branches where nothing happened); its four genome-list
families are built on the device.  One STEP = one full SPR search round on that frozen tree, i.e. what the
reference does in Pool.map(startTopologyUpdatesParallel, ...) (MAPLEv0.7.5.4.py:12283-12312): for every non-root
node the current placement cost, then findBestParentTopology with the deep-search stop rules (allowedFails 4,
threshold 14*ln lRef, non-strict) including all merges, branch-length optimisations and phase-2 refinements.

metric : SPR candidate placements scored per second = phase-1 appendProbNode calls (:7011/:7223) / search time
value  : node list already on the device (lists, tree and model are always device-resident)
e2e    : node ids in pinned host memory -> device, search, 64-byte result records -> host, every step
--impl reference : the CPU restatement of the reference algorithm (oracle/maple_oracle.c, OpenMP over all host
         cores) running the same searches on a bounded sample of the nodes, on the same frozen tree built on the CPU
         (no GPU, no CUDA library on this path).  The reference itself is a pure-Python script that cannot travel to
         the GPU box; BASELINE.md has its CPython rates measured while surveying.
"""
import argparse
import json
import math
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
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nseq", type=int, default=int(os.environ.get("MAPLE_BENCH_NSEQ", 100000)))
    ap.add_argument("--round", default=os.environ.get("MAPLE_BENCH_ROUND", "deep"), choices=["fast", "deep"])
    ap.add_argument("--workload", default=os.environ.get("MAPLE_BENCH_WORKLOAD", "search"), choices=["search", "place"],
                    help="search: one SPR search round (the headline); place: one batch of new samples placed on the frozen tree "
                         "(findBestParentForNewSample, BASELINE.json config 5 shape)")
    ap.add_argument("--new-samples", type=int, default=int(os.environ.get("MAPLE_BENCH_NEW_SAMPLES", 50000)), help="--workload place: samples per batch")
    ap.add_argument("--cpu-searches", type=int, default=1500, help="searches in the CPU sample")
    ap.add_argument("--cpu-samples", type=int, default=400, help="--workload place: samples in the CPU sample")
    ap.add_argument("--error-model", action="store_true", default=bool(int(os.environ.get("MAPLE_BENCH_ERROR_MODEL", "0"))),
                    help="site-specific sequencing-error rates on (the reference's --estimateSiteSpecificErrorRate shape, BASELINE config 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the untimed side measurements reported under 'extra'")
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
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def round_params(args, lRef):
    from maple_b200.search import search_params
    L = math.log(lRef)
    if args.round == "fast":  # the reference's first, fast round (:222-225, :12149-12153)
        return search_params(lRef, True, 2, 6.0 * L)
    return search_params(lRef, False, 4, 14.0 * L)  # its deep rounds (:53-55, :12155-12159)


def workload_name(args):
    return ("synthetic 29903-bp, %d seqs ~10 diffs, UNREST+rateVariation%s, MAPLE-like branch lengths; one full SPR search round "
            "(%s stop rules) over every non-root node of the frozen tree"
            % (args.nseq, "+site-specific error rates" if args.error_model else "", args.round))


def build_problem(args, device_index):
    import torch
    from maple_b200.engine import MapleEngine
    from maple_b200.genome_list import pack_lists
    from maple_b200.search import dirty_nodes
    from maple_b200.synthetic import generate
    from maple_b200.tree import DeviceTree
    t0 = time.time()
    d = generate(args.nseq, lRef=29903, mean_diffs=10.0, rate_variation=True, error_model=args.error_model,
                 site_specific_errors=args.error_model, seed=1, ml_like_blens=True)
    eng = MapleEngine(d.model, device_index)
    tree = DeviceTree(eng, d.up, d.child0, d.child1, d.dist, d.root)
    tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, d.model.usingErrorRate))
    tree.prepare_search()
    nodes = dirty_nodes(tree)
    torch.cuda.synchronize()
    return d, eng, tree, nodes, round(time.time() - t0, 1)


def oracle_tree(d, tree):
    return {"up": d.up, "child0": d.child0, "child1": d.child1, "dist": tree.dist, "isTip": tree.isTip, "root": d.root}


def params_dict(p):
    return {f[0]: getattr(p, f[0]) for f in p._fields_ if f[0] != "reserved"}


def cpu_sample(nodes, k):
    """every (len/k)-th node of the pre-order node list: an unbiased spread over the tree"""
    import numpy as np
    if k >= len(nodes):
        return nodes
    return nodes[np.linspace(0, len(nodes) - 1, k).astype(np.int64)]


def build_problem_cpu(args):
    """The same frozen tree as build_problem, built without the CUDA library: the oracle's mergeVectors in level-synchronous
    batches (oracle/host_tree.py).  Verified equivalent: the oracle's searches on both give identical records."""
    import numpy as np
    from maple_b200.synthetic import generate
    from oracle.host_tree import build_tree_lists
    from oracle.oracle import Oracle
    t0 = time.time()
    d = generate(args.nseq, lRef=29903, mean_diffs=10.0, rate_variation=True, error_model=args.error_model,
                 site_specific_errors=args.error_model, seed=1, ml_like_blens=True)
    orc = Oracle(d.model)
    lists, dist, isTip = build_tree_lists(orc, d.up, d.child0, d.child1, d.dist, d.root, d.tip_nodes, d.tip_lists, d.model.lRef,
                                          d.model.usingErrorRate)
    out, stack = [], [int(d.root)]  # the nodes startTopologyUpdatesParallel visits (:9615-9626), in its pre-order
    while stack:
        n = stack.pop()
        if d.child0[n] >= 0:
            stack.append(int(d.child0[n]))
            stack.append(int(d.child1[n]))
        if n != d.root:
            out.append(n)
    ta = {"up": d.up, "child0": d.child0, "child1": d.child1, "dist": dist, "isTip": isTip, "root": d.root}
    return d, orc, ta, lists, np.array(out, np.int32), round(time.time() - t0, 1)


def run_reference(args):
    """CPU arm: no GPU, no CUDA library anywhere on this path."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d, orc, ta, host, nodes, setup_s = build_problem_cpu(args)
    orc.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every core of this process's affinity mask at every N
    p = round_params(args, d.model.lRef)
    sample = cpu_sample(nodes, args.cpu_searches)
    pd = params_dict(p)
    for _ in range(max(1, min(args.warmup, 1))):
        orc.search_batch(ta, host, pd, sample[: max(16, len(sample) // 8)], lazy_mode=1)
    t0 = time.perf_counter()
    tot = 0
    for _ in range(args.steps):
        rec = orc.search_batch(ta, host, pd, sample, lazy_mode=1)
        tot += int(rec["phase1"].sum())
    dt = time.perf_counter() - t0
    cores = orc.num_threads()
    val = tot / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "nseq": args.nseq, "round": args.round, "searches_per_step": int(len(sample)),
                       "placements_per_step": tot // args.steps, "setup_s": setup_s,
                       "note": "tree and lists built on the CPU (oracle/host_tree.py); the reference itself is a Python script "
                               "that cannot travel to this box, BASELINE.md has its CPython rates"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d of the %d searches of one round (evenly spread over the tree), oracle/maple_oracle.c "
                                       "search with OpenMP on %d threads" % (len(sample), len(nodes), cores)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def callers_extra(args, d, eng, tree, d_nodes):
    """Side measurements on the same device-resident tree, after everything that is timed for the headline and each guarded on its
    own: the other stop-rule setting of the search, and the callers either side of it: whole-tree log-likelihood
    (calculateTreeLikelihood as one merge batch), the branch-length sweep with frozen lists (traverseTreeToOptimizeBranchLengths
    fastPass as one maple_blen_batch launch) and the rebuild of all four list families (reCalculateAllGenomeLists)."""
    import torch
    from maple_b200.genome_list import pack_lists
    out = {"note": "wall clock with a device synchronize on both sides, one run each after one warm-up where it is cheap"}

    def timed(fn, warm=True):
        if warm:
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return r, time.perf_counter() - t0

    try:
        other = "fast" if args.round == "deep" else "deep"
        p2 = round_params(type("A", (), {"round": other})(), d.model.lRef)
        o, dt = timed(lambda: tree.spr_search(d_nodes, p2))
        rec = tree.search_records(o)
        out["search_round_" + other] = {"placements": int(rec["phase1"].sum()), "seconds": dt, "placements_per_s": float(rec["phase1"].sum()) / dt,
                                        "searched": int((rec["status"] == 0).sum()), "proposals": int((rec["placement"] >= 0).sum())}
        # the CPU port on the same stop rules, same bounded sample as the headline's cpu_baseline
        from oracle.oracle import Oracle
        import numpy as np
        orc = Oracle(d.model)
        cores = orc.use_all_cores()
        nodes_all = d_nodes.cpu().numpy()
        sample = cpu_sample(nodes_all, args.cpu_searches)
        host = tree.arena.to_host()
        ta, pd = oracle_tree(d, tree), params_dict(p2)
        orc.search_batch(ta, host, pd, sample[:64], lazy_mode=1)
        t0 = time.perf_counter()
        ref = orc.search_batch(ta, host, pd, sample, lazy_mode=1)
        cdt = time.perf_counter() - t0
        pos = {int(nd): i for i, nd in enumerate(nodes_all)}
        got = rec[[pos[int(nd)] for nd in sample]]
        same = all(np.array_equal(got[f], ref[f]) for f in ("placement", "bestNode", "status", "phase1", "bLenTop", "bLenBottom", "bLenAppend"))
        out["search_round_" + other]["cpu_baseline"] = {"value": float(ref["phase1"].sum()) / cdt, "unit": UNIT, "cores": cores, "kind": "port",
                                                        "sample": "%d of the %d searches (evenly spread)" % (len(sample), len(nodes_all)),
                                                        "gpu_matches_oracle_on_sample": bool(same)}
    except Exception as e:
        out["search_round_error"] = repr(e)[:300]
    try:
        lk, dt = timed(tree.tree_likelihood)
        out["tree_likelihood"] = {"logLK": lk, "seconds": dt, "internal_nodes": int((tree.child0 >= 0).sum())}
    except Exception as e:
        out["tree_likelihood_error"] = repr(e)[:300]
    try:
        dist0 = tree.dist.copy()
        (upd, dirty), dt = timed(lambda: tree.optimize_branch_lengths(1.0 / (10 * d.model.lRef)), warm=False)
        out["fast_branch_length_sweep"] = {"branches": int(tree.n - 3), "updated": int(upd), "seconds": dt,
                                           "mean_abs_change_in_mutations": float(abs(tree.dist - dist0).mean() * d.model.lRef)}
        lists_t0 = time.perf_counter()
        tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, d.model.usingErrorRate))
        torch.cuda.synchronize()
        out["recalculate_all_lists"] = {"nodes": int(tree.n), "seconds": time.perf_counter() - lists_t0,
                                        "note": "includes packing the tip lists on the host"}
        lk2 = tree.tree_likelihood()
        out["fast_branch_length_sweep"]["logLK_after_sweep_and_rebuild"] = lk2
    except Exception as e:
        out["branch_length_sweep_error"] = repr(e)[:300]
    try:  # the reference's default mode: one branch at a time, updatePartials after every change, all on the device (one lane)
        tree.dist[:] = dist0
        tree.d_dist.copy_(torch.from_numpy(tree.dist))
        tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, d.model.usingErrorRate))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        upd, dirty = tree.optimize_branch_lengths_sequential(1.0 / (10 * d.model.lRef))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["sequential_branch_length_sweep"] = {"branches": int(tree.n - 3), "updated": int(upd), "seconds": dt,
                                                 "logLK_after_sweep": tree.tree_likelihood(),
                                                 "note": "traverseTreeToOptimizeBranchLengths(fastPass=False) + updatePartials after every change, "
                                                         "device-resident, sequential by definition"}
        tree.dist[:] = dist0
        tree.d_dist.copy_(torch.from_numpy(tree.dist))
        tree.recalculate_all_lists(d.tip_nodes, pack_lists(d.tip_lists, d.model.lRef, d.model.usingErrorRate))
    except Exception as e:
        out["sequential_sweep_error"] = repr(e)[:300]
    return out


def placement_extra(args):
    """Side measurement, outside every timed region and in its own process (a failure there cannot take the headline down): the
    placement workload of this file (`--workload place`: maple_place_batch = findBestParentForNewSample for a batch of new samples
    on the same frozen tree) with a smaller batch; its JSON line is embedded as it is."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "place", "--nseq", str(args.nseq), "--new-samples", "20000",
                            "--steps", "2", "--warmup", "1", "--cpu-samples", "200"], capture_output=True, text=True, timeout=400, cwd=ROOT)
        got = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        return json.loads(got[-1]) if got else {"error": (r.stderr or r.stdout)[-400:]}
    except Exception as e:  # timeout included
        return {"error": repr(e)[:400]}


# ------------------------------------------------------------------------------------------------------------------
# --workload place: a batch of new samples placed on the frozen tree (findBestParentForNewSample, :7912; the reference's own
# batch form is process_chunk under joblib, :11190-11287)
PLACE_METRIC = "new_sample_candidate_placements_per_sec"


def place_workload_name(args):
    return ("synthetic 29903-bp, frozen tree of %d seqs (UNREST+rateVariation), %d new samples (tips with one extra substitution) placed "
            "with findBestParentForNewSample, non-strict stop rules" % (args.nseq, args.new_samples))


def new_samples(d, k, seed=3):
    """k new samples: existing tips with one extra substitution in their first long reference run (so that they are not simply
    absorbed as minor sequences)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = []
    for i in rng.choice(len(d.tip_lists), min(k, len(d.tip_lists)), replace=False):
        gl, new, pos, done = d.tip_lists[i], [], 0, False
        for e in gl:
            end = e[1] if e[0] in (4, 5) else pos + 1
            if not done and e[0] == 4 and end - pos > 40:
                mid = pos + 20
                ref = int(d.model.refIdx[mid])
                new += [(4, mid), ((ref + 1 + int(rng.integers(3))) % 4, ref), (4, end)]
                done = True
            else:
                new.append(e)
            pos = end
        out.append(new)
    return out


def place_params_dict(lRef):
    L = math.log(lRef)
    return {"strictStopRules": 0, "allowedFails": 5, "deeperSearchForLongBranches": 0, "onlyFindIdentical": 0,
            "thresholdLogLK": 18.0 * L, "thresholdLogLKoptimization": 1.0 * L, "thresholdLogLKconsecutivePlacement": 1.0,
            "effectivelyNon0BLen": 1.0 / (10 * lRef), "BLenThresholdDeeperSearch": (L + 5) / lRef, "oneMutBLen": 1.0 / lRef}


def run_place_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from maple_b200.genome_list import pack_lists
    d, orc, ta, host, nodes, setup_s = build_problem_cpu(args)
    cores = orc.use_all_cores()
    samples = new_samples(d, args.new_samples)
    sub = samples[:: max(1, len(samples) // args.cpu_samples)][: args.cpu_samples]
    packed = pack_lists(sub, d.model.lRef, d.model.usingErrorRate)
    pd = place_params_dict(d.model.lRef)
    orc.place_batch(ta, host, pd, pack_lists(sub[:16], d.model.lRef, d.model.usingErrorRate))
    t0 = time.perf_counter()
    tot = 0
    for _ in range(args.steps):
        rec = orc.place_batch(ta, host, pd, packed)
        tot += int(rec["phase1"].sum())
    dt = time.perf_counter() - t0
    val = tot / dt
    print(json.dumps({
        "impl": "reference", "metric": PLACE_METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": place_workload_name(args), "nseq": args.nseq, "samples_per_step": len(sub), "placements_per_step": tot // args.steps,
                   "samples_per_s": len(sub) * args.steps / dt, "setup_s": setup_s},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d of the %d new samples, oracle/maple_oracle.c placement (C, OpenMP on %d threads)" % (len(sub), len(samples), cores)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_place(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from maple_b200 import capi
    from maple_b200.genome_list import pack_lists
    from maple_b200.sharding import all_gather_raw
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: maple_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d, eng, tree, nodes, setup_s = build_problem(args, local)
    dev = eng.device
    samples = new_samples(d, args.new_samples)
    mine = samples[rank::world]  # strong scaling: the tree on every GPU, the samples dealt round-robin (the reference's joblib chunks, :11280)
    packed = pack_lists(mine, d.model.lRef, d.model.usingErrorRate)
    pp = capi.PlaceParams()
    for k, v in place_params_dict(d.model.lRef).items():
        setattr(pp, k, v)
    n_total = len(samples)

    def gather(out):
        if world > 1:
            all_gather_raw(out, n_total, world)

    ids, mark = tree.stage_samples(packed)
    for _ in range(args.warmup):
        out = tree.place_staged(ids, pp)
        gather(out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 5.0:
            time.sleep(0.05)
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = eng.launches
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        out = tree.place_staged(ids, pp)
        ev[k][1].record()
        gather(out)
        ev[k][2].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launches - l0
    step_ms = sum(a.elapsed_time(z) for a, _, z in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in ev) / args.steps
    rec = tree.place_records(out)
    tree.release_samples(mark)
    cand = torch.tensor([int(rec["phase1"].sum()), int((rec["status"] == 0).sum()), int((rec["status"] == 1).sum()), int((rec["status"] >= 2).sum())],
                        dtype=torch.int64, device=dev)
    tmax = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cand)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    cand_total, placed, absorbed, failed = (int(x) for x in cand.tolist())
    total_ms = float(tmax.item())
    value = cand_total * args.steps / (total_ms / 1e3)
    # ---- e2e: the packed sample lists come from host memory every step, the records go back to the host
    h_out = torch.empty((len(mine), 48), dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids2, mark2 = tree.stage_samples(packed)
        o = tree.place_staged(ids2, pp)
        gather(o)
        h_out.copy_(o, non_blocking=True)
        torch.cuda.synchronize()
        _ = int(h_out[0, 0])
        tree.release_samples(mark2)
    e2e_dt = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        dist.destroy_process_group()
        return
    A, n = tree.arena, tree.n
    tot_lists = slice(3 * n, 4 * n)
    have = A.key_start[tot_lists] >= 0
    mean_tot_bytes = float((A.nkeys[tot_lists][have].double() * 4 + A.npay[tot_lists][have].double() * 8).mean().item()) + 16
    sample_bytes = float(packed.nkeys.mean() * 4 + packed.npay.mean() * 8) + 16
    alg_bytes = cand_total / world * mean_tot_bytes + len(mine) * (sample_bytes + 48)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9
    h2d = int(packed.key.nbytes + packed.pay.nbytes + packed.key_start.nbytes * 2 + packed.nkeys.nbytes * 2)
    line = {
        "metric": PLACE_METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": place_workload_name(args), "nseq": args.nseq, "nodes": tree.n, "samples_per_step": n_total,
                   "placements_per_step": cand_total, "samples_per_s": n_total * args.steps / (total_ms / 1e3), "placed": placed,
                   "absorbed_as_minor": absorbed, "failed": failed, "setup_s": setup_s, "l2": "flushed between timed iterations (160 MB memset)",
                   "kernel_variant": capi.DEFAULT_PLACE_VARIANT,
                   "parallelism": "whole tree on every GPU; samples dealt round-robin to %d GPU(s); one NCCL all-gather of the 48-byte records" % world},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": cand_total / float(e2e_dt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(len(mine) * 48),
                "samples_per_s": n_total / float(e2e_dt.item()), "note": "packed sample lists from host memory, staged in the arena, records out, per rank"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "k_place_samples_warp_mat", "kernel_ms": kern_ms, "alg_bytes_per_launch": int(alg_bytes),
                     "mean_mid_branch_list_bytes": round(mean_tot_bytes, 1),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle.oracle import Oracle
        host = tree.arena.to_host()
        orc = Oracle(d.model)
        cores = orc.use_all_cores()
        step_ = max(1, len(samples) // args.cpu_samples)
        sub = samples[::step_][: args.cpu_samples]
        sp_ = pack_lists(sub, d.model.lRef, d.model.usingErrorRate)
        ta, pd = oracle_tree(d, tree), place_params_dict(d.model.lRef)
        orc.place_batch(ta, host, pd, pack_lists(sub[:16], d.model.lRef, d.model.usingErrorRate))
        t0 = time.perf_counter()
        ref = orc.place_batch(ta, host, pd, sp_)
        cdt = time.perf_counter() - t0
        got = rec[::step_][: args.cpu_samples]
        same = all(np.array_equal(got[f], ref[f]) for f in ("bestNode", "status", "phase1", "missedMinors", "bLenTop", "bLenBottom", "bLenAppend"))
        with np.errstate(invalid="ignore"):
            diff = np.nanmax(np.abs(got["bestScore"] - ref["bestScore"]))
        line["cpu_baseline"] = {"value": float(ref["phase1"].sum()) / cdt, "unit": UNIT, "cores": cores, "kind": "port",
                                "samples_per_s": len(sub) / cdt,
                                "sample": "%d of the %d new samples (evenly spread), %d candidate branches, oracle/maple_oracle.c placement "
                                          "(C, OpenMP)" % (len(sub), len(samples), int(ref["phase1"].sum())),
                                "gpu_matches_oracle_on_sample": bool(same), "max_abs_score_diff": float(diff)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.workload == "place":
        return run_place_reference(args) if args.impl == "reference" else run_place(args)
    if args.impl == "reference":
        return run_reference(args)
    import numpy as np
    import torch
    import torch.distributed as dist
    from maple_b200 import capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: maple_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d, eng, tree, nodes, setup_s = build_problem(args, local)
    dev = eng.device
    p = round_params(args, d.model.lRef)
    # strong scaling: every rank holds the whole tree.  First warm-up round: searches dealt round-robin like
    # coreNum[node]==corNum (:9619).  From then on the deal follows what the previous round measured (per-search SM cycles,
    # all-gathered once): cost-balanced shards, longest search first inside every shard (sharding.shard_nodes, DeviceTree.spr_search)
    # -- a run has tens of rounds on a tree that hardly changes, so "the previous round" always exists after the first.
    from maple_b200.sharding import all_gather_raw, shard_nodes, shard_positions
    mine = shard_nodes(nodes, rank, world)
    d_nodes = torch.as_tensor(mine, dtype=torch.int32, device=dev)

    def step(src_nodes):
        out = tree.spr_search(src_nodes, p)
        if world > 1:  # one collective per round: everybody gets every proposal
            all_gather_raw(out, len(nodes), world, per_rank)
        return out

    per_rank = (len(nodes) + world - 1) // world
    cost = None
    for w in range(max(args.warmup, 1)):
        out = step(d_nodes)
        if w == 0 and world > 1:  # re-deal by measured cost
            # two measures per search, all-gathered: the cycles it took (they order a shard and pick the searches that get an SM
            # each, DeviceTree.spr_search) and the candidates it scored -- the cycles of a search include waiting for the
            # lanes it shares its warp with, so shards of equal summed cycles are not shards of equal work; candidates are
            both = torch.zeros((per_rank, 2), dtype=torch.int64, device=dev)
            both[: d_nodes.numel(), 0] = tree.search_cost[d_nodes.long()]
            both[: d_nodes.numel(), 1] = torch.from_numpy(tree.search_records(out)["phase1"].astype(np.int64)).to(dev)
            allc = torch.empty((world * per_rank, 2), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allc, both)
            allc = allc.cpu().numpy().reshape(world, per_rank, 2)
            cyc_all = np.zeros(len(nodes), np.float64)
            cost = np.zeros(len(nodes), np.float64)
            for r in range(world):
                pos = shard_positions(len(nodes), r, world)
                cyc_all[pos] = allc[r, : len(pos), 0]
                cost[pos] = allc[r, : len(pos), 1] + 500.0  # + the search's own merges and branch lengths, in candidates' worth
            tree.search_cost[torch.as_tensor(nodes, dtype=torch.int64, device=dev)] = torch.as_tensor(cyc_all, dtype=torch.int64, device=dev)
            mine = shard_nodes(nodes, rank, world, cost)
            d_nodes = torch.as_tensor(mine, dtype=torch.int32, device=dev)
            from maple_b200.sharding import shard_sizes
            per_rank = int(shard_sizes(len(nodes), world, cost).max())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 5.0:
            time.sleep(0.05)
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)  # > 126 MB L2
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    l0 = eng.launches
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations
        ev[k][0].record()
        out = tree.spr_search(d_nodes, p)
        ev[k][1].record()
        if world > 1:
            all_gather_raw(out, len(nodes), world, per_rank)
        ev[k][2].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launches - l0
    step_ms = sum(a.elapsed_time(z) for a, _, z in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in ev) / args.steps
    rec = tree.search_records(out)
    cand = torch.tensor([int(rec["phase1"].sum()), int((rec["status"] == 0).sum()), int((rec["status"] == 3).sum()),
                         int((rec["placement"] >= 0).sum())], dtype=torch.int64, device=dev)
    tmax = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    kern_all = [kern_ms]
    if world > 1:
        dist.all_reduce(cand)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        kt = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(kt, torch.tensor([kern_ms], dtype=torch.float64, device=dev))
        kern_all = [round(float(x), 1) for x in kt.tolist()]
    cand_total, searched, overflowed, proposals = (int(x) for x in cand.tolist())
    total_ms = float(tmax.item())
    value = cand_total * args.steps / (total_ms / 1e3)

    # ---- e2e: node ids from pinned host memory, records back to the host, every step
    h_nodes = torch.empty(len(mine), dtype=torch.int32, pin_memory=True)
    h_nodes.copy_(torch.from_numpy(np.ascontiguousarray(mine)))
    h_out = torch.empty((len(mine), 64), dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dn = h_nodes.to(dev, non_blocking=True)
        o = step(dn)
        h_out.copy_(o, non_blocking=True)
        torch.cuda.synchronize()
        _ = int(h_out[0, 0])
    e2e_dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_val = cand_total / float(e2e_dt.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        dist.destroy_process_group()
        return

    # algorithmic bytes of one round: every scored candidate reads the stored mid-branch list of its branch and
    # writes nothing; every search reads its removed list once and writes one 64-byte record
    A, n = tree.arena, tree.n
    tot_lists = slice(3 * n, 4 * n)
    have = A.key_start[tot_lists] >= 0
    mean_tot_bytes = float((A.nkeys[tot_lists][have].double() * 4 + A.npay[tot_lists][have].double() * 8).mean().item()) + 16
    low = slice(0, n)
    mean_low_bytes = float((A.nkeys[low].double() * 4 + A.npay[low].double() * 8).mean().item()) + 16
    alg_bytes = cand_total / world * mean_tot_bytes + (searched / world) * (mean_low_bytes + 64)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9
    traffic = None  # ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, one launch (profiles/)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "search_traffic.json")))
        if tr.get("nseq") == args.nseq and tr.get("round") == args.round and world == 1:
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "nseq": args.nseq, "round": args.round, "nodes": tree.n,
                   "searches_per_step": searched, "placements_per_step": cand_total, "proposals": proposals,
                   "scratch_overflows": overflowed, "arena_MB": round(tree.arena.used_bytes() / 1e6, 1), "setup_s": setup_s,
                   "critical_searches": int(getattr(tree, "_critical_set", 0)),  # searches given an SM each (rank 0's choice)
                   "l2": "flushed between timed iterations (160 MB memset)",
                   "parallelism": "whole tree on every GPU; searches dealt to %d GPU(s) in cost-balanced shards (the previous round's per-search "
                                  "cycles), longest first inside a shard; one NCCL all-gather of the 64-byte result records per round" % world},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(len(mine) * 4), "d2h_bytes_per_step": int(len(mine) * 64),
                "note": "pinned host node ids in, result records out, per rank"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "k_spr_search_fsm<7,true,false>", "kernel_ms": kern_ms, "kernel_ms_per_rank": kern_all, "alg_bytes_per_launch": int(alg_bytes),
                     "mean_mid_branch_list_bytes": round(mean_tot_bytes, 1),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle.oracle import Oracle
        host = tree.arena.to_host()
        sample = cpu_sample(nodes, args.cpu_searches)
        orc = Oracle(d.model)
        orc.use_all_cores()
        ta, pd = oracle_tree(d, tree), params_dict(p)
        orc.search_batch(ta, host, pd, sample[: max(16, len(sample) // 8)], lazy_mode=1)
        t0 = time.perf_counter()
        ref = orc.search_batch(ta, host, pd, sample, lazy_mode=1)
        cdt = time.perf_counter() - t0
        pos = {int(nd): i for i, nd in enumerate(nodes)}
        got = rec[[pos[int(nd)] for nd in sample]]
        same = all(np.array_equal(got[f], ref[f]) for f in ("placement", "bestNode", "status", "phase1", "bLenTop", "bLenBottom", "bLenAppend"))
        fin = np.isfinite(ref["bestScore"])
        line["cpu_baseline"] = {"value": float(ref["phase1"].sum()) / cdt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                "sample": "%d of the %d searches of the round (evenly spread), %d placements, oracle/maple_oracle.c "
                                          "search (C, OpenMP)" % (len(sample), len(nodes), int(ref["phase1"].sum())),
                                "gpu_matches_oracle_on_sample": bool(same),
                                "max_abs_score_diff": float(np.max(np.abs(got["bestScore"][fin] - ref["bestScore"][fin]), initial=0.0))}
    if world == 1 and not args.no_extras and args.nseq >= 20000:
        line["extra"] = callers_extra(args, d, eng, tree, d_nodes)
        line["extra"]["placement_batch"] = placement_extra(args)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
