#!/usr/bin/env python3
"""Time the UNMODIFIED reference (MAPLEv0.7.5.4.py under whatever interpreter runs this script) on an alignment written by the
bench's generator, counting SPR candidate placements the way bench.py does: phase-1 appendProbNode calls made from
findBestParentTopology (lines 7011 / 7223), and the initial-placement ones (8033 / 8050), against the time the reference itself
reports for the search.  Build-container only (needs /root/reference); the result goes to BASELINE.md / profiles/.

usage: time_reference.py NSEQ [extra reference options ...]      e.g.  time_reference.py 10000 --numTopologyImprovements 1
"""
import contextlib
import io
import json
import os
import platform
import re
import runpy
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/MAPLEv0.7.5.4.py"


def main():
    nseq = int(sys.argv[1])
    extra = sys.argv[2:]
    from maple_b200.synthetic import generate, write_maple_file
    d = generate(nseq, rate_variation="--rateVariation" in extra, seed=1, ml_like_blens=True)
    inp = "/tmp/ref_time_%d.maple" % nseq
    write_maple_file(d, inp)
    counts = {"spr": 0, "place": 0, "all": 0}
    timers = {"spr_search_s": 0.0, "spr_searches": 0}
    # The reference defines its functions at module level and calls them through its globals, so they are wrapped once they
    # exist: a profile hook that fires at the first call of a reference function after the definitions, installs the counting
    # wrappers into the module's globals and removes itself (no per-call tracing afterwards).
    argv = sys.argv
    sys.argv = [REF, "--input", inp, "--output", "/tmp/ref_time_%d_out" % nseq, "--overwrite", "--model", "UNREST"] + extra
    buf = io.StringIO()
    t0 = time.time()
    installed = {"done": False}
    orig_time = time.time

    def install(frame_globals):
        o_app = frame_globals["appendProbNode"]

        def appendProbNode(*a, **k):
            ln = sys._getframe(1).f_lineno
            counts["all"] += 1
            if ln in (7011, 7223):
                counts["spr"] += 1
            elif ln in (8033, 8050):
                counts["place"] += 1
            return o_app(*a, **k)

        o_fb = frame_globals["findBestParentTopology"]

        def findBestParentTopology(*a, **k):
            t = orig_time()
            r = o_fb(*a, **k)
            timers["spr_search_s"] += orig_time() - t
            timers["spr_searches"] += 1
            return r

        frame_globals["appendProbNode"] = appendProbNode
        frame_globals["findBestParentTopology"] = findBestParentTopology

    def profiler(frame, event, arg):  # fires once: at the first call of a reference function after its definitions exist
        g = frame.f_globals
        if (event == "call" and not installed["done"] and str(g.get("__file__", "")).endswith("MAPLEv0.7.5.4.py")
                and "findBestParentTopology" in g and "appendProbNode" in g and "startTopologyUpdatesParallel" in g):
            installed["done"] = True
            install(frame.f_globals)
            sys.setprofile(None)

    sys.setprofile(profiler)
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(REF, run_name="__main__")
    except SystemExit:
        pass
    finally:
        sys.setprofile(None)
        sys.argv = argv
    wall = time.time() - t0
    out = buf.getvalue()
    m = re.search(r"[Tt]ime.*finding.*?([0-9.]+)", out)
    res = {"nseq": nseq, "options": ["--model", "UNREST"] + extra, "interpreter": "%s %s" % (platform.python_implementation(), platform.python_version()),
           "cores": 1 if "--numCores" not in extra else int(extra[extra.index("--numCores") + 1]), "wall_s": wall,
           "spr_candidate_placements": counts["spr"], "initial_placement_candidates": counts["place"], "appendProbNode_calls": counts["all"],
           "spr_searches": timers["spr_searches"], "time_in_findBestParentTopology_s": timers["spr_search_s"],
           "spr_placements_per_s": counts["spr"] / timers["spr_search_s"] if timers["spr_search_s"] else None,
           "reference_reported_lines": [ln for ln in out.splitlines() if "ime" in ln and ("parent" in ln.lower() or "topolog" in ln.lower())][-12:]}
    print(json.dumps(res, indent=1))
    with open(os.path.join(ROOT, "profiles", "r02_reference_cpython_%d.json" % nseq), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
