"""Load the committed golden fixtures (tests/golden/*.json.gz, written by make_golden.py)."""
import functools
import glob
import gzip
import json
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-len(".json.gz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json.gz")))


# fixtures added after the last hardware run of round 1: their device tests live in tests/test_zz*_gpu_*.py (collected last), so that
# a surprise there cannot stop the device tests that have already run on a B200
NOT_YET_ON_HARDWARE = ("ay_unrest_1000",)


def hw_names():
    return [n for n in golden_names() if n not in NOT_YET_ON_HARDWARE]


def _as_tuples(gl):
    return [tuple(e) for e in gl]


@functools.lru_cache(maxsize=None)
def load_golden(name):
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".json.gz"), "rt") as f:
        fx = json.load(f)
    fx["lists"] = [_as_tuples(gl) for gl in fx["lists"]]
    return fx


@functools.lru_cache(maxsize=None)
def load_extras(name):
    """tests/golden/extras/<name>.json.gz: Newick strings, re-read trees, the input-tree set-up and the branch-length sweeps
    recorded from the reference on the same frozen tree (make_golden.py: harvest_extras)."""
    with gzip.open(os.path.join(GOLDEN_DIR, "extras", name + ".json.gz"), "rt") as f:
        fx = json.load(f)
    fx["lists"] = [_as_tuples(gl) for gl in fx["lists"]]
    return fx
