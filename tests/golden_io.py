"""Load the committed golden fixtures (tests/golden/*.json.gz, written by make_golden.py)."""
import functools
import glob
import gzip
import json
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-len(".json.gz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json.gz")))


def hw_names():
    """Fixtures the device tests run on (all of them; kept as a function because the tests parametrize over it)."""
    return golden_names()


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
