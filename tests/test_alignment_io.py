"""Input producers of the path against the reference: tip genome lists (probVectTerminalNode, :3882) recorded from the
reference under each fixture's flags, and the MAPLE-format reader (readConciseAlignment, :3498) through a write/read round
trip of the recorded differences."""
import gzip
import os

import numpy as np
import pytest

from golden_io import golden_names, load_golden
from maple_b200.alignment import AlignmentError, read_maple_alignment, tip_genome_list
from maple_b200.genome_list import lists_equal
from maple_b200.model import MapleModel


@pytest.mark.parametrize("name", [n for n in golden_names() if "tipInputs" in load_golden(n)])
def test_tip_lists_match_reference(name):
    g = load_golden(name)
    ti = g["tipInputs"]
    model = MapleModel.from_reference_snapshot(g["env"], g["model"])
    er = model.errorRates if (ti["usingErrorRate"] and ti["errorRateSiteSpecific"]) else None
    n_amb = 0
    for t in ti["tips"]:
        diffs = [tuple(d) for d in t["diffs"]]
        got = tip_genome_list(diffs, model.refIdx, model.lRef, ti["usingErrorRate"], ti["errorRateGlobal"] or 0.0, er, ti["onlyNambiguities"])
        ref = g["lists"][t["list"]]
        if ti["usingErrorRate"] and ti["errorRateSiteSpecific"]:
            # The reference takes the ambiguity vector from a table that earlier in-place updates left at whatever site-specific
            # error rate they last saw (see alignment.py); ours uses the rate of the site itself.  Same entries, same support,
            # values within the size of an error rate.
            assert len(got) == len(ref)
            for a, b in zip(got, ref):
                assert list(a[:2]) == list(b[:2]), t["name"]
                if a[0] == 6:
                    assert [x > 0.2 for x in a[-1]] == [x > 0.2 for x in b[-1]] and max(abs(x - y) for x, y in zip(a[-1], b[-1])) < 0.05
        else:
            assert lists_equal(got, ref), t["name"]
        n_amb += sum(1 for d in diffs if d[0] not in "acgtn-")
    assert len(ti["tips"]) >= 50
    if name.startswith("ex_"):
        assert n_amb > 0  # the example alignment carries IUPAC codes: the O-entry path is exercised


def _write(path, ref, data, with_ref=True):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "wt") as f:
        if with_ref:
            f.write(">reference\n")
            for i in range(0, len(ref), 70):
                f.write(ref[i:i + 70].upper() + "\n")
        for name, diffs in data.items():
            f.write(">" + name + "\n")
            for d in diffs:
                f.write("\t".join(str(x) for x in d) + "\n")


@pytest.mark.parametrize("ext", [".txt", ".gz"])
def test_reader_round_trip(tmp_path, ext):
    g = load_golden("ex_unrest")
    ref = g["env"]["ref"]
    data = {t["name"]: [tuple(d) for d in t["diffs"]] for t in g["tipInputs"]["tips"]}
    p = str(tmp_path / ("aln" + ext))
    _write(p, ref, data)
    ref2, data2 = read_maple_alignment(p)
    assert ref2 == ref.lower() and data2 == data and list(data2) == list(data)
    p2 = str(tmp_path / ("noref" + ext))
    _write(p2, ref, data, with_ref=False)
    _, data3 = read_maple_alignment(p2, reference=ref)
    assert data3 == data


def test_reader_rejects_what_the_reference_rejects(tmp_path):
    ref = "acgtacgtac"
    for body, msg in ((">s\nc\t2\n", "reference nucleotide"), (">s\nt\t3\nt\t3\n", "overlaps"), (">s\nn\t2\t4\na\t4\n", "overlaps"),
                      (">s\nt\n", "one column")):
        p = str(tmp_path / "bad.txt")
        with open(p, "w") as f:
            f.write(">ref\n" + ref + "\n" + body)
        with pytest.raises(AlignmentError, match=msg):
            read_maple_alignment(p)
    assert tip_genome_list(None, np.zeros(10, np.int8), 10) == [(5, 10)]
