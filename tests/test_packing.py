"""Packed genome-list format: lossless round trip of every list in the golden fixtures, and
bit-exact reconstruction of the reference's cumulative tables."""
import numpy as np
import pytest

from golden_io import golden_names, load_golden
from maple_b200.genome_list import lists_equal, pack_lists, split_key
from maple_b200.model import MapleModel


@pytest.mark.parametrize("name", golden_names())
def test_round_trip(name):
    g = load_golden(name)
    pl = pack_lists(g["lists"] + [None], g["env"]["lRef"], g["env"]["usingErrorRate"])
    assert pl.key_start[-1] == -1 and pl.get(len(pl) - 1) is None
    for i, gl in enumerate(g["lists"]):
        assert lists_equal(pl.get(i), gl)
        assert pl.key_start[i] % 4 == 0 and pl.pay_start[i] % 2 == 0
        assert split_key(pl.key[pl.key_start[i] + pl.nkeys[i] - 1])[4] == g["env"]["lRef"]


def test_rejects_bad_lists():
    with pytest.raises(ValueError):
        pack_lists([[(4, 50)]], 100, 0)  # does not reach lRef
    with pytest.raises(ValueError):
        pack_lists([[(0, 1, 0.5), (4, 100)]], 100, 1)  # 3-field nucleotide entry under the error model
    with pytest.raises(ValueError):
        pack_lists([[(4, 1 << 24)]], 1 << 24, 0)


def test_cumulative_tables_match_python_running_sum():
    g = load_golden("ex_unrest_rv_sse")
    m = MapleModel.from_reference_snapshot(g["env"], g["model"])
    Q, sr, er = g["model"]["mutMatrixGlobal"], g["model"]["siteRates"], g["model"]["errorRates"]
    cr, ce = [0.0], [0.0]
    for i in range(m.lRef):
        r = int(m.refIdx[i])
        cr.append(cr[-1] + Q[r][r] * sr[i])
        ce.append(ce[-1] + er[i])
    assert np.array_equal(m.cumulativeRate, np.array(cr))
    assert np.array_equal(m.cumulativeErrorRate, np.array(ce))
    assert m.totError == g["model"]["totError"]
