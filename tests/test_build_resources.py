"""The default search kernel must come out of ptxas as measured: 128 registers (8 CTAs = 16 warps per SM) and well under
1 kB of spill stores.  The same source with a different set of instantiations has produced 2 kB of spills and a 40 % slower
kernel, silently -- this test makes that loud.  No GPU needed (reads the ptxas report of the in-tree build)."""
import os

import pytest

from maple_b200.build import LIB, PTXAS_LOG, build_extension, kernel_resources


def test_default_search_kernel_register_allocation():
    build_extension()
    if not os.path.isfile(PTXAS_LOG) or os.path.getmtime(PTXAS_LOG) + 5 < os.path.getmtime(LIB):
        pytest.skip("no ptxas report for the current library (built elsewhere)")
    res = kernel_resources()
    # <7, true, false>: the default (second form of the subtree scans, scan service and dense pass compiled out);
    # <7, false, false>: the first form, kept for A/B runs
    default = [v for k, v in res.items() if "k_spr_search_fsmILi7ELb1ELb0E" in k]
    assert len(default) == 1, sorted(res)
    r = default[0]
    assert r["registers"] == 128, r
    assert r["spill_stores"] <= 1200 and r["spill_loads"] <= 1200, r
    first_form = [v for k, v in res.items() if "k_spr_search_fsmILi7ELb0ELb0E" in k]
    assert len(first_form) == 1 and first_form[0]["registers"] == 128 and first_form[0]["spill_stores"] <= 1200, first_form
    wide = [v for k, v in res.items() if "k_spr_search_fsmILi6ELb1ELb0E" in k]
    assert wide and wide[0]["registers"] > 128
