"""The C restatement of the reference step (oracle/fdtd_c.c, OpenMP) against the numpy oracle: bit for bit, on every
small field case (PML on all / some / no axes, 2-D, 1-D, Nx = 1), and against the golden vectors produced by the
reference itself."""
import os

import numpy as np
import pytest

from oracle import cases
from oracle.fdtd_c import OracleFDTDC
from oracle.fdtd_numpy import FIELD_KEYS, OracleFDTD


@pytest.mark.parametrize("name", cases.SMALL_FIELD_CASES)
def test_c_oracle_equals_numpy_oracle_bitwise(name):
    case = cases.field_case(name)
    A = OracleFDTD(case["eps"], case["dL"], case["npml"])
    B = OracleFDTDC(case["eps"], case["dL"], case["npml"])
    sa, _ = A.run(case["steps"], case["sources"], case["probes"])
    sb, _ = B.run(case["steps"], case["sources"], case["probes"])
    assert np.array_equal(sa, sb)
    for k in FIELD_KEYS:
        assert np.array_equal(A.fields()[k], B.fields()[k]), k
    for fam in ("ICE", "IH", "ICH", "ID"):
        for c in range(3):
            assert np.array_equal(getattr(A, fam)[c], getattr(B, fam)[c]), (fam, c)


def test_c_oracle_against_reference_golden(golden_dir):
    case = cases.field_case("c1_tm")
    gold = np.load(os.path.join(golden_dir, "fields_c1_tm.npz"))
    O = OracleFDTDC(case["eps"], case["dL"], case["npml"])
    series, snaps = O.run(case["steps"], case["sources"], case["probes"], case["snapshots"])
    assert np.array_equal(series, gold["series"])
    s = int(gold["stride"])
    for t in case["snapshots"]:
        for k in FIELD_KEYS:
            assert np.array_equal(snaps[t][k][::s, ::s, :], gold["t%d_%s" % (t, k)]), (t, k)


def test_c_oracle_reproduces_scaled_config2_prefix(golden_dir):
    """The 96^3 / 2000-step fixture of config 2 (its first 150 steps come from the reference itself): the first 200
    steps recomputed here, bit for bit."""
    case = cases.scaled_case("c2_96")
    gold = np.load(os.path.join(golden_dir, "fields_c2_96.npz"))
    n = 200
    O = OracleFDTDC(case["eps"], case["dL"], case["npml"])
    series, _ = O.run(n, [(c, p, w[:n]) for c, p, w in case["sources"]], case["probes"])
    assert np.array_equal(series, gold["series"][:n])
    assert int(gold["reference_steps"]) == 150
