"""The numpy oracle against the committed vectors (generated from the reference itself by
oracle/make_golden.py).  CPU only; this is what pins the oracle on a box without /root/reference."""
import os

import numpy as np
import pytest

from oracle import cases
from oracle.fdtd_numpy import FIELD_KEYS, OracleFDTD


@pytest.mark.parametrize("name", cases.FIELD_CASES)
def test_oracle_matches_golden_fields(name, golden_dir):
    case = cases.field_case(name)
    gold = np.load(os.path.join(golden_dir, "fields_%s.npz" % name))
    sim = OracleFDTD(case["eps"], case["dL"], case["npml"])
    assert sim.dt == float(gold["dt"])
    series, snaps = sim.run(case["steps"], case["sources"], case["probes"], case["snapshots"])
    # bit-for-bit: the oracle restates the reference's numpy op sequence exactly
    assert np.array_equal(series, gold["series"])
    s = int(gold["stride"])
    for t in case["snapshots"]:
        for k in FIELD_KEYS:
            assert np.array_equal(snaps[t][k][::s, ::s, :], gold["t%d_%s" % (t, k)]), (t, k)
            assert np.linalg.norm(snaps[t][k]) == float(gold["t%d_%s_norm" % (t, k)])


def test_zero_fields_stay_zero_in_2d_tm(golden_dir):
    """A Jz-driven Nz=1 run excites only Ez, Hx, Hy: the others are exactly 0."""
    gold = np.load(os.path.join(golden_dir, "fields_c1_tm.npz"))
    for k in ("Ex", "Ey", "Hz", "Dx", "Dy"):
        assert float(gold["t1000_%s_norm" % k]) == 0.0
    assert float(gold["t1000_Ez_norm"]) > 0.0
