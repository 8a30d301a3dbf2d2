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


def test_c_oracle_matches_the_reference_on_the_splitter_example(golden_dir):
    """tests/golden/example_splitter.npz holds what the REFERENCE's own measure_fields returned on the straight guide of
    examples/simulate_splitter_fdtd.py (320 x 180, 4500 steps): the C restatement reproduces the series bit for bit."""
    import importlib.util
    from oracle.fdtd_c import OracleFDTDC
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("simulate_splitter_fdtd", os.path.join(root, "examples", "simulate_splitter_fdtd.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    gold = np.load(os.path.join(golden_dir, "example_splitter.npz"))
    Nx, Ny, steps, npml = int(gold["Nx"]), int(gold["Ny"]), int(gold["steps"]), int(gold["npml"])
    eps_wg, _, J_in, J_wg, _ = ex.geometry(Nx, Ny, npml)
    O = OracleFDTDC(eps_wg, float(gold["dL"]), [npml, npml, 0])
    assert O.dt == float(gold["dt"])
    wave = ex.pulse(steps, O.dt, int(gold["t0"]), int(gold["sigma"]))
    series, _ = O.run(steps, [("z", J_in, wave)], [("Ez", J_wg)])
    assert np.array_equal(series, gold["measured_wg"])
    T, f_max = ex.transmission(gold["measured"], gold["measured_wg"], O.dt)
    assert np.array_equal(T, gold["T"]) and f_max == float(gold["f_max"])
    assert 0.9 < T.sum() < 1.0 and abs(T[0] - T[1]) < 0.01 * T[0]          # a working 50 / 50 splitter
    assert abs(f_max - 299792458.0 / 2e-6) < 0.03 * 299792458.0 / 2e-6
