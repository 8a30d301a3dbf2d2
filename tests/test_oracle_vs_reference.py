"""The oracle against the reference's own code, where /root/reference exists (build container).
Skipped on the GPU box -- there tests/golden/ carries the pin."""
import numpy as np
import pytest
import torch

from oracle import cases, ref_loader
from oracle.fdtd_numpy import FIELD_KEYS, OracleFDTD, pad_to_3d, sigma_profiles, time_step
from oracle.fdtd_torch import TorchFDTD

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")


@pytest.mark.parametrize("name", cases.SMALL_FIELD_CASES)
def test_step_bit_identical(name):
    ref = ref_loader.load()
    case = cases.field_case(name)
    F = ref.fdtd(case["eps"], case["dL"], case["npml"])
    O = OracleFDTD(case["eps"], case["dL"], case["npml"])
    T = TorchFDTD(case["eps"], case["dL"], case["npml"])
    assert F.dt == O.dt == T.dt
    for t in range(min(case["steps"], 60)):
        J = {"x": None, "y": None, "z": None}
        for comp, profile, wave in case["sources"]:
            term = pad_to_3d(profile) * wave[t]
            J[comp] = term if J[comp] is None else J[comp] + term
        f = F.forward(Jx=J["x"], Jy=J["y"], Jz=J["z"])
        g = O.step(Jx=J["x"], Jy=J["y"], Jz=J["z"])
        h = T.step(**{"J" + c: (None if J[c] is None else torch.as_tensor(J[c])) for c in "xyz"})
    for k in FIELD_KEYS:
        assert np.array_equal(f[k], g[k]), k
        assert np.array_equal(f[k], h[k].numpy()), k
    for c, n in enumerate("xyz"):
        for nm in ("ICE", "IH", "ICH", "ID"):
            assert np.array_equal(getattr(F, nm + n), getattr(O, nm)[c])


def test_sigma_profiles_and_quirks():
    ref = ref_loader.load()
    shape, npml = (12, 7, 9), [3, 2, 0]
    F = ref.fdtd(np.ones(shape) * 2.0, 5e-8, npml)
    dt = time_step(5e-8)
    assert dt == F.dt
    sH, sD = sigma_profiles(shape, npml, dt)
    assert np.array_equal(F.sigHx[:, 0, 0], sH[0]) and np.array_equal(F.sigDx[:, 0, 0], sD[0])
    assert np.array_equal(F.sigHy[0, :, 0], sH[1]) and np.array_equal(F.sigDy[0, :, 0], sD[1])
    assert np.array_equal(F.sigHz[0, 0, :], sH[2]) and np.array_equal(F.sigDz[0, 0, :], sD[2])
    # SURVEY appendix A verified example N=12, p=3
    s0 = 0.5 * 8.85418782e-12 / dt
    np.testing.assert_allclose(sH[0] / s0, [0, 8 / 27, 1 / 27, 0, 0, 0, 0, 0, 0, 1 / 216, 1 / 8, 125 / 216], atol=1e-15)
    np.testing.assert_allclose(sD[0] / s0, [0, 125 / 216, 1 / 8, 1 / 216, 0, 0, 0, 0, 0, 0, 1 / 27, 8 / 27], atol=1e-15)
    O = OracleFDTD(np.ones(shape) * 2.0, 5e-8, npml)
    for c, n in enumerate("xyz"):
        for q in range(4):
            assert np.array_equal(getattr(F, "mH%s%d" % (n, q + 1)) * np.ones(shape), O.mH[c][q])
            assert np.array_equal(getattr(F, "mD%s%d" % (n, q + 1)) * np.ones(shape), O.mD[c][q])


def test_reference_rejects_complex_J():
    """`self.Dx += Jx` (fdtd.py:125-127) is an in-place add into a float64 array: numpy raises a TypeError
    (UFuncTypeError) for a complex J.  ceviche_b200.fdtd._as_J mirrors that (tests/test_gpu_fields.py)."""
    ref = ref_loader.load()
    F = ref.fdtd(np.ones((6, 5, 1)), 5e-8, [1, 1, 0])
    J = np.zeros((6, 5, 1), dtype=complex)
    J[3, 2, 0] = 1 + 2j
    with pytest.raises(TypeError):
        F.forward(Jz=J)



def test_sigma_profiles_sweep_including_overlapping_and_oversized_pml():
    """Every (N, npml) with N = 1..11 and npml = 0..8 on one axis: the host layer's 1-D profiles equal the reference's
    sigma arrays bit for bit where the reference builds them (also where the low and high PML ramps overlap and overwrite
    each other), and raise the same exception type where the reference's indices run out of its doubled grid."""
    import itertools
    from ceviche_b200.fdtd import sigma_profiles as host_profiles
    ref = ref_loader.load()
    dt = time_step(5e-8)
    agree = errors = 0
    for N, p in itertools.product(range(1, 12), range(0, 9)):
        shape, npml = (N, 3, 2), [p, 0, 0]
        try:
            F = ref.fdtd(np.ones(shape), 5e-8, npml)
            want, want_err = (F.sigHx[:, 0, 0].copy(), F.sigDx[:, 0, 0].copy()), None
        except Exception as e:          # noqa: BLE001 (whatever numpy raises is the reference's error convention)
            want, want_err = None, type(e)
        for fn in (host_profiles, sigma_profiles):           # the product's host layer and the oracle's restatement
            try:
                sH, sD = fn(shape, npml, dt)
                got, got_err = (sH[0], sD[0]), None
            except Exception as e:      # noqa: BLE001
                got, got_err = None, type(e)
            assert got_err == want_err, (N, p, fn.__module__, got_err, want_err)
            if want is not None:
                assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (N, p, fn.__module__)
        agree += want is not None
        errors += want is None
    assert agree > 60 and errors > 0
