"""Helper mirrors of ceviche/utils.py (Yee-grid averaging, shape / value helpers, finite-difference checkers, plotting
orientation) against the reference's own functions loaded unmodified (oracle/ref_loader.py) where its sources are present,
and against closed forms everywhere."""
import sys

import numpy as np
import pytest
import torch

from oracle import ref_loader
from ceviche_b200 import utils


def _ref_utils():
    ref_loader.load()
    return sys.modules["ceviche.utils"]


needs_reference = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")


def test_grid_averaging_closed_form_numpy_and_torch():
    rng = np.random.default_rng(0)
    q = rng.random((5, 4, 3))
    xx, yy, zz = utils.grid_center_to_xyz(q)
    assert np.array_equal(xx, (q + np.roll(q, 1, 0)) / 2)
    assert np.array_equal(yy[:, 0], (q[:, 0] + q[:, -1]) / 2)            # periodic wrap
    assert np.array_equal(zz, (q + np.roll(q, 1, 2)) / 2)
    txx, tyy, tzz = utils.grid_center_to_xyz(torch.as_tensor(q))
    assert torch.is_tensor(txx) and np.array_equal(txx.numpy(), xx) and np.array_equal(tyy.numpy(), yy) and np.array_equal(tzz.numpy(), zz)
    a, b, c = utils.grid_center_to_xyz(q, averaging=False)
    assert np.array_equal(a, q) and a is not q and b is not c
    cx, cy, cz = utils.grid_xyz_to_center(xx, yy, zz)
    assert np.array_equal(cx, (xx + np.roll(xx, 1, 0)) / 2) and np.array_equal(cz, (zz + np.roll(zz, 1, 2)) / 2)
    ints = np.arange(24).reshape(2, 3, 4)
    assert utils.grid_xyz_to_center(ints, ints, ints)[1].dtype == np.float64
    vx, vy = utils.vec_zz_to_xy({'shape': (5, 4)}, q[:, :, 0].flatten())
    assert np.array_equal(vx, ((q[:, :, :1] + np.roll(q[:, :, :1], 1, 0)) / 2).flatten()) and vy.shape == (20,)
    # the eps_r setter of the fdtd object is this very average (fdtd.py:63-72): same bits as the oracle's restatement
    from oracle.fdtd_numpy import OracleFDTD
    O = OracleFDTD(1 + q, 5e-8, [1, 1, 0])
    for mine, theirs in zip(utils.grid_center_to_xyz(1 + q), O.eps_yee):
        assert np.array_equal(mine, theirs)


@needs_reference
def test_helpers_equal_the_reference_functions():
    ref = _ref_utils()
    rng = np.random.default_rng(1)
    q = rng.random((6, 5, 4))
    for mine, theirs in zip(utils.grid_center_to_xyz(q), ref.grid_center_to_xyz(q)):
        assert np.array_equal(mine, theirs)
    for mine, theirs in zip(utils.grid_center_to_xyz(q, averaging=False), ref.grid_center_to_xyz(q, averaging=False)):
        assert np.array_equal(mine, theirs)
    three = [rng.random((6, 5, 4)) for _ in range(3)]
    for mine, theirs in zip(utils.grid_xyz_to_center(*three), ref.grid_xyz_to_center(*three)):
        assert np.array_equal(mine, theirs)
    v = rng.random(30)
    for mine, theirs in zip(utils.vec_zz_to_xy({'shape': (6, 5)}, v), ref.vec_zz_to_xy({'shape': (6, 5)}, v)):
        assert np.array_equal(mine, theirs)
    for x in (3.0, 2, np.arange(3.0)):
        assert np.array_equal(utils.float_2_array(x), ref.float_2_array(x))
    for x in (3.0, 2, (1, 2, 3), [1, 2]):
        assert utils.get_shape(x) == ref.get_shape(x)
    # (ref.imarr / ref.jac_num go through get_value, which needs the real HIPS autograd's ArrayBox: closed forms instead)
    assert np.array_equal(utils.imarr(q), np.flipud(q[:, :, 0].T)) and np.array_equal(utils.imarr(q[:, :, 0]), np.flipud(q[:, :, 0].T))
    fn = lambda a: np.sum(np.sin(a) * np.arange(1, a.size + 1))
    x = rng.random(5)
    assert np.allclose(utils.der_num(fn, x, 2, 1e-6), ref.der_num(fn, x, 2, 1e-6), rtol=0, atol=0)
    # grad_num: the reference adds the derivative along the imaginary axis even for real arguments, which numpy refuses for a
    # float64 array (`arg_i_for[index] += 1j * delta / 2` raises); the mirror returns the real gradient there
    g = utils.grad_num(fn, x)
    assert g.dtype == np.complex128 and np.allclose(g.real, np.cos(x) * np.arange(1, 6), atol=1e-8) and np.all(g.imag == 0)
    vec = lambda a: np.array([np.sum(a ** 2), a[0] * a[1], np.sin(a[2])])
    exact = np.array([2 * x, [x[1], x[0], 0, 0, 0], [0, 0, np.cos(x[2]), 0, 0]]).T               # (n_in, n_out)
    assert np.allclose(utils.jac_num(vec, x), exact, atol=1e-6)
    assert np.array_equal(utils.reshape_to_ND(q[:, :, 0], 3), ref.reshape_to_ND(q[:, :, 0], 3))
    with pytest.raises(ValueError):
        utils.reshape_to_ND(np.zeros((1, 1, 1, 1)), 3)


def test_value_and_shape_helpers_on_tensors():
    import torch.autograd.forward_ad as fwAD
    x = torch.arange(3.0, dtype=torch.float64, requires_grad=True)
    y = utils.get_value(x * 2)
    assert not y.requires_grad and torch.equal(y, torch.tensor([0.0, 2.0, 4.0], dtype=torch.float64))
    with fwAD.dual_level():
        d = fwAD.make_dual(torch.ones(2), torch.full((2,), 3.0))
        v = utils.get_value(d * 2)
        assert fwAD.unpack_dual(v).tangent is None and torch.equal(v, torch.full((2,), 2.0))
    assert utils.get_value(4.5) == 4.5
    assert utils.get_shape(torch.zeros(2, 3)) == (2, 3) and utils.get_shape(np.zeros((4,))) == (4,) and utils.get_shape(1.5) == (1,)
    assert torch.is_tensor(utils.float_2_array(torch.zeros(2))) and utils.float_2_array(2.0).shape == (1,)
    # finite-difference checkers on tensors
    fn = lambda a: (a ** 3).sum()
    t = torch.tensor([1.0, 2.0], dtype=torch.float64)
    assert abs(float(utils.der_num(fn, t, 1, 1e-5)) - 12.0) < 1e-6
    assert np.allclose(utils.grad_num(fn, t).real, [3.0, 12.0], atol=1e-6)
    J = utils.jac_num(lambda a: torch.stack([a.sum(), (a ** 2).sum()]), t, step_size=1e-7)
    assert J.shape == (2, 2) and np.allclose(J, [[1.0, 2.0], [1.0, 4.0]], atol=1e-5)         # (n_in, n_out)
    assert utils.imarr(torch.arange(6.0).reshape(2, 3, 1)).shape == (3, 2)
    freqs, power = utils.plot_spectral_power(np.sin(0.2 * np.arange(64)), 1e-15, show=False)
    assert freqs.shape == power.shape == (32, ) or power.shape == (32, 1)
