"""CPU-only checks of the host layer and of the C-ABI library's export table (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import fdtd_numpy as onp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constants_match_reference_digits():
    from ceviche_b200 import constants as c
    assert c.EPSILON_0 == 8.85418782e-12 and c.MU_0 == 1.25663706e-6
    assert c.C_0 == onp.C_0 == 299792458.13099605


@pytest.mark.parametrize("shape,npml", [((12, 7, 9), (3, 2, 0)), ((200, 200, 1), (20, 20, 0)), ((5, 6, 40), (0, 1, 9))])
def test_host_sigma_profiles_equal_oracle(shape, npml):
    from ceviche_b200.fdtd import sigma_profiles
    dt = onp.time_step(5e-8)
    sH, sD = sigma_profiles(shape, npml, dt)
    oH, oD = onp.sigma_profiles(shape, npml, dt)
    for a in range(3):
        assert np.array_equal(sH[a], oH[a]) and np.array_equal(sD[a], oD[a])


def test_reshape_to_nd_raises_beyond_3d():
    from ceviche_b200.fdtd import reshape_to_ND
    assert reshape_to_ND(np.ones((4, 5)), 3).shape == (4, 5, 1)
    with pytest.raises(ValueError):
        reshape_to_ND(np.ones((2, 2, 2, 2)), 3)


def test_library_exports_every_declared_symbol():
    """Every function declared in include/ceviche_b200.h is exported by the built .so."""
    from ceviche_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "ceviche_b200.h")).read()
    declared = set(re.findall(r"\b(cev_[a-z_A-Z0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(_lib.EXPORTS) == declared
    assert lib.cev_abi_version() == 3


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import ceviche_b200
    from ceviche_b200._lib import CevicheB200Error
    with pytest.raises(CevicheB200Error):
        ceviche_b200.fdtd(np.ones((4, 4)), 5e-8, [1, 1, 0])


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under ceviche_b200/ may reference it."""
    pkg = os.path.join(ROOT, "ceviche_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f


def test_spectrum_helpers_match_numpy_formulas():
    """ceviche/utils.py:373-400 restated over torch.fft (CPU tensors here)."""
    import torch
    from ceviche_b200 import utils
    rng = np.random.default_rng(0)
    series = rng.standard_normal((64, 2))
    dt = 1e-16
    freqs, sig = utils.get_spectrum(series, dt)
    ref = np.fft.fft(np.hamming(64).reshape(64, 1) * series, axis=0)
    # the reference calls np.fft.fft on a (steps, n) array, i.e. along the LAST axis; for the (steps, 1)
    # series its callers use that is a no-op per row -- we transform along time, which is what is meant
    np.testing.assert_allclose(sig.numpy(), ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(freqs.numpy(), np.fft.fftfreq(64, d=dt))
    f2, pw = utils.get_spectral_power(series[:, 0], dt)
    np.testing.assert_allclose(pw.numpy()[:, 0], np.abs(ref[:, 0]) ** 2, rtol=1e-12)


def test_coupled_components_closure():
    """Which field components a source can reach (curl couplings of ceviche/derivatives.py:16-30 on
    grids with singleton axes), checked against what the numpy oracle actually produces."""
    from ceviche_b200.fdtd import coupled_components
    assert coupled_components((200, 200, 1), 0b100) == 0b011100      # TM: Ez, Hx, Hy
    assert coupled_components((200, 200, 1), 0b001) == 0b100011      # TE: Ex, Ey, Hz
    assert coupled_components((64, 1, 1), 0b100) == 0b010100
    assert coupled_components((64, 1, 1), 0b001) == 0b000001
    assert coupled_components((8, 8, 8), 0b010) == 0b111111
    assert coupled_components((8, 8, 8), 0) == 0
    for shape in ((7, 6, 1), (9, 1, 1), (1, 6, 5), (1, 1, 8), (5, 1, 4), (4, 3, 5)):
        for c, comp in enumerate("xyz"):
            prof = np.zeros(shape)
            prof[tuple(n // 2 for n in shape)] = 1.0
            O = onp.OracleFDTD(1 + np.random.default_rng(1).random(shape), 5e-8, [0, 0, 0])
            O.run(12, [(comp, prof, np.ones(12))], [])
            f = O.fields()
            seen = sum(1 << q for q, n in enumerate("xyz") if np.any(f["D" + n])) | \
                sum(8 << q for q, n in enumerate("xyz") if np.any(f["H" + n]))
            assert seen == coupled_components(shape, 1 << c), (shape, comp)


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: the header must compile as C99 (and as C++), with no torch / CUDA types, and a C
    translation unit that calls through it must link against the built library."""
    import shutil
    import subprocess
    import tempfile
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    header = os.path.join(ROOT, "include", "ceviche_b200.h")
    subprocess.run([gcc, "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", header], check=True)
    subprocess.run([shutil.which("g++") or gcc, "-x", "c++", "-std=c++17", "-fsyntax-only", header], check=True)
    assert "#include <torch" not in open(header).read() and "#include <cuda" not in open(header).read()
    from ceviche_b200 import build
    lib = build.LIB_PATH
    if not os.path.isfile(lib):
        pytest.skip("library not built")
    src = r"""
#include "ceviche_b200.h"
#include <stdio.h>
int main(void) {
    cev_fdtd* plan = 0;
    /* no GPU needed: an invalid dtype is rejected before any CUDA call */
    int rc = cev_fdtd_create(&plan, 0, 7, 1, 4, 4, 4, 5e-8, 1e-16, 0, 0);
    printf("%d %d %s\n", cev_abi_version(), rc, cev_last_error());
    return (rc != 0 && plan == 0) ? 0 : 1;
}
"""
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), c, lib, "-Wl,-rpath," + os.path.dirname(lib),
                        "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert out.stdout.split()[0] == "3" and "dtype" in out.stdout


def test_hips_binding_is_a_soft_dependency():
    """Without HIPS autograd the module imports and exposes no primitive; with an `extend` module the registration uses
    exactly the reference's protocol (one maker for eps_r, None for the other arguments)."""
    import types
    from ceviche_b200 import hips
    import importlib.util
    if importlib.util.find_spec("autograd") is None:
        assert hips.fdtd_series is None
    calls = []
    ext = types.SimpleNamespace(primitive=lambda f: f, defvjp=lambda f, *m: calls.append(("vjp", len(m), m[1:])),
                                defjvp=lambda f, *m: calls.append(("jvp", len(m), m[1:])))
    prim = hips.register(ext)
    assert callable(prim) and [c[0] for c in calls] == ["vjp", "jvp"]
    assert all(c[1] == 5 and all(x is None for x in c[2]) for c in calls)


def test_design_region_boxes():
    """Host logic of the design-box gradients: eps_r[i, j, k] enters the Yee averages of cells (i..i+1, j..j+1, k..k+1)
    (utils.py:167-174), so the 1/eps cotangent box is one cell larger on the high side; on x-slabs it is clipped to the
    slab (possibly empty); a box whose +1 would wrap falls back to the whole grid."""
    import types
    from ceviche_b200 import autodiff
    from ceviche_b200.slab import SlabFDTD, partition
    sim = types.SimpleNamespace(design_region=((2, 5), (1, 4), (0, 3)), grid_shape=(8, 6, 4))
    assert autodiff._grad_box(sim) == [(2, 6), (1, 5), (0, 4)]
    sim.design_region = ((2, 8), (1, 4), (0, 3))           # x1 + 1 would wrap: everywhere
    assert autodiff._grad_box(sim) is None
    sim.design_region = None
    assert autodiff._grad_box(sim) is None
    sim2d = types.SimpleNamespace(design_region=((2, 5), (1, 4), (0, 1)), grid_shape=(8, 6, 1))
    assert autodiff._grad_box(sim2d) == [(2, 6), (1, 5), (0, 1)]          # a full axis needs no extra cell
    sim.design_region = ((2, 9), (1, 4), (0, 3))
    with pytest.raises(ValueError):
        autodiff._grad_box(sim)
    parts = partition(12, 3)
    assert parts == [(0, 4), (4, 8), (8, 12)]
    boxes = []
    for lo, hi in parts:
        slab = types.SimpleNamespace(design_region=((3, 6), (1, 4), (0, 3)), Nx=12, Ny=6, Nz=5, lo=lo, hi=hi)
        boxes.append(SlabFDTD._local_grad_box(slab))
    assert boxes[0] == [3, 4, 1, 5, 0, 4]                   # global planes 3..6 (+1) cut at the slab boundary
    assert boxes[1] == [0, 3, 1, 5, 0, 4]
    assert boxes[2][0] >= boxes[2][1]                       # does not touch the third slab: empty x-range
    slab = types.SimpleNamespace(design_region=None, Nx=12, Ny=6, Nz=5, lo=0, hi=4)
    assert SlabFDTD._local_grad_box(slab) is None


def test_jacobian_forward_mode_batches_directions_on_cpu():
    """Host logic of jacobian(mode='forward') (ceviche/jacobians.py:38-51): a batch of directions through ONE evaluation of
    `fun` (torch.vmap over torch.func.jvp), in chunks; a `fun` torch.vmap cannot trace falls back to one pass per
    direction.  Checked on torch-only functions (the FDTD nodes need a GPU: tests/test_gpu_gradients.py)."""
    import torch
    from ceviche_b200 import jacobian, jacobians
    calls = []

    def fun(x):
        calls.append(1)
        y = torch.zeros(3, dtype=torch.float64)
        y[0:2] = x[0:2] ** 2
        return torch.cat([y, (x[0] * x[2]).reshape(1)])
    x = np.array([1.0, 2.0, 3.0])
    exact = np.array([[2.0, 0, 0], [0, 4.0, 0], [0, 0, 0], [3.0, 0, 1.0]])
    J = jacobian(fun, mode='forward')(x)
    assert jacobians.last_forward_path == "batched" and len(calls) == 1
    assert np.array_equal(J.numpy(), exact)
    assert np.array_equal(jacobian(fun, mode='reverse')(x).numpy(), exact)
    assert np.allclose(jacobian(fun, mode='numerical')(x).numpy(), exact, atol=1e-5)
    del calls[:]
    old = jacobians.forward_chunk
    try:
        jacobians.forward_chunk = 2
        assert np.array_equal(jacobian(fun, mode='forward')(x).numpy(), exact) and len(calls) == 2
    finally:
        jacobians.forward_chunk = old

    def untraceable(x):
        return fun(x) * float(np.asarray(x.detach())[0])         # a numpy round trip: no data pointer under torch.vmap
    J = jacobian(untraceable, mode='forward')(x)
    assert jacobians.last_forward_path.startswith("per-direction")
    assert np.array_equal(J.numpy(), exact)
    assert tuple(jacobian(lambda c: c * 3, mode='forward')(2.0).shape) == (1, 1)     # scalar in, scalar out
    with pytest.raises(ValueError):
        jacobian(fun, mode='sideways')


def test_host_tensors_stay_plain_under_torch_func_transforms():
    """Under torch.vmap(torch.func.jvp(.)) every op returns a wrapped tensor without a data pointer; the object's
    bookkeeping runs inside autodiff.plain() and strips wrappers with autodiff.base(), so the C ABI gets real pointers.
    torch.func.grad (whose wrappers do not say requires_grad) is refused instead of silently treated as a constant."""
    import types
    import torch
    from torch._C._functorch import is_functorch_wrapped_tensor as wrapped
    from ceviche_b200 import autodiff
    x = torch.tensor([1.0, 2.0], dtype=torch.float64)
    seen = {}

    def fun(c):
        m = 1 / (x * c)
        seen["top_level_factory_is_wrapped"] = wrapped(torch.zeros(2))
        with autodiff.plain():
            z = torch.zeros(2, dtype=torch.float64)
            seen["plain_factory_is_wrapped"] = wrapped(z)
            seen["ptr"] = z.data_ptr() != 0 and (z.clone() * 2).data_ptr() != 0
            b = autodiff.base(m).to(torch.float64).contiguous()
            seen["base"] = (not wrapped(b)) and b.data_ptr() != 0 and torch.equal(b, 1 / (x * x))
        sim = types.SimpleNamespace(_mE64=[m, m, m], _H=[z] * 3, _D=[z] * 3, _pml=None)
        seen["needs_grad"] = autodiff.needs_grad(sim, [None, None, None])
        return (m + z).sum()
    t = torch.vmap(lambda d: torch.func.jvp(fun, (x,), (d,))[1])(torch.eye(2, dtype=torch.float64))
    assert torch.allclose(t, -1 / x ** 3)        # evaluated at c = x: m = 1 / x^2
    assert seen == {"top_level_factory_is_wrapped": True, "plain_factory_is_wrapped": False, "ptr": True, "base": True,
                    "needs_grad": True}
    with pytest.raises(NotImplementedError):                 # a batch of permittivities has no single plain value
        torch.vmap(lambda e: autodiff.base(e).sum())(torch.ones(2, 3))

    def through_func_grad(c):
        m = 1 / (x * c)
        z = torch.zeros(2, dtype=torch.float64)
        autodiff.needs_grad(types.SimpleNamespace(_mE64=[m, m, m], _H=[z] * 3, _D=[z] * 3, _pml=None), [None, None, None])
        return m.sum()
    with pytest.raises(NotImplementedError):
        torch.func.grad(through_func_grad)(x)
    with autodiff.plain():                                   # no transform active: a no-op context
        assert not wrapped(torch.zeros(1))
