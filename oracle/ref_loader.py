"""Load the reference's own FDTD files, unmodified, from /root/reference (TEST INFRASTRUCTURE).

HIPS autograd and matplotlib are not installed in the build container, so
``import ceviche`` fails as shipped.  The FDTD forward path only uses
``autograd.numpy`` as a numpy alias, so we alias it, stub the two decorators
``ceviche/utils.py`` evaluates at import time, and load constants / utils /
derivatives / fdtd by file path under a synthetic ``ceviche`` package
(``ceviche/__init__.py`` is NOT executed: it would pull in the FDFD modules,
which need the real autograd).

Only usable where /root/reference exists (the build container).  It never
travels to the GPU box; tests that need it skip there and rely on
``tests/golden/`` instead.
"""
import importlib.util
import os
import sys
import types
import warnings

REFERENCE_DIR = os.environ.get("CEVICHE_REFERENCE_DIR", "/root/reference/ceviche")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "fdtd.py"))


def load():
    """Return the reference ``ceviche.fdtd`` module (class ``fdtd`` inside)."""
    if "ceviche.fdtd" in sys.modules and getattr(sys.modules["ceviche"], "_oracle_shim", False):
        return sys.modules["ceviche.fdtd"]
    if not available():
        raise RuntimeError("reference sources not found under %s" % REFERENCE_DIR)
    import numpy

    if "autograd" not in sys.modules:
        ag = types.ModuleType("autograd")
        ag.numpy = numpy
        ext = types.ModuleType("autograd.extend")
        ext.primitive = lambda f: f
        ext.defvjp = ext.defjvp = lambda *a, **k: None
        ext.vspace = None
        ag.extend = ext
        sys.modules["autograd"] = ag
        sys.modules["autograd.numpy"] = numpy
        sys.modules["autograd.extend"] = ext
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        pylab = types.ModuleType("matplotlib.pylab")
        mpl.pylab = pylab
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pylab"] = pylab

    pkg = types.ModuleType("ceviche")
    pkg.__path__ = [REFERENCE_DIR]
    pkg._oracle_shim = True
    sys.modules["ceviche"] = pkg
    mods = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        for name in ("constants", "utils", "derivatives", "fdtd"):
            spec = importlib.util.spec_from_file_location(
                "ceviche." + name, os.path.join(REFERENCE_DIR, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules["ceviche." + name] = mod
            spec.loader.exec_module(mod)
            setattr(pkg, name, mod)
            mods[name] = mod
    return mods["fdtd"]


def load_modes():
    """The reference's ``ceviche.modes`` (get_modes / insert_mode), unmodified.  It imports
    ``compute_derivative_matrices`` from ``ceviche.fdfd``, which cannot be imported without the real autograd; the
    function itself lives in ``ceviche/derivatives.py`` (already loaded), so a stand-in ``ceviche.fdfd`` module
    re-exports it."""
    load()
    if "ceviche.modes" in sys.modules:
        return sys.modules["ceviche.modes"]
    fdfd = types.ModuleType("ceviche.fdfd")
    fdfd.compute_derivative_matrices = sys.modules["ceviche.derivatives"].compute_derivative_matrices
    sys.modules["ceviche.fdfd"] = fdfd
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        spec = importlib.util.spec_from_file_location("ceviche.modes", os.path.join(REFERENCE_DIR, "modes.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ceviche.modes"] = mod
        spec.loader.exec_module(mod)
    return mod
