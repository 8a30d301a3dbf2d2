"""Load the reference's own FDTD files, unmodified, from /root/reference (TEST INFRASTRUCTURE).

HIPS autograd and matplotlib are not installed in the build container, so
``import ceviche`` fails as shipped.  The FDTD forward path only uses
``autograd.numpy`` as a numpy alias, so we alias it, stub the two decorators
``ceviche/utils.py`` evaluates at import time, and load constants / utils /
derivatives / fdtd by file path under a synthetic ``ceviche`` package
(``ceviche/__init__.py`` is NOT executed: it would pull in the FDFD modules,
which need the real autograd).

Where it loads from: /root/reference/ceviche in the build container; else ``oracle/_ref/ceviche`` -- UNMODIFIED
copies of the four files (+ LICENSE) that ``__graft_entry__.build()`` makes where /root/reference exists.  That
directory is git-ignored (never part of the repository's history) but travels with the gpurun snapshot, so that
``bench.py --impl reference`` can time the reference's own code on the GPU box's host cores.  Tests that need the
reference skip where neither exists and rely on ``tests/golden/`` instead.
"""
import importlib.util
import os
import sys
import types
import warnings

VENDORED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ceviche")
FILES = ("constants.py", "utils.py", "derivatives.py", "fdtd.py")


def _find():
    for d in (os.environ.get("CEVICHE_REFERENCE_DIR"), "/root/reference/ceviche", VENDORED_DIR):
        if d and all(os.path.isfile(os.path.join(d, f)) for f in FILES):
            return d
    return os.environ.get("CEVICHE_REFERENCE_DIR") or "/root/reference/ceviche"


REFERENCE_DIR = _find()


def available():
    return all(os.path.isfile(os.path.join(REFERENCE_DIR, f)) for f in FILES)


def vendor(src="/root/reference"):
    """Copy the reference's four FDTD files and its licence, unmodified, into git-ignored oracle/_ref/ (called by
    __graft_entry__.build() in the build container).  Returns the directory, or None when there is no reference."""
    import shutil
    pkg = os.path.join(src, "ceviche")
    if not all(os.path.isfile(os.path.join(pkg, f)) for f in FILES):
        return None
    os.makedirs(VENDORED_DIR, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(pkg, f), os.path.join(VENDORED_DIR, f))
    for lic in ("LICENSE", "LICENSE.md", "LICENSE.txt"):
        if os.path.isfile(os.path.join(src, lic)):
            shutil.copyfile(os.path.join(src, lic), os.path.join(os.path.dirname(VENDORED_DIR), lic))
    return VENDORED_DIR


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
