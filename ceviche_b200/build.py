"""Build the C-ABI shared library in-tree with nvcc for sm_100a (no JIT cache: the .so
travels with the repo snapshot to the GPU box)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("CEV_LIB_PATH") or os.path.join(HERE, "libceviche_b200.so")
HEADER = os.path.join(ROOT, "include", "ceviche_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--threads", "0",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    deps = [HEADER] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return [d for d in deps if os.path.isfile(d)]


def is_stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    return None


def build_library(force=False, verbose=False, extra_flags=()):
    """Compile csrc/*.cu -> libceviche_b200.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build %s" % LIB_PATH)
    tmp = LIB_PATH + ".tmp%d" % os.getpid()
    extra_flags = list(extra_flags) + os.environ.get("CEV_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-I", os.path.join(ROOT, "include"), "-o", tmp] + sources()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n%s\n%s" % (res.stdout, res.stderr))
    if verbose:
        print(res.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force=True, verbose="-v" in sys.argv))
