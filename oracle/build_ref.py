#!/usr/bin/env python
"""Build the UNMODIFIED reference (minatosato/cymf) into oracle/_ref/ -- test infrastructure only.

The reference's hot path is 9 Cython modules (cymf/{math,linalg,metrics,optimizer,model,bpr,wmf,
glove,evaluator}.pyx).  They are cythonized from where they lie under /root/reference (through a
scratch package of symlinks, so no reference source enters this repository) and compiled with the
system g++ + OpenMP.  Only build products land in oracle/_ref/ (git-ignored, NOT gpurun-ignored, so
the compiled modules travel to the GPU box, which has the same image and therefore the same
Python / NumPy / SciPy ABI).

Accommodations (none touches hot-path arithmetic):
  (a) cymf/evaluator.pyx:89,137 contain the f-string typo `{metric)}` that Cython 3 rejects; a
      sed-patched scratch copy of that one file is cythonized instead of the symlink.
  (b) oracle/shim/cblas.h stands in for the absent system CBLAS header (dead code on our paths).
  (c) compiler directive legacy_implicit_noexcept=True restores the Cython 0.29 semantics the
      reference was written for (without it every cdef call inside prange re-acquires the GIL).
  (d) the package __init__ written into oracle/_ref/cymf/ imports only the compiled modules; the
      reference's own __init__ also pulls in cymf.dataset -> wget (network loaders, out of scope).

Usage:  python oracle/build_ref.py [--force]
It is a no-op (exit 0) when /root/reference is absent (the GPU box): the prebuilt files are used.
"""
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CYMF_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PKG = os.path.join(OUT, "cymf")
MODULES = ["math", "linalg", "metrics", "optimizer", "model", "bpr", "wmf", "glove", "evaluator", "relmf"]
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

INIT_PY = '''"""Import shim for the compiled reference modules (written by oracle/build_ref.py)."""
from .bpr import BPR
from .wmf import WMF
from .glove import GloVe
from .relmf import RelMF
from . import evaluator
from .evaluator import Evaluator, AverageOverAllEvaluator, AoaEvaluator, UnbiasedEvaluator
'''


def built() -> bool:
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(PKG, m + suffix)) for m in MODULES)


def main(force: bool = False) -> int:
    if not os.path.isdir(os.path.join(REF, "cymf")):
        print(f"[build_ref] {REF} not present; using prebuilt oracle/_ref ({'found' if built() else 'MISSING'})")
        return 0
    if built() and not force:
        print("[build_ref] oracle/_ref already built")
        return 0
    import numpy as np

    os.makedirs(PKG, exist_ok=True)
    gen = os.path.join(OUT, "build")
    os.makedirs(gen, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    pyinc = sysconfig.get_paths()["include"]
    with tempfile.TemporaryDirectory(prefix="cymf_ref_") as scratch:
        spkg = os.path.join(scratch, "cymf")
        os.makedirs(spkg)
        open(os.path.join(spkg, "__init__.py"), "w").close()
        for name in os.listdir(os.path.join(REF, "cymf")):
            if name.endswith((".pyx", ".pxd", ".h")):
                os.symlink(os.path.join(REF, "cymf", name), os.path.join(spkg, name))
        # accommodation (a): patched scratch copy of evaluator.pyx only
        ev = os.path.join(spkg, "evaluator.pyx")
        text = open(ev).read()
        os.unlink(ev)
        with open(ev, "w") as f:
            f.write(text.replace("{metric)}", "{metric}"))
        for mod in MODULES:
            cpp = os.path.join(gen, mod + ".cpp")
            subprocess.check_call(
                [sys.executable, "-m", "cython", "--cplus", "-3", "-X", "legacy_implicit_noexcept=True",
                 "-I", scratch, os.path.join(spkg, mod + ".pyx"), "-o", cpp],
                cwd=scratch)
            so = os.path.join(PKG, mod + suffix)
            subprocess.check_call(
                [CXX, "-O3", "-fopenmp", "-std=c++11", "-w", "-shared", "-fPIC",
                 "-I", pyinc, "-I", np.get_include(), "-I", os.path.join(HERE, "shim"), "-I", spkg,
                 cpp, "-o", so, "-fopenmp"])
            print("[build_ref] built", os.path.relpath(so, HERE))
    with open(os.path.join(PKG, "__init__.py"), "w") as f:
        f.write(INIT_PY)
    shutil.rmtree(gen, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main(force="--force" in sys.argv))
