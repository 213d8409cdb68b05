"""Bring up an importable copy of the reference (ARC 3.10.2) in THIS container.

Test/fixture infrastructure only -- nothing in the product path imports this.
It exists so that `tools/gen_golden.py` and `tools/export_atom_data.py` can run
the reference's own CPU path and freeze its outputs as
fixtures under tests/golden/.  `/root/reference` is read-only and does not exist
on the GPU box, so everything here works on a scratch copy under /tmp.

What it does (none of it touches hot-path arithmetic):
  1. copies /root/reference/arc to <scratch>/arc and compiles the reference's own
     arc_c_extensions.c there with gcc -O3 (the reference's setup.py does the same);
  2. stubs matplotlib / h5py / lmfit / asteval (imported at module top by the
     reference, unused on the hot path) with MagicMock modules;
  3. regenerates the LFS blob arc/data/precalculated3j.npy (absent from the
     checkout) from the reference's own Racah-formula fallback (wigner.py:159-196),
     index map wigner.py:77-83, for integer AND half-integer j2 (the survey's
     first regeneration missed half-integer j2, which silently zeroed every
     spin Clebsch-Gordan that is not memoised in the shipped sqlite seeds);
  4. points HOME at a private directory so the reference's sqlite memo tables
     start EMPTY (every matrix element is computed live by the C Numerov path).

Usage:   import tools.refenv as refenv; arc = refenv.load()   # returns module `arc`
"""
import glob
import os
import shutil
import subprocess
import sys
import sysconfig
from unittest.mock import MagicMock

REFERENCE = os.environ.get("ARC_REFERENCE", "/root/reference")
SCRATCH = os.environ.get("ARC_REF_SCRATCH", "/tmp/arcb200_refenv")


def _stub_modules():
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.colors",
                 "matplotlib.ticker", "matplotlib.collections",
                 "matplotlib.colorbar", "h5py", "lmfit", "asteval"]:
        if name in sys.modules and not isinstance(sys.modules[name], MagicMock):
            continue
        m = MagicMock(name=name)
        m.__path__ = []
        m.__name__ = name
        sys.modules[name] = m
    sys.modules["matplotlib"].rcParams = {}


def _build_cext(arcdir):
    import numpy
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(arcdir, "arc_c_extensions" + suffix)
    if os.path.exists(out):
        return
    subprocess.check_call([
        "gcc", "-O3", "-shared", "-fPIC",
        "-I" + numpy.get_include(), "-I" + sysconfig.get_paths()["include"],
        os.path.join(arcdir, "arc_c_extensions.c"), "-o", out, "-lm"])


def _gen_3j(arcdir):
    """tab[2*j1, 2*(23+m1), 2*j2, m2+j2, 2-j3+j1] = Wigner3j(j1,j2,j3,m1,m2,-m1-m2)."""
    import importlib.util
    import numpy as np
    target = os.path.join(arcdir, "data", "precalculated3j.npy")
    if os.path.exists(target):
        return
    orig = np.load

    def fake(path, *a, **k):
        if str(path).endswith("precalculated3j.npy"):
            return np.zeros((1,))
        return orig(path, *a, **k)
    np.load = fake
    try:
        spec = importlib.util.spec_from_file_location(
            "_ref_wigner_boot", os.path.join(arcdir, "wigner.py"))
        w = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(w)
    finally:
        np.load = orig
    w.wignerPrecal = False
    J = w.wignerPrecalJmax
    tab = np.zeros((2 * J, 4 * J + 1, 5, 5, 5))
    rp2 = w.roundPy2
    for j1x2 in range(0, 2 * J):
        j1 = j1x2 / 2
        for j2x2 in range(0, 5):
            j2 = j2x2 / 2
            for m2i in range(0, j2x2 + 1):
                m2 = m2i - j2
                j3 = abs(j1 - j2)
                while j3 < j1 + j2 + 0.1:
                    # same index expression as the lookup at wigner.py:77-83 (integer AND
                    # half-integer j2: round(2 - j3 + j1 + 1e-15) maps x.5 upward)
                    k = round(rp2(2 - j3 + j1))
                    for m1x2 in range(-j1x2, j1x2 + 1, 2):
                        m1 = m1x2 / 2
                        m3 = -m1 - m2
                        if abs(m3) > j3 + 0.1:
                            continue
                        v = w.Wigner3j(j1, j2, j3, m1, m2, m3)
                        tab[round(rp2(2 * j1)), round(rp2(2 * (J + m1))),
                            round(rp2(2.0 * j2)), round(rp2(m2 + j2)), k] = v
                    j3 += 1.0
    np.save(target, tab)


def load(fresh_home=True, home_name="home"):
    """Return the reference `arc` module, imported from the scratch copy."""
    arcdir = os.path.join(SCRATCH, "arc")
    if not os.path.isdir(arcdir):
        os.makedirs(SCRATCH, exist_ok=True)
        shutil.copytree(os.path.join(REFERENCE, "arc"), arcdir)
        subprocess.check_call(["chmod", "-R", "u+w", SCRATCH])
    _build_cext(arcdir)
    _stub_modules()
    _gen_3j(arcdir)
    home = os.path.join(SCRATCH, home_name)
    os.makedirs(home, exist_ok=True)
    os.environ["HOME"] = home
    if SCRATCH not in sys.path:
        sys.path.insert(0, SCRATCH)
    import arc  # noqa: E402  (reference package from scratch copy)
    if fresh_home:
        purge_memo()
    return arc


def purge_memo():
    """Empty the reference's matrix-element memo (SURVEY.md 8c rule ii)."""
    d = os.path.expanduser("~/.arc-data")
    for f in (glob.glob(os.path.join(d, "*_matrix_elements.npy"))
              + glob.glob(os.path.join(d, "*_precalculated.db"))):
        os.remove(f)


if __name__ == "__main__":
    a = load()
    print("reference arc imported from", a.__file__)
