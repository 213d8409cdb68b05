"""CPU tests of the C-ABI boundary: the library loads, exports every symbol that
include/arcb200.h declares, and the product path has no CPU fallback and never touches
oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "arc-alkali-rydberg-calculator_b200")


def _declared():
    src = open(os.path.join(ROOT, "include", "arcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(arcb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from arc_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), "libarcb200.so does not export %s" % n
    # the ctypes table binds exactly the declared entry points
    assert sorted(_lib.exported_symbols()) == names
    assert _lib.load().arcb200_abi_version() == 2


def test_state_struct_layout_and_lengths():
    from arc_b200 import _lib, radial
    from oracle import oracle
    assert _lib.STATE_DTYPE.itemsize == 144
    lib = _lib.load()
    recs = np.array([radial.make_state(2.0858, 2.0 * n * (n + 15.0), 0.001, 0.01, 0.01, 1, .5, 1.5,
                                       -1e-3, 9.076, 0.0073, 37, 1., 1., 1., 1., 1., 0.99999)
                     for n in (20, 61, 120)])
    offs = np.zeros(4, dtype=np.int64)
    total = lib.arcb200_numerov_lengths(3, recs.ctypes.data, offs.ctypes.data)
    assert total > 0 and offs[0] == 0 and np.all(np.diff(offs) > 0) and np.all(offs % 32 == 0)
    for r in recs:
        L = oracle._lib().arc_oracle_numerov_length(float(r["innerLimit"]), float(r["outerLimit"]), 0.001)
        assert int(r["length"]) == L            # (int)((sqrt(outer)-sqrt(inner))/step), cext:131
    bad = recs.copy()
    bad["step"] = 0.0
    assert lib.arcb200_numerov_lengths(3, bad.ctypes.data, offs.ctypes.data) < 0
    assert b"invalid" in lib.arcb200_last_error()


def test_workspace_size_is_monotone():
    from arc_b200 import _lib
    lib = _lib.load()
    a = lib.arcb200_sweep_workspace_bytes(410, 1, 410)
    b = lib.arcb200_sweep_workspace_bytes(410, 8, 410)
    c = lib.arcb200_sweep_workspace_bytes(2363, 1, 250)
    assert 0 < a < b and c > a
    assert lib.arcb200_sweep_workspace_bytes(0, 1, 1) < 0


def test_product_never_imports_the_oracle():
    for dirpath, _d, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle\.|liboracle|oracle/", txt, re.M), \
                    "%s uses the oracle" % f


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import arc_b200
    from arc_b200._lib import ArcB200Error
    with pytest.raises(ArcB200Error):
        arc_b200.Rubidium().getRadialMatrixElement(30, 0, .5, 30, 1, 1.5)
    with pytest.raises(ArcB200Error):
        arc_b200.sweep.syevd_batched(np.eye(4))
