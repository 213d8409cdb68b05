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


def test_bad_arguments_return_an_error_before_any_cuda_call():
    """The sweep entry points validate sizes and pointers on the host: a bad call comes back as -1
    with a message (no hang on batch = 0, no device fault on a null pointer).  No GPU is needed
    because the checks run before the first CUDA call."""
    from arc_b200 import _lib
    lib = _lib.load()
    buf = np.zeros(64)
    p = buf.ctypes.data
    nl = _lib.LaunchCount()

    def stark(N=4, nF=2, idx0=0, topk=0, batch=1, h0=p, work=p):
        return lib.arcb200_stark_sweep(0, None, N, h0, p, nF, p, idx0, None, topk, 0.95, p, p, None, None,
                                       work, 1 << 20, batch, nl.addr)
    assert stark(nF=0) == 0                                   # empty sweep: nothing to do
    for kw, msg in ((dict(batch=0), b"batch"), (dict(batch=70000), b"batch"), (dict(N=0), b"N = 0"),
                    (dict(idx0=4), b"index"), (dict(idx0=-1), b"index"), (dict(h0=None), b"null"),
                    (dict(work=None), b"null"), (dict(topk=2), b"composition"), (dict(topk=-1), b"topk")):
        assert stark(**kw) == -1, kw
        assert msg in lib.arcb200_last_error(), (kw, lib.arcb200_last_error())

    def pair(N=4, nR=2, k=2, opi=0, batch=1, norders=1, M=p, topk=0):
        return lib.arcb200_pair_sweep(0, None, N, p, norders, M, nR, p, k, 0.0, opi, None, topk, 0.95, p, p,
                                      None, None, None, p, 1 << 20, batch, nl.addr)
    assert pair(nR=0) == 0
    for kw, msg in ((dict(batch=0), b"batch"), (dict(k=0), b"k must"), (dict(k=5), b"k must"),
                    (dict(opi=9), b"index"), (dict(M=None), b"operator"), (dict(norders=-1), b"norders"),
                    (dict(topk=3), b"composition"), (dict(N=0), b"N = 0")):
        assert pair(**kw) == -1, kw
        assert msg in lib.arcb200_last_error(), (kw, lib.arcb200_last_error())
    assert lib.arcb200_syevd_batched(0, None, 4, 1, None, p, p, p, 1 << 20, nl.addr) == -1
    assert lib.arcb200_syevd_batched(0, None, 4, 0, None, p, p, p, 1 << 20, nl.addr) == 0
    assert lib.arcb200_sytrd_batched(0, None, 0, 1, p, p, p, p, p, 1 << 20, 0, nl.addr) == -1
    assert lib.arcb200_stedc_batched(0, None, 4, 1, p, None, p, p, p, 1 << 20, nl.addr) == -1
    assert lib.arcb200_sy2sb_batched(0, None, 256, 70000, p, p, p, 1 << 20, nl.addr) == -1
    assert lib.arcb200_radial_integrals(0, None, 3, None, p, p, p, p, p) == -1
    assert lib.arcb200_numerov_batch(0, None, 2, None, p, p, p, p, p, p, 1, p) == -1
    assert lib.arcb200_semiclassical_table(0, None, 2, p, 3, p, 256, p) == -1
    assert lib.arcb200_semiclassical_table(0, None, 2, None, 1, p, 256, p) == -1
    assert lib.arcb200_scatter_coo(0, None, 4, 2, None, p, p, p) == -1
    assert nl.value == 0


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
