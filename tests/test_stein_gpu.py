"""GPU tests of the bisection + inverse-iteration path for k << N (eig_stein.cu) behind
arcb200_pair_sweep: against numpy.linalg.eigh on the same matrices, including exactly degenerate
and tightly clustered spectra (cluster re-orthogonalisation), both tridiagonalisation paths, and
against the divide & conquer path (ARCB200_STEIN=0)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sweep(h0, M, R, k, sigma, opi):
    from arc_b200 import sweep
    N = len(h0)
    rr, cc = np.nonzero(M)
    ops = sweep.PairOperators(h0, [(rr, cc, M[rr, cc])], N)
    return sweep.pair_sweep(ops, R, k, sigma, opi)


def _ref(h0, M, rv, k, sigma):
    w, v = np.linalg.eigh(np.diag(h0) + M / (rv * 1e-6) ** 3)
    sel = np.sort(np.argsort(np.abs(w - sigma), kind="stable")[:k])
    return w, v, sel


@pytest.mark.parametrize("N,k", [(600, 50), (1100, 40), (1536, 300)])
def test_stein_random_matrices(N, k):
    rng = np.random.default_rng(N)
    M = rng.standard_normal((N, N))
    M = (M + M.T) * 1e-18
    h0 = rng.standard_normal(N) * 3
    R = np.array([1.4, 1.0, 0.8])
    res = _sweep(h0, M, R, k, 0.3, 11)
    for t, rv in enumerate(R):
        w, v, sel = _ref(h0, M, rv, k, 0.3)
        scale = np.abs(w).max()
        assert np.abs(res.y[t] - w[sel]).max() <= 1e-9 * scale
        assert np.abs(res.hl[t] - v[11, sel] ** 2).max() <= 1e-6


def test_stein_degenerate_and_clustered_spectrum():
    """Three identical diagonal blocks (every eigenvalue exactly three-fold) plus a weak coupling
    that splits them by ~1e-10 of the scale: highlight summed over clusters must match LAPACK."""
    rng = np.random.default_rng(3)
    nb = 220
    B = rng.standard_normal((nb, nb))
    B = B + B.T
    N = 3 * nb
    H = np.zeros((N, N))
    for b in range(3):
        H[b * nb:(b + 1) * nb, b * nb:(b + 1) * nb] = B
    for eps in (0.0, 1e-9):
        C = rng.standard_normal((N, N))
        Hc = H + eps * (C + C.T)
        h0 = np.diag(Hc).copy()
        M = (Hc - np.diag(h0)) * 1e-18
        k = 60
        res = _sweep(h0, M, np.array([1.0]), k, 0.0, 5)
        w, v, sel = _ref(h0, M, 1.0, k, 0.0)
        scale = np.abs(w).max()
        # the window edge may cut a 3-fold cluster: compare interior clusters
        assert np.abs(np.sort(res.y[0])[3:-3] - np.sort(w[sel])[3:-3]).max() <= 1e-9 * scale
        cid = np.concatenate([[0], np.cumsum(np.diff(w[sel]) > 1e-6 * scale)])
        a = np.bincount(cid, weights=res.hl[0])
        b_ = np.bincount(cid, weights=v[5, sel] ** 2)
        assert np.abs(a[1:-1] - b_[1:-1]).max() <= 1e-6


def test_stein_matches_divide_and_conquer(monkeypatch):
    rng = np.random.default_rng(8)
    N, k = 1200, 64
    M = rng.standard_normal((N, N))
    M = (M + M.T) * 1e-18
    h0 = rng.standard_normal(N)
    R = np.array([1.1, 0.9])
    a = _sweep(h0, M, R, k, -0.2, 3)
    monkeypatch.setenv("ARCB200_STEIN", "0")
    b = _sweep(h0, M, R, k, -0.2, 3)
    scale = np.abs(a.y).max()
    assert np.abs(a.y - b.y).max() <= 1e-11 * scale
    assert np.abs(a.hl - b.hl).max() <= 1e-8
    assert np.array_equal(a.comp_i[:, :, 0], b.comp_i[:, :, 0])
