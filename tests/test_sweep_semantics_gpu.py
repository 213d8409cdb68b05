"""GPU parity tests of the sweep epilogue semantics against the oracle on the same matrices:
composition (_stateComposition2) for EVERY eigenvector, driving-weighted highlight, k-clamp,
empty / single-point sweeps, exact degeneracies (divalent triplet manifold)."""
import os

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


def _rand_stark(N, seed, coupling=1.0):
    rng = np.random.default_rng(seed)
    h0 = np.sort(rng.standard_normal(N)) * 50.0
    V = rng.standard_normal((N, N)) * coupling
    V = (V + V.T) / 2
    np.fill_diagonal(V, 0.0)
    return h0, V


def test_composition_all_vectors_match_oracle():
    from arc_b200 import sweep
    h0, V = _rand_stark(97, 3)
    F = np.array([0.0, 0.7, 3.1])
    res = sweep.stark_sweep(h0, V, F, idx0=5, topk=3, contrib_max=0.95)
    y_o, hl_o, comp_o = oracle.stark_diagonalise(np.diag(h0), V, F, 5, upTo=4, totalContributionMax=0.95)
    assert np.abs(res.y - y_o).max() <= 1e-9 * np.abs(y_o).max()
    assert np.abs(res.hl - hl_o).max() <= 1e-6
    for f in range(len(F)):
        for i in range(97):
            ref = comp_o[f][i]
            got = [(res.comp_c[f, i, t], res.comp_i[f, i, t]) for t in range(3) if res.comp_i[f, i, t] >= 0]
            assert len(got) == len(ref), (f, i, got, ref)          # same stopping rule
            for (gc, gi), (rc, ri) in zip(got, ref):
                assert gi == ri
                assert abs(abs(gc) - abs(rc)) <= 1e-6


def test_driving_weighted_highlight_matches_oracle():
    from arc_b200 import sweep
    h0, V = _rand_stark(64, 4)
    rng = np.random.default_rng(9)
    w = np.abs(rng.standard_normal(64))
    w[::3] = 0.0
    w = w / w.max()
    F = np.linspace(0, 2, 4)
    res = sweep.stark_sweep(h0, V, F, idx0=0, drive_w=w)
    _y, hl_o, _ = oracle.stark_diagonalise(np.diag(h0), V, F, 0, drive_w=w, want_composition=False)
    assert np.abs(res.hl - hl_o).max() <= 1e-6


def test_empty_and_single_point_sweeps():
    from arc_b200 import sweep
    h0, V = _rand_stark(40, 5)
    r0 = sweep.stark_sweep(h0, V, np.zeros(0), idx0=0)
    assert r0.y.shape == (0, 40) and r0.kernel_launches == 0
    r1 = sweep.stark_sweep(h0, V, np.array([0.0]), idx0=7)
    assert np.abs(r1.y[0] - h0).max() <= 1e-12 * np.abs(h0).max()       # F = 0: H is diagonal
    assert abs(r1.hl[0].sum() - 1.0) < 1e-12
    # tiny bases
    for N in (1, 2, 3):
        h0, V = _rand_stark(N, 6)
        r = sweep.stark_sweep(h0, V, np.array([0.5, 1.5]), idx0=0, topk=min(3, N))
        y_o, hl_o, _ = oracle.stark_diagonalise(np.diag(h0), V, [0.5, 1.5], 0, want_composition=False)
        assert np.abs(r.y - y_o).max() <= 1e-12 * max(1.0, np.abs(y_o).max())
        assert np.abs(r.hl - hl_o).max() <= 1e-9


def test_exactly_degenerate_diagonal():
    """divalent-like manifold: many exactly equal diagonal entries (SURVEY hard part 4)."""
    from arc_b200 import sweep
    N = 120
    rng = np.random.default_rng(8)
    h0 = np.repeat(np.arange(12.0), 10)               # 12 levels, each 10-fold degenerate
    V = np.zeros((N, N))
    idx = rng.integers(0, N, size=(300, 2))
    for a, b in idx:
        if a != b:
            V[a, b] = V[b, a] = rng.standard_normal()
    F = np.array([0.0, 1e-3, 0.5])
    res = sweep.stark_sweep(h0, V, F, idx0=3)
    y_o, hl_o, _ = oracle.stark_diagonalise(np.diag(h0), V, F, 3, want_composition=False)
    assert np.abs(res.y - y_o).max() <= 1e-9 * np.abs(y_o).max()
    for f in range(3):                                 # highlight summed over degenerate clusters
        cid = np.concatenate([[0], np.cumsum(np.diff(y_o[f]) > 1e-7)])
        assert np.abs(np.bincount(cid, weights=res.hl[f]) - np.bincount(cid, weights=hl_o[f])).max() <= 1e-6


def test_pair_sweep_window_semantics():
    """k nearest sigma, ascending; sigma != 0; k == N (all)."""
    from arc_b200 import sweep
    rng = np.random.default_rng(12)
    N = 150
    h0 = rng.standard_normal(N) * 3
    M = rng.standard_normal((N, N)) * 1e-18
    M = (M + M.T) / 2
    np.fill_diagonal(M, 0.0)
    r, c = np.nonzero(M)
    ops = sweep.PairOperators(h0, [(r, c, M[r, c])], N)
    R = np.array([5.0, 2.0, 1.0])
    for k, sigma in ((20, 0.0), (37, 1.3), (N, -0.4)):
        res = sweep.pair_sweep(ops, R, k, sigma, opi=11)
        for i, rv in enumerate(R):
            H = np.diag(h0) + M / (rv * 1e-6) ** 3
            w, v = np.linalg.eigh(H)
            sel = np.sort(np.argsort(np.abs(w - sigma), kind="stable")[:k])
            assert np.abs(res.y[i] - w[sel]).max() <= 1e-9 * np.abs(w).max()
            assert np.abs(res.hl[i] - v[11, sel] ** 2).max() <= 1e-6
            top = np.argmax(np.abs(v[:, sel]), axis=0)
            assert np.array_equal(res.comp_i[i, :, 0], top)


def test_stark_c4_divalent_triplet_runs_and_matches_numpy():
    """BASELINE configs[3]-type case at reduced n range: Sr88 triplet manifold (s=1)."""
    import arc_b200
    calc = arc_b200.StarkMap(arc_b200.Strontium88())
    calc.defineBasis(56, 0, 1, 0, 54, 58, 14, s=1)
    N = len(calc.basisStates)
    assert N > 150
    F = np.linspace(0, 150, 5)
    calc.diagonalise(F)
    y_o, hl_o, _ = oracle.stark_diagonalise(calc.mat1, calc.mat2, F, calc.indexOfCoupledState, want_composition=False)
    y = np.array(calc.y)
    assert np.abs(y - y_o).max() <= 1e-9 * np.abs(y_o).max()
    k = np.argmax(hl_o, axis=1)
    hl = np.array(calc.highlight)
    for f in range(len(F)):
        cid = np.concatenate([[0], np.cumsum(np.diff(y_o[f]) > 1e-7 * np.abs(y_o).max())])
        assert np.abs(np.bincount(cid, weights=hl[f]) - np.bincount(cid, weights=hl_o[f])).max() <= 1e-6


def test_stark_driving_from_state_matches_reference(gold_dir):
    """drivingFromState highlight (calculations_atom_single.py:887-962, 1003-1021)."""
    import arc_b200
    g = np.load(os.path.join(gold_dir, "stark_driving_golden.npz"))
    a = g["args"]
    calc = arc_b200.StarkMap(arc_b200.Rubidium())
    calc.defineBasis(int(a[0]), int(a[1]), float(a[2]), float(a[3]), int(a[4]), int(a[5]), int(a[6]))
    d = g["drive"]
    calc.diagonalise(g["fields"], drivingFromState=[int(d[0]), int(d[1]), float(d[2]), float(d[3]), int(d[4])])
    assert abs(calc.maxCoupling - float(g["maxCoupling"])) <= 1e-8 * float(g["maxCoupling"])
    y = np.array(calc.y)
    scale = np.abs(g["y"]).max()
    assert np.abs(y - g["y"]).max() <= 1e-9 * scale
    hl = np.array(calc.highlight)
    for f in range(len(g["fields"])):
        cid = np.concatenate([[0], np.cumsum(np.diff(g["y"][f]) > 1e-7 * scale)])
        assert np.abs(np.bincount(cid, weights=hl[f]) - np.bincount(cid, weights=g["highlight"][f])).max() <= 1e-6


def test_pair_driving_from_state_and_k_clamp(gold_dir):
    """drivingFromState for pair states (calculations_atom_pairstate.py:3332-3413) and the
    noOfEigenvectors >= N-1 clamp (:3319-3327)."""
    import arc_b200
    g = np.load(os.path.join(gold_dir, "pair_driving_golden.npz"))
    c, b, d = g["ctor"], g["basis"], g["drive"]
    calc = arc_b200.PairStateInteractions(arc_b200.Rubidium(), int(c[0]), int(c[1]), float(c[2]), int(c[3]),
                                          int(c[4]), float(c[5]), float(c[6]), float(c[7]))
    calc.defineBasis(float(b[0]), float(b[1]), int(b[2]), int(b[3]), float(b[4]))
    assert len(calc.basisStates) == int(g["N"])
    calc.diagonalise(g["r"], int(g["k"]), drivingFromState=[int(d[0]), int(d[1]), float(d[2]), float(d[3]), int(d[4])])
    assert calc.maxCoupledStateIndex == int(g["maxCoupledStateIndex"])
    assert abs(calc.maxCoupling - float(g["maxCoupling"])) <= 1e-8 * float(g["maxCoupling"])
    y = np.array(calc.y)
    assert y.shape == g["y"].shape                      # clamped to N-1 eigenpairs
    scale = max(np.abs(g["y"]).max(), 1e-3)
    assert np.abs(y - g["y"]).max() <= 1e-7 * scale      # eigsh tol = 1e-6
    hl = np.array(calc.highlight)
    for i in range(len(g["r"])):
        cid = np.concatenate([[0], np.cumsum(np.diff(g["y"][i]) > 1e-6 * scale)])
        a_ = np.bincount(cid, weights=hl[i]); b_ = np.bincount(cid, weights=g["highlight"][i])
        assert np.abs(a_[1:-1] - b_[1:-1]).max() <= 1e-6 if len(a_) > 2 else True
