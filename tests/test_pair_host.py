"""CPU tests of the PairStateInteractions host logic (channel enumeration, coupling
predicate, angular algebra, matDiagonal / matR assembly) against golden vectors from the
live reference.  Radial elements come from the oracle here (no GPU needed)."""
import os

import numpy as np
import pytest

from oracle.host_radial import patch_atom_with_oracle


def _build(g, procs=1):
    import arc_b200
    atom = patch_atom_with_oracle(arc_b200.Rubidium(), procs=procs)
    c, b = g["ctor"], g["basis"]
    calc = arc_b200.PairStateInteractions(atom, int(c[0]), int(c[1]), float(c[2]), int(c[3]),
                                          int(c[4]), float(c[5]), float(c[6]), float(c[7]),
                                          interactionsUpTo=int(g["interactionsUpTo"]))
    kw = {"Bz": float(b[5])} if len(b) > 5 else {}
    calc.defineBasis(float(b[0]), float(b[1]), int(b[2]), int(b[3]), float(b[4]), **kw)
    return calc


@pytest.mark.parametrize("name", ["pair_small_theta0_golden.npz", "pair_small_theta_bz_golden.npz",
                                  "pair_small_quad_golden.npz"])
def test_define_basis_matches_reference(gold_dir, name):
    g = np.load(os.path.join(gold_dir, name))
    calc = _build(g)
    assert len(calc.channel) == len(g["channel"])
    assert np.allclose(np.array(calc.channel)[:, :6], g["channel"][:, :6])
    assert np.abs(np.array(calc.channel)[:, 6] - g["channel"][:, 6]).max() < 1e-9
    assert np.allclose(np.array(calc.basisStates), g["basisStates"])
    assert np.array_equal(calc.index, g["index"])
    assert calc.originalPairStateIndex == int(g["originalPairStateIndex"])
    d = calc.matDiagonal.diagonal()
    assert np.abs(d - g["matDiagonal"]).max() <= 1e-9 * max(1.0, np.abs(d).max())
    for o, m in enumerate(calc.matR):
        assert m.nnz == int(g["matR%d_nnz" % o]), o
        ref = g["matR%d_probe" % o]
        got = m.dot(g["probes"])
        assert np.abs(got - ref).max() <= 1e-8 * np.abs(ref).max(), o
        if "matR%d_rows" % o in g:
            dense = np.zeros(m.shape)
            dense[g["matR%d_rows" % o], g["matR%d_cols" % o]] = g["matR%d_vals" % o]
            assert np.abs(m.toarray() - dense).max() <= 1e-8 * np.abs(dense).max()
    for o, c in enumerate(calc.coupling):
        assert c.nnz == int(g["coupling%d_nnz" % o])
        assert abs(abs(c).sum() - float(g["coupling%d_sumabs" % o])) <= 1e-8 * float(g["coupling%d_sumabs" % o])


@pytest.mark.parametrize("name", ["pair_c3_golden.npz", "pair_c5_golden.npz"])
def test_define_basis_full_size_matches_reference(gold_dir, name):
    """The C3 (N = 2363) and C5 (N = 10 384, theta = pi/4, Bz) operators at their stated size against
    the live reference, on the CPU: host logic of defineBasis with radial elements from the oracle
    (the GPU twins are tests/test_pair_gpu.py::test_pair_c3 and test_round2_gpu.py::test_c5_full_size...)."""
    g = np.load(os.path.join(gold_dir, name))
    calc = _build(g, procs=min(8, os.cpu_count() or 1))
    assert len(calc.channel) == len(g["channel"]) and len(calc.basisStates) == len(g["basisStates"])
    assert np.allclose(np.array(calc.channel)[:, :6], g["channel"][:, :6])
    assert np.abs(np.array(calc.channel)[:, 6] - g["channel"][:, 6]).max() < 1e-9
    assert np.allclose(np.array(calc.basisStates), g["basisStates"])
    assert np.array_equal(calc.index, g["index"])
    assert calc.originalPairStateIndex == int(g["originalPairStateIndex"])
    d = calc.matDiagonal.diagonal()
    assert np.abs(d - g["matDiagonal"]).max() <= 1e-9 * max(1.0, np.abs(d).max())
    m = calc.matR[0]
    assert m.nnz == int(g["matR0_nnz"])
    assert abs(abs(m).sum() - float(g["matR0_sumabs"])) <= 1e-8 * float(g["matR0_sumabs"])
    ref = g["matR0_probe"]
    assert np.abs(m.dot(g["probes"]) - ref).max() <= 1e-8 * np.abs(ref).max()
    c = calc.coupling[0]
    assert c.nnz == int(g["coupling0_nnz"])
    assert abs(abs(c).sum() - float(g["coupling0_sumabs"])) <= 1e-8 * float(g["coupling0_sumabs"])
    if name == "pair_c3_golden.npz":
        # the whole CPU path of the headline configuration at one R point: our operators + the oracle's
        # restatement of the reference's eigsh call against the reference's own eigsh output (tol 1e-6)
        # and against LAPACK on the reference's matrices
        from oracle import oracle
        k = int(g["k"])
        r, y, hl, _v = oracle.pair_diagonalise(calc.matDiagonal, calc.matR, g["r"][:1], k,
                                               calc.originalPairStateIndex)
        assert r[0] == g["r"][0]
        scale = max(np.abs(g["y_dense"]).max(), 1e-3)
        assert np.abs(y[0] - g["y"][0]).max() <= 1e-6 * scale
        assert np.abs(y[0] - g["y_dense"][0]).max() <= 1e-6 * scale
        top = int(np.argmax(g["highlight"][0]))
        assert abs(hl[0][top] - g["highlight"][0][top]) <= 1e-6


def test_phi_nonzero_rejected():
    import arc_b200
    calc = arc_b200.PairStateInteractions(arc_b200.Rubidium(), 60, 0, .5, 60, 0, .5, .5, .5)
    with pytest.raises(NotImplementedError):
        calc.defineBasis(0.3, 0.2, 1, 2, 1e9)


def test_c6_angle_pairs_host_logic(gold_dir):
    """getC6perturbatively_anglePairs (calculations_atom_pairstate.py:2065-2287), the part that
    needs no eigensolve: angular paths, radial sums, rotated interaction matrices, C6 and hopping
    C6 against the live reference (radial elements from the oracle here)."""
    import arc_b200
    from tools.gen_golden import anglepairs_cases
    g = np.load(os.path.join(gold_dir, "anglepairs_golden.npz"))
    atom = patch_atom_with_oracle(arc_b200.Rubidium())
    for name, ctor, angles, nRange, dE in anglepairs_cases():
        calc = arc_b200.PairStateInteractions(atom, *ctor)
        c6, c6hop, mats = calc.getC6perturbatively_anglePairs(angles, nRange, dE, returnInteractionMatrix=True)
        ref = g[name + "_c6"]
        assert np.abs(np.array(c6) - ref).max() <= 1e-7 * np.abs(ref).max(), name
        assert len(c6hop) == len(g[name + "_c6hop"])
        if len(c6hop):
            rh = g[name + "_c6hop"]
            assert np.abs(np.array(c6hop) - rh).max() <= 1e-7 * max(np.abs(rh).max(), np.abs(ref).max()), name
        rm = g[name + "_mats"]
        assert np.array(mats).shape == rm.shape
        assert np.abs(np.array(mats) - rm).max() <= 1e-7 * np.abs(rm).max(), name
