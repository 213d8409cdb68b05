"""GPU parity tests for StarkMap (subsystems 1+2+3 end to end) against golden vectors
frozen from the live reference.  Tolerances (BASELINE.json north_star): energies 1e-9
relative (to the spectral scale max|lambda|), highlight 1e-6, matrix elements 1e-8."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _load(gold_dir, name):
    return np.load(os.path.join(gold_dir, name))


def _dense_mat2(g):
    N = len(g["mat1_diag"])
    m = np.zeros((N, N))
    m[g["mat2_rows"], g["mat2_cols"]] = g["mat2_vals"]
    return m


def _check_sweep(calc, g, fields_idx=None):
    y = np.array(calc.y)
    hl = np.array(calc.highlight)
    gy, gh = g["y"], g["highlight"]
    if fields_idx is not None:
        gy, gh = gy[fields_idx], gh[fields_idx]
    scale = np.abs(gy).max()
    assert np.abs(y - gy).max() <= 1e-9 * scale
    # highlight: compare sums over (near-)degenerate clusters, eigenvector basis inside a
    # degenerate subspace is arbitrary
    for f in range(len(gy)):
        gaps = np.diff(gy[f]) > 1e-7 * scale
        cid = np.concatenate([[0], np.cumsum(gaps)])
        a = np.bincount(cid, weights=hl[f])
        b = np.bincount(cid, weights=gh[f])
        assert np.abs(a - b).max() <= 1e-6, f


@pytest.mark.parametrize("name,ctor", [("stark_c1_golden.npz", "Rubidium"),
                                       ("stark_small_bz_golden.npz", "Rubidium")])
def test_stark_define_basis_and_diagonalise(gold_dir, name, ctor):
    import arc_b200
    g = _load(gold_dir, name)
    atom = getattr(arc_b200, ctor)()
    calc = arc_b200.StarkMap(atom)
    a = g["args"]
    calc.defineBasis(int(a[0]), int(a[1]), float(a[2]), float(a[3]), int(a[4]), int(a[5]),
                     int(a[6]), Bz=float(g["Bz"]), s=float(g["s"]))
    assert len(calc.basisStates) == len(g["basisStates"])
    assert np.allclose(np.array(calc.basisStates), g["basisStates"])
    assert calc.indexOfCoupledState == int(g["indexOfCoupledState"])
    d = np.diag(calc.mat1)
    assert np.abs(d - g["mat1_diag"]).max() <= 1e-12 * np.abs(d).max()
    m2 = _dense_mat2(g)
    nz = m2 != 0
    assert np.array_equal(calc.mat2 != 0, nz)
    assert np.abs(calc.mat2[nz] / m2[nz] - 1).max() <= 1e-8          # matrix elements
    calc.diagonalise(g["fields"])
    _check_sweep(calc, g)
    # composition: leading coefficient of the tracked state
    f = len(g["fields"]) - 1
    k = int(np.argmax(calc.highlight[f]))
    comp = calc.composition[f][k]
    assert comp[0][1] == int(g["comp_i"][f][k][0])
    assert abs(abs(comp[0][0]) - abs(g["comp_c"][f][k][0])) <= 1e-6
    assert calc.gpu_kernel_launches > 0


def test_stark_sweep_from_golden_matrices(gold_dir):
    """subsystems 2+3 only: golden mat1/mat2 in, y/highlight out (all golden fields)."""
    from arc_b200 import sweep
    g = _load(gold_dir, "stark_rb75s_golden.npz")
    res = sweep.stark_sweep(g["mat1_diag"], _dense_mat2(g), g["fields"], int(g["indexOfCoupledState"]))
    scale = np.abs(g["y"]).max()
    assert np.abs(res.y - g["y"]).max() <= 1e-9 * scale
    k = np.argmax(g["highlight"], axis=1)
    rows = np.arange(len(k))
    assert np.abs(res.hl[rows, k] - g["highlight"][rows, k]).max() <= 1e-6
    # composition agrees with the reference's _stateComposition2 on the tracked state
    assert np.array_equal(res.comp_i[rows, k, 0], g["comp_i"][rows, k, 0])


def test_reference_polarizability_pin(gold_dir):
    """reference test/test_single_atom.py:5-14: Rb 75S polarizability = 853.2 +- 0.1."""
    import arc_b200
    pin = json.load(open(os.path.join(gold_dir, "stark_rb75s_pol.json")))
    calc = arc_b200.StarkMap(arc_b200.Rubidium())
    calc.defineBasis(75, 0, 0.5, 0.5, 70, 80, 20)
    calc.diagonalise(np.linspace(0.3, 20, 100))
    p = calc.getPolarizability(minStateContribution=0.9)
    assert abs(p - pin["reference_test_value"]) < pin["tol"]
    assert abs(p - pin["polarizability"]) < 1e-6 * abs(pin["polarizability"])


def test_stark_divalent_singlet(gold_dir):
    import arc_b200
    g = _load(gold_dir, "stark_sr88_singlet_golden.npz")
    calc = arc_b200.StarkMap(arc_b200.Strontium88())
    a = g["args"]
    calc.defineBasis(int(a[0]), int(a[1]), float(a[2]), float(a[3]), int(a[4]), int(a[5]),
                     int(a[6]), s=float(g["s"]))
    m2 = _dense_mat2(g)
    nz = m2 != 0
    assert np.abs(calc.mat2[nz] / m2[nz] - 1).max() <= 1e-8
    calc.diagonalise(g["fields"])
    _check_sweep(calc, g)


def test_syevd_matches_numpy_edge_cases():
    from arc_b200 import sweep
    rng = np.random.default_rng(11)
    for N in (1, 2, 5, 33, 70):
        M = rng.standard_normal((2, N, N)); M = M + M.transpose(0, 2, 1)
        w, v = sweep.syevd_batched(M)
        for b in range(2):
            wr = np.linalg.eigvalsh(M[b])
            assert np.abs(w[b] - wr).max() <= 1e-12 * max(1.0, np.abs(wr).max())
            assert np.abs(M[b] @ v[b] - v[b] * w[b]).max() <= 1e-11 * max(1.0, np.abs(wr).max())
