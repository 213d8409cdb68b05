"""GPU parity tests for ShirleyMethod (Floquet Stark maps) against golden
vectors frozen from the live reference (tools/gen_golden.py shirley) and against numpy.linalg.eigh
on the same Floquet matrices (calculations_atom_single.py:3931-3957).
Tolerances: matrix elements 1e-8, energies 1e-9 of the spectral scale, probabilities 1e-6."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(g, k):
    import arc_b200
    kw = g[k + "_kw"]
    calc = arc_b200.ShirleyMethod(arc_b200.Rubidium())
    b = g[k + "_basis"]
    calc.defineBasis(int(b[0]), int(b[1]), float(b[2]), float(b[3]), q=int(kw[0]), nMin=int(kw[1]),
                     nMax=int(kw[2]), maxL=int(kw[3]), edN=int(kw[4]), Bz=float(kw[5]))
    calc.defineShirleyHamiltonian(fn=int(kw[6]))
    return calc


@pytest.mark.parametrize("k", ["a", "b"])
def test_shirley_against_reference_golden(gold_dir, k):
    g = np.load(os.path.join(gold_dir, "shirley_golden.npz"))
    calc = _case(g, k)
    assert np.allclose(np.array(calc.basisStates, dtype=float), g[k + "_basisStates"])
    assert calc.indexOfCoupledState == int(g[k + "_index"])
    E = g[k + "_bareEnergies"]
    assert np.abs(calc.bareEnergies - E).max() <= 1e-12 * np.abs(E).max()
    V = g[k + "_V"]
    nz = V != 0
    assert np.array_equal(calc.V != 0, nz)
    assert np.abs(calc.V[nz] / V[nz] - 1).max() <= 1e-8                     # matrix elements
    calc.diagonalise(g[k + "_eFields"], g[k + "_freqs"])
    eigs, tp, ts = g[k + "_eigs"], g[k + "_transProbs"], g[k + "_targetShifts"]
    assert calc.eigs.shape == eigs.shape and calc.transProbs.shape == tp.shape
    scale = np.abs(eigs).max()
    assert np.abs(calc.eigs - eigs).max() <= 1e-9 * scale
    # probabilities and shifts follow the eigenvector with the largest target overlap; compare
    # where that choice is unambiguous in the reference
    assert np.abs(calc.transProbs.sum(axis=-1) - tp.sum(axis=-1)).max() <= 1e-6
    assert np.abs(calc.transProbs - tp).max() <= 1e-6
    assert np.abs(calc.targetShifts - ts).max() <= 1e-9 * scale + 1e-6 * np.abs(ts).max()


def test_shirley_two_stage_window_against_numpy():
    """Floquet dimension inside the two-stage window (3 x 352 = 1056): eigenvalues, transition
    probabilities and target shifts against numpy.linalg.eigh on the same matrices."""
    import arc_b200
    calc = arc_b200.ShirleyMethod(arc_b200.Rubidium())
    calc.defineBasis(40, 0, 0.5, 0.5, q=0, nMin=36, nMax=46, maxL=16, edN=0)
    dim0 = len(calc.basisStates)
    calc.defineShirleyHamiltonian(fn=1)
    N = 3 * dim0
    assert 1024 <= N <= 4096, N
    eF, fr = [30.0, 120.0], [4.0e9]
    calc.diagonalise(eF, fr)
    tar = calc.indexOfCoupledState + dim0
    for ei, f in enumerate(eF):
        Hf = (calc.H0 + calc.dT * fr[0] + calc.B * f).toarray()
        ev, vec = np.linalg.eigh(Hf)
        scale = np.abs(ev).max()
        assert np.abs(calc.eigs[ei] - ev).max() <= 1e-9 * scale
        tp = (np.abs(vec * vec[tar]) ** 2).reshape((3, dim0, N)).sum(axis=(0, -1))
        assert np.abs(calc.transProbs[ei] - tp).max() <= 1e-6
        evInd = int(np.argmax(np.abs(vec[tar]) ** 2))
        assert abs(calc.targetShifts[ei] - (calc.targetEnergy - ev[evInd])) <= 1e-9 * scale


def test_shirley_keep_eigenvectors():
    import arc_b200
    calc = arc_b200.ShirleyMethod(arc_b200.Rubidium())
    calc.defineBasis(30, 0, 0.5, 0.5, q=0, nMin=29, nMax=31, maxL=2, edN=0)
    calc.defineShirleyHamiltonian(fn=1)
    calc.diagonalise([80.0], [2.0e9], keepEigenvectors=True)
    N = 3 * len(calc.basisStates)
    Hf = (calc.H0 + calc.dT * 2.0e9 + calc.B * 80.0).toarray()
    v = np.real(calc.eigVectors)
    assert v.shape == (N, N)
    assert np.abs(Hf @ v - v * calc.eigs).max() <= 1e-9 * np.abs(calc.eigs).max()
