"""CPU tests of the host-side mirror of the reference interface (no GPU): angular algebra,
energies, Stark basis / matrix assembly, result containers -- against golden vectors frozen
from the live reference (tools/gen_golden.py)."""
import json
import os
import pickle

import numpy as np
import pytest

from oracle.host_radial import patch_atom_with_oracle


@pytest.fixture(scope="module")
def ang(gold_dir):
    return json.load(open(os.path.join(gold_dir, "angular_golden.json")))


def test_wigner_symbols(ang):
    from arc_b200 import CG, Wigner3j, Wigner6j
    for r in ang["CG"]:
        assert abs(CG(*r[:6]) - r[6]) < 1e-13
    for r in ang["w3j"]:
        assert abs(Wigner3j(*r[:6]) - r[6]) < 1e-13
    for r in ang["w6j"]:
        l, s, j, j1, c, l1 = r[:6]
        want = r[6]
        got = Wigner6j(l, s, j, j1, c, l1)
        if l1 - l == 2:            # the reference's shipped 6j table is empty for l1 = l+2
            assert want == 0.0
        else:
            assert abs(got - want) < 1e-13


def test_wigner_d_matrices(ang):
    from arc_b200 import WignerDmatrix
    for d in ang["wignerD"]:
        m = WignerDmatrix(d["theta"], d["phi"]).get(d["j"])
        ref = np.array(d["re"]) + 1j * np.array(d["im"])
        assert np.abs(m - ref).max() < 1e-13
    assert np.array_equal(WignerDmatrix(0, 0).get(1.5), np.eye(4))


def test_efield_angular_and_zeeman(ang):
    import arc_b200
    from arc_b200.stark import efield_angular
    for r in ang["efield"]:
        assert abs(efield_angular(*r[:6], s=r[6]) - r[7]) < 1e-13
    a = arc_b200.Rubidium()
    for r in ang["zeeman"]:
        v = a.getZeemanEnergyShift(r[0], r[1], r[2], r[3], s=r[4])
        assert abs(v - r[5]) <= 1e-13 * max(abs(r[5]), 1e-30)


def test_energies_match_reference(gold_dir):
    import arc_b200
    g = json.load(open(os.path.join(gold_dir, "radial_golden.json")))["atoms"]
    for cls in g:
        a = getattr(arc_b200, cls)()
        for n, l, j, E in g[cls]["energies_eV"]:
            assert a.getEnergy(n, l, j) == E
    d = json.load(open(os.path.join(gold_dir, "divalent_golden.json")))
    sr = arc_b200.Strontium88()
    for n, l, j, s, E in d["energies_eV"]:
        assert sr.getEnergy(n, l, j, s=s) == E
    with pytest.raises(ValueError):
        sr.getEnergy(50, 1, 1)            # spin must be explicit (div:322-327)
    with pytest.raises(ValueError):
        arc_b200.Rubidium().getEnergy(5, 5, 5.5)


def test_stark_define_basis_host_logic(gold_dir):
    """basis order, mat1 (incl. Zeeman), mat2 pattern and values; radial elements from the
    oracle so this runs without a GPU."""
    import arc_b200
    g = np.load(os.path.join(gold_dir, "stark_small_bz_golden.npz"))
    calc = arc_b200.StarkMap(patch_atom_with_oracle(arc_b200.Rubidium()))
    a = g["args"]
    calc.defineBasis(int(a[0]), int(a[1]), float(a[2]), float(a[3]), int(a[4]), int(a[5]), int(a[6]),
                     Bz=float(g["Bz"]))
    assert np.allclose(np.array(calc.basisStates), g["basisStates"])
    assert calc.indexOfCoupledState == int(g["indexOfCoupledState"])
    assert np.abs(np.diag(calc.mat1) - g["mat1_diag"]).max() <= 1e-12 * np.abs(g["mat1_diag"]).max()
    N = len(g["mat1_diag"])
    m2 = np.zeros((N, N))
    m2[g["mat2_rows"], g["mat2_cols"]] = g["mat2_vals"]
    assert np.array_equal(calc.mat2 != 0, m2 != 0)
    assert np.abs(calc.mat2 - m2).max() <= 1e-10 * np.abs(m2).max()
    assert np.array_equal(calc.mat2, calc.mat2.T)


@pytest.mark.parametrize("name,cls", [("stark_rb_lown_golden.npz", "Rubidium"),
                                      ("stark_k_lown_golden.npz", "Potassium")])
def test_stark_basis_below_ground_state_matches_reference(gold_dir, name, cls):
    """nMin < groundStateN.  The reference tests `[tn, tl, tj] in atom.extraLevels` with a list
    against a list of TUPLES (calculations_atom_single.py:749-751, alkali_atom_data.py:286-295),
    which is never true: every state below groundStateN is dropped, extra levels included.
    Golden from the live reference pins that behaviour."""
    import arc_b200
    g = np.load(os.path.join(gold_dir, name))
    atom = patch_atom_with_oracle(getattr(arc_b200, cls)())
    calc = arc_b200.StarkMap(atom)
    a = g["args"]
    calc.defineBasis(int(a[0]), int(a[1]), float(a[2]), float(a[3]), int(a[4]), int(a[5]), int(a[6]))
    got = np.array(calc.basisStates)
    assert got.shape == g["basisStates"].shape and np.allclose(got, g["basisStates"])
    assert atom.extraLevels and not (got[:, 0] < atom.groundStateN).any()
    assert calc.indexOfCoupledState == int(g["indexOfCoupledState"])
    assert np.abs(np.diag(calc.mat1) - g["mat1_diag"]).max() <= 1e-12 * np.abs(g["mat1_diag"]).max()
    N = len(g["mat1_diag"])
    m2 = np.zeros((N, N))
    m2[g["mat2_rows"], g["mat2_cols"]] = g["mat2_vals"]
    assert np.array_equal(calc.mat2 != 0, m2 != 0)
    assert np.abs(calc.mat2 - m2).max() <= 1e-8 * np.abs(m2).max()


@pytest.mark.parametrize("name,N", [("stark_c4_singlet_golden.npz", 651), ("stark_c4_triplet_golden.npz", 1932)])
def test_stark_c4_full_size_host_logic(gold_dir, name, N):
    """BASELINE configs[3] at its stated size on the CPU: Sr88 basis, `mat1`, `mat2` against the live
    reference with the semiclassical radial elements from the oracle (mpmath Anger functions); the
    GPU twin is tests/test_round2_gpu.py::test_c4_sr88_full_size_matches_reference."""
    import arc_b200
    from oracle.host_radial import patch_divalent_with_oracle
    g = np.load(os.path.join(gold_dir, name))
    calc = arc_b200.StarkMap(patch_divalent_with_oracle(arc_b200.Strontium88(), procs=min(8, os.cpu_count() or 1)))
    a = g["args"]
    calc.defineBasis(int(a[0]), int(a[1]), float(a[2]), float(a[3]), int(a[4]), int(a[5]), int(a[6]),
                     s=float(g["s"]))
    assert len(calc.basisStates) == N == len(g["basisStates"])
    assert np.allclose(np.array(calc.basisStates), g["basisStates"])
    assert calc.indexOfCoupledState == int(g["indexOfCoupledState"])
    d = np.diag(calc.mat1)
    assert np.abs(d - g["mat1_diag"]).max() <= 1e-12 * np.abs(d).max()
    m2 = np.zeros((N, N))
    m2[g["mat2_rows"], g["mat2_cols"]] = g["mat2_vals"]
    # dipole-forbidden triplet pairs carry the reference's Clebsch-Gordan rounding noise instead of an
    # exact zero (divalent_atom_functions.py:405): compared absolutely below 1e-13 of the scale
    floor = 1e-13 * np.abs(m2).max()
    nz = np.abs(m2) > floor
    assert np.array_equal(np.abs(calc.mat2) > floor, nz)
    assert np.abs(calc.mat2[nz] / m2[nz] - 1).max() <= 1e-10
    assert np.abs(calc.mat2[~nz] - m2[~nz]).max() <= floor


def test_composition_view_and_pickle():
    from arc_b200.stark import CompositionView
    c = np.array([[[0.9, -0.3, 0.0], [1.0, 0.0, 0.0]]])
    i = np.array([[[4, 2, -1], [7, -1, -1]]], dtype=np.int32)
    v = CompositionView(c, i)
    assert len(v) == 1
    assert v[0][0] == [[0.9, 4], [-0.3, 2]] and v[0][1] == [[1.0, 7]]
    assert v[0][1][0][1] == 7                       # the access pattern of pair:4766-4771
    v2 = pickle.loads(pickle.dumps(v))
    assert v2[0] == v[0]
    import arc_b200
    atom = pickle.loads(pickle.dumps(arc_b200.Rubidium()))       # saveCalculation-style
    assert atom.getEnergy(60, 0, .5) == arc_b200.Rubidium().getEnergy(60, 0, .5)


def test_selection_rules_return_zero_without_compute():
    import arc_b200
    a = arc_b200.Rubidium()
    assert a.getRadialMatrixElement(30, 0, .5, 31, 2, 1.5) == 0
    assert a.getQuadrupoleMatrixElement(30, 0, .5, 31, 3, 2.5) == 0
    assert a.getRadialCoupling(30, 0, .5, 31, 4, 3.5) == 0
    # literature value (rubidium_literature_dme.csv) is returned without any computation
    assert a.getRadialMatrixElement(5, 0, .5, 5, 1, .5) == a._lit[(5, 0, 1, 5, 1, 1)][0]
    assert a.gpu_kernel_launches == 0


def test_memo_persistence_reference_formats(tmp_path):
    """Radial memo <-> the reference's own table formats (SURVEY.md 8b seam B2, 8f rank 4):
    `.npy` rows (n1,l1,j1_x2,n2,l2,j2_x2,value) and the sqlite tables dipoleME / quadrupoleME
    (alkali_atom_functions.py:246-335, 1860-1906); divalent keys carry j and s
    (divalent_atom_functions.py:164-171).  No GPU: the memo is filled by hand."""
    import sqlite3
    import arc_b200
    a = arc_b200.Rubidium()
    a._dipole[(60, 0, 1, 60, 1, 3)] = 4131.5
    a._dipole[(59, 1, 1, 60, 0, 1)] = -3820.25
    a._quadrupole[(60, 0, 1, 61, 2, 5)] = 1.25e6
    d, q = str(tmp_path / "rb_dipole_matrix_elements.npy"), str(tmp_path / "rb_quadrupole_matrix_elements.npy")
    a.saveMatrixElementTables(d, q)
    rows = np.load(d)
    assert rows.shape == (2, 7) and rows.dtype == np.float64
    b = arc_b200.Rubidium()
    assert b.loadMatrixElementTables(d, q) == 3
    assert b._dipole == a._dipole and b._quadrupole == a._quadrupole
    # element present in the memo: answered without any GPU call (no device in this test)
    assert b.getRadialMatrixElement(60, 0, 0.5, 60, 1, 1.5) == 4131.5
    assert b.getRadialMatrixElement(60, 1, 1.5, 60, 0, 0.5) == 4131.5
    db = str(tmp_path / "precalculated_pair.db")
    assert a.exportToArcDatabase(db) == (2, 1)
    conn = sqlite3.connect(db)
    c = conn.cursor()
    # the reference's own lookup (alkali_atom_functions.py:1057-1062)
    c.execute("SELECT dme FROM dipoleME WHERE n1= ? AND l1 = ? AND j1_x2 = ? AND n2 = ? AND l2 = ? AND j2_x2 = ?",
              (60, 0, 1, 60, 1, 3))
    assert c.fetchone()[0] == 4131.5
    c.execute("SELECT qme FROM quadrupoleME WHERE n1= ? AND l1 = ? AND j1_x2 = ? AND n2 = ? AND l2 = ? AND j2_x2 = ?",
              (60, 0, 1, 61, 2, 5))
    assert c.fetchone()[0] == 1.25e6
    conn.close()
    e = arc_b200.Rubidium()
    assert e.importFromArcDatabase(db) == 3 and e._dipole == a._dipole
    sr = arc_b200.Strontium88()
    sr._dipole[(60, 0, 0, 60, 1, 1, 0)] = 3000.0
    db2 = str(tmp_path / "sr.db")
    assert sr.exportToArcDatabase(db2) == (1, 0)
    conn = sqlite3.connect(db2)
    assert conn.execute("SELECT dme FROM dipoleME WHERE n1=60 AND l1=0 AND j1=0 AND n2=60 AND l2=1 AND j2=1 AND s=0").fetchone()[0] == 3000.0
    conn.close()
