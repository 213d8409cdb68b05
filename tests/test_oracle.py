"""CPU tests that pin the oracle (the checker) before it is trusted.

* bit-for-bit against the reference's own C extension compiled into oracle/_ref
  (skipped on boxes without it);
* against golden vectors frozen from the LIVE reference (tools/gen_golden.py).
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle

RB = dict(alphaC=9.076, alpha=0.0072973525643, Z=37,
          a1=[3.69628474, 4.44088978, 3.78717363, 2.39848933],
          a2=[1.64915255, 1.92828831, 1.57027864, 1.76810544],
          a3=[-9.86069196, -16.7959777, -11.6558897, -12.0710678],
          a4=[0.19579987, -0.8163314, 0.52942835, 0.77256589],
          rc=[1.66242117, 1.50195124, 4.86851938, 4.79831327],
          mu=(1.4099934427170326e-25 - 9.1093837139e-31) / 1.4099934427170326e-25)


def rb_state(n, l, j, E_au):
    sp = dict(n=n, l=l, j=j, E_au=E_au, alphaC=RB["alphaC"], alpha=RB["alpha"], Z=RB["Z"],
              mu=RB["mu"])
    for k in ("a1", "a2", "a3", "a4", "rc"):
        sp[k] = RB[k][l] if l < 4 else 0.0
    return sp


def _numerov_args(sp, step=0.001):
    inner = max(4 * step, sp["alphaC"] ** (1 / 3.0))
    outer = 2.0 * sp["n"] * (sp["n"] + 15.0)
    return (inner, outer, step, 0.01, 0.01, sp["l"], 0.5, sp["j"], sp["E_au"], sp["alphaC"],
            sp["alpha"], sp["Z"], sp["a1"], sp["a2"], sp["a3"], sp["a4"], sp["rc"], sp["mu"])


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("n,l,j", [(20, 0, .5), (45, 1, 1.5), (60, 3, 2.5), (61, 4, 4.5),
                                   (60, 10, 10.5), (100, 20, 19.5)])
def test_restatement_matches_reference_build(n, l, j):
    """oracle/numerov_oracle.c == arc_c_extensions.c:108-285, bit for bit."""
    E = -0.5 / (n - (3.13 if l == 0 else 0.0)) ** 2
    args = _numerov_args(rb_state(n, l, j, E))
    psi, r, dp = oracle.numerov(*args)
    psi_ref, r_ref = oracle.ref_numerov(*args)
    assert len(psi) == len(psi_ref)
    assert np.array_equal(r, r_ref)
    assert np.array_equal(psi, psi_ref)


def test_oracle_radial_elements_match_golden(gold_dir):
    """oracle radial dipole/quadrupole elements == live reference values (<=1e-12)."""
    g = json.load(open(os.path.join(gold_dir, "radial_golden.json")))["atoms"]["Rubidium"]
    en = {}
    import arc_b200
    atom = arc_b200.Rubidium()

    def sp(n, l, j):
        return rb_state(n, l, j, atom.getEnergy(n, l, j) / 27.211386245981)
    for rows, power in ((g["dipole"][:12], 1), (g["quadrupole"][:6], 2)):
        for n1, l1, j1, n2, l2, j2, val in rows:
            a, b = sp(int(n1), int(l1), j1), sp(int(n2), int(l2), j2)
            if a["E_au"] > b["E_au"]:
                a, b = b, a
            v = oracle.radial_integral(a, b, power)
            assert abs(v - val) <= 1e-12 * max(1.0, abs(val)), (n1, l1, j1, n2, l2, j2, v, val)


def test_oracle_wavefunction_pins(gold_dir):
    g = json.load(open(os.path.join(gold_dir, "radial_golden.json")))["atoms"]["Rubidium"]
    for w in g["wavefunctions"]:
        r, psi = oracle.radial_wavefunction(rb_state(w["n"], w["l"], w["j"], w["E_au"]))
        assert len(r) == w["L"]
        assert int(np.argmax(psi != 0)) == w["first_nonzero"]
        np.testing.assert_allclose(r[w["idx"]], w["r"], rtol=0, atol=0)
        np.testing.assert_allclose(psi[w["idx"]], w["psi"], rtol=1e-13, atol=1e-300)


def test_oracle_semiclassical_matches_golden(gold_dir):
    g = json.load(open(os.path.join(gold_dir, "divalent_golden.json")))
    import arc_b200
    atom = arc_b200.Strontium88()
    for n1, l1, j1, n2, l2, j2, s, val in g["dipole"][:20]:
        key = atom._ordered(int(n1), int(l1), j1, int(n2), int(l2), j2, s)
        p = atom._semiclassical_params([key], 1, s)[0]
        v = oracle.semiclassical_dipole(*p)
        assert abs(v - val) <= 1e-10 * max(1.0, abs(val))
    for n1, l1, j1, n2, l2, j2, s, val in g["quadrupole"][:20]:
        key = atom._ordered(int(n1), int(l1), j1, int(n2), int(l2), j2, s)
        p = atom._semiclassical_params([key], 2, s)[0]
        v = oracle.semiclassical_quadrupole(*p)
        assert abs(v - val) <= 1e-10 * max(1.0, abs(val))


def test_oracle_stark_matches_golden(gold_dir):
    """numpy eigh on the golden mat1/mat2 reproduces the live reference's y/highlight."""
    g = np.load(os.path.join(gold_dir, "stark_c1_golden.npz"))
    N = len(g["mat1_diag"])
    mat1 = np.diag(g["mat1_diag"])
    mat2 = np.zeros((N, N))
    mat2[g["mat2_rows"], g["mat2_cols"]] = g["mat2_vals"]
    fields = g["fields"][[0, 6, 12]]
    y, hl, comp = oracle.stark_diagonalise(mat1, mat2, fields, int(g["indexOfCoupledState"]))
    np.testing.assert_allclose(y, g["y"][[0, 6, 12]], rtol=0, atol=1e-9)
    np.testing.assert_allclose(hl, g["highlight"][[0, 6, 12]], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name,pick", [("stark_c4_singlet_golden.npz", [0, 3, 7]), ("stark_c4_triplet_golden.npz", [5])])
def test_oracle_stark_c4_matches_golden(gold_dir, name, pick):
    """The oracle's sweep restatement on the Sr88 matrices of BASELINE configs[3] (many exactly equal
    diagonal entries): eigenvalues against the live reference, highlight as sums over degenerate clusters."""
    g = np.load(os.path.join(gold_dir, name))
    N = len(g["mat1_diag"])
    mat2 = np.zeros((N, N))
    mat2[g["mat2_rows"], g["mat2_cols"]] = g["mat2_vals"]
    y, hl, _ = oracle.stark_diagonalise(np.diag(g["mat1_diag"]), mat2, g["fields"][pick],
                                        int(g["indexOfCoupledState"]), want_composition=False)
    scale = np.abs(g["y"]).max()
    assert np.abs(y - g["y"][pick]).max() <= 1e-9 * scale
    for q, f in enumerate(pick):
        edges = np.flatnonzero(np.diff(g["y"][f]) > 1e-7 * scale) + 1
        A = np.add.reduceat(hl[q], np.r_[0, edges])
        B = np.add.reduceat(g["highlight"][f], np.r_[0, edges])
        assert np.abs(A - B).max() <= 1e-6


def test_oracle_numerov_back_matches_live_reference(gold_dir):
    """Restatement of the pure-Python NumerovBack (aaf:3939-4063) against outputs of the live
    reference for analytic callables: bit-exact, including the divergence cut."""
    from tools.gen_golden import numerovback_cases
    g = np.load(os.path.join(gold_dir, "numerovback_golden.npz"))
    for name, inner, outer, k, step, i1, i2 in numerovback_cases():
        r, psi, dp = oracle.numerov_back(inner, outer, k, step, i1, i2)
        assert len(r) == int(g[name + "_L"])
        assert np.array_equal(r[::7], g[name + "_r"]) and np.array_equal(psi[::7], g[name + "_psi"])
        assert np.array_equal(psi[:400], g[name + "_head_psi"]) and np.array_equal(r[:400], g[name + "_head_r"])
        assert float((psi == 0).sum()) == g[name + "_sum"][3]
