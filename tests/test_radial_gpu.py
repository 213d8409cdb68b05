"""GPU parity tests for subsystem (1): CUDA Numerov / radial integrals / Anger path
against the oracle and the golden vectors frozen from the live reference.
Tolerances (BASELINE.json north_star): matrix elements 1e-8 relative."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_ME = 1e-8


def _rb():
    import arc_b200
    return arc_b200.Rubidium()


def test_library_reports_b200():
    from arc_b200 import _lib
    dev = _lib.require_device()
    assert dev >= 0


def test_radial_elements_match_golden(gold_dir):
    g = json.load(open(os.path.join(gold_dir, "radial_golden.json")))["atoms"]
    import arc_b200
    for cls in ("Rubidium", "Caesium", "Potassium39"):
        atom = getattr(arc_b200, cls)()
        d, q = g[cls]["dipole"], g[cls]["quadrupole"]
        atom.precompute_radial(dipole_pairs=[tuple(r[:6]) for r in d],
                               quadrupole_pairs=[tuple(r[:6]) for r in q])
        for r in d:
            v = atom.getRadialMatrixElement(*r[:6])
            assert abs(v - r[6]) <= REL_ME * abs(r[6]), (cls, r, v)
        for r in q:
            v = atom.getQuadrupoleMatrixElement(*r[:6])
            assert abs(v - r[6]) <= REL_ME * abs(r[6]), (cls, r, v)
        assert atom.gpu_kernel_launches >= 2


def test_wavefunctions_match_oracle():
    """psi, r, divergence point and norm for seeded states vs the C restatement."""
    from oracle import oracle
    from arc_b200 import radial
    atom = _rb()
    rng = np.random.default_rng(5)
    sts = [(20, 0, .5), (33, 1, .5), (47, 2, 2.5), (60, 3, 3.5), (61, 4, 4.5), (80, 7, 6.5),
           (120, 20, 19.5), (25, 24, 24.5)]
    for _ in range(6):
        n = int(rng.integers(20, 100)); l = int(rng.integers(0, min(n - 1, 25)))
        sts.append((n, l, l + .5))
    recs = np.array([atom._state_record(*s) for s in sts])
    rb = radial.RadialBatch(recs, normalise=True)
    dps = rb.divergence_points()
    for i, (n, l, j) in enumerate(sts):
        rec = recs[i]
        args = [float(rec[k]) for k in ("innerLimit", "outerLimit", "step", "init1", "init2")] + \
               [int(rec["l"]), float(rec["s"]), float(rec["j"]), float(rec["stateEnergy"]),
                float(rec["alphaC"]), float(rec["alpha"]), int(rec["Z"])] + \
               [float(rec[k]) for k in ("a1", "a2", "a3", "a4", "rc", "mu")]
        psi_o, r_o, dp_o = oracle.numerov(*args)
        norm = oracle.trapezoid(psi_o ** 2, r_o)
        psi_o = psi_o / np.sqrt(norm)
        r_g, psi_g = rb.wavefunction(i)
        assert len(r_g) == len(r_o)
        assert np.array_equal(r_g, r_o), "x grid must be bit-identical"
        assert abs(int(dps[i]) - dp_o) <= 100, (n, l, j, dps[i], dp_o)
        lo = max(int(dps[i]), dp_o) + 1
        scale = np.abs(psi_o).max()
        assert np.abs(psi_g[lo:] - psi_o[lo:]).max() <= 1e-9 * scale, (n, l, j)


def test_numerov_wavefunction_signature():
    """18-argument C-extension signature: d[0] = R r (raw), d[1] = r."""
    from oracle import oracle
    import arc_b200
    a = _rb()
    E = a.getEnergy(40, 1, 1.5) / 27.211386245981
    args = (a.alphaC ** (1 / 3.0), 2.0 * 40 * 55.0, 0.001, 0.01, 0.01, 1, 0.5, 1.5, E,
            a.alphaC, a.alpha, a.Z, a.a1[1], a.a2[1], a.a3[1], a.a4[1], a.rc[1],
            (a.mass - 9.1093837139e-31) / a.mass)
    d = arc_b200.NumerovWavefunction(*args)
    psi_o, r_o, _ = oracle.numerov(*args)
    assert d.shape == (2, len(r_o)) and d.flags["C_CONTIGUOUS"]
    assert np.array_equal(d[1], r_o)
    assert np.abs(d[0] - psi_o).max() <= 1e-9 * np.abs(psi_o).max()


def test_radial_wavefunction_api_normalised():
    a = _rb()
    E = a.getEnergy(30, 0, 0.5) / 27.211386245981
    r, psi = a.radialWavefunction(0, 0.5, 0.5, E, a.alphaC ** (1 / 3.0), 2.0 * 30 * 45.0, 0.001)
    from oracle import oracle
    assert abs(oracle.trapezoid(psi ** 2, r) - 1.0) < 1e-12


def test_semiclassical_matches_golden(gold_dir):
    g = json.load(open(os.path.join(gold_dir, "divalent_golden.json")))
    import arc_b200
    atom = arc_b200.Strontium88()
    for s in (0, 1):
        atom.precompute_radial(dipole_pairs=[tuple(r[:6]) for r in g["dipole"] if r[6] == s],
                               quadrupole_pairs=[tuple(r[:6]) for r in g["quadrupole"] if r[6] == s],
                               s=s)
    for r in g["dipole"]:
        v = atom.getRadialMatrixElement(*r[:6], s=r[6])
        assert abs(v - r[7]) <= REL_ME * max(abs(r[7]), 1e-3), (r, v)
    for r in g["quadrupole"]:
        v = atom.getQuadrupoleMatrixElement(*r[:6], s=r[6])
        assert abs(v - r[7]) <= REL_ME * max(abs(r[7]), 1e-3), (r, v)


def test_empty_and_tiny_batches():
    from arc_b200 import radial
    atom = _rb()
    assert atom.precompute_radial(dipole_pairs=[]) == 0
    assert atom.getRadialMatrixElement(30, 0, .5, 31, 2, 1.5) == 0     # forbidden: dl=2
    recs = np.array([atom._state_record(21, 0, .5)])
    rb = radial.RadialBatch(recs)
    out = rb.integrals(np.zeros((0, 3), dtype=np.int32))
    assert out.shape == (0,)


def test_gram_path_matches_per_pair_kernel(gold_dir):
    """Tensor-core Gram path (opt-in, atom.fast_radial=True): golden elements to 1e-8, and
    against the exact per-pair kernel to 1e-8 relative OR 3e-12 of the Cauchy-Schwarz scale
    sqrt(<r^p>_a <r^p>_b) for strongly cancelling elements (documented in atoms.py)."""
    import arc_b200
    from arc_b200 import radial
    g = json.load(open(os.path.join(gold_dir, "radial_golden.json")))["atoms"]["Rubidium"]
    atom = arc_b200.Rubidium()
    atom.fast_radial = True
    atom.precompute_radial(dipole_pairs=[tuple(r[:6]) for r in g["dipole"]],
                           quadrupole_pairs=[tuple(r[:6]) for r in g["quadrupole"]])
    for r in g["dipole"]:
        assert abs(atom.getRadialMatrixElement(*r[:6]) - r[6]) <= REL_ME * abs(r[6])
    for r in g["quadrupole"]:
        assert abs(atom.getQuadrupoleMatrixElement(*r[:6]) - r[6]) <= REL_ME * abs(r[6])
    # dense block cross-check
    sts = [(n, l, l + .5) for l in (0, 1, 2, 9) for n in range(25, 60, 3)]
    recs = np.array([atom._state_record(*s) for s in sts])
    rb = radial.RadialBatch(recs)
    byl = {l: [i for i, s in enumerate(sts) if s[1] == l] for l in (0, 1, 2, 9)}
    blocks = [(byl[0], byl[1], 1), (byl[1], byl[2], 1), (byl[0], byl[2], 2), (byl[9], byl[9], 2)]
    out, offs, nref = rb.gram_refined_device(blocks, thresh=1e-4)
    out = out.cpu().numpy()
    mom = rb.moments_device().cpu().numpy()
    E = recs["stateEnergy"]
    for gi, (A, B, p) in enumerate(blocks):
        pr = []
        for a in A:
            for b in B:
                lo, hi = (a, b) if E[a] <= E[b] else (b, a)
                pr.append((lo, hi, p))
        ref = rb.integrals(np.array(pr, dtype=np.int32))
        got = out[offs[gi]:offs[gi] + len(pr)]
        scale = np.sqrt(np.abs(np.outer(mom[A, p - 1], mom[B, p - 1]))).ravel()
        assert np.all(np.abs(got - ref) <= 1e-8 * np.abs(ref) + 3e-12 * scale)
