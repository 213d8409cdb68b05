"""Freeze golden vectors from the LIVE reference (ARC 3.10.2 CPU path) into tests/golden/.

Runs only in the build container (needs /root/reference); the fixtures it writes
are committed and are what the GPU box compares against.  The reference is run
with an EMPTY matrix-element memo so every radial element comes from its live C
Numerov path (SURVEY.md 8c oracle rules).

    python tools/gen_golden.py [radial] [angular] [stark] [pair] [divalent]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tools.refenv as refenv  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def _comp_arrays(comp, width):
    """list[list[[coef,idx]...]] -> (coef[n,width], idx[n,width]) padded with idx=-1."""
    n = len(comp)
    c = np.zeros((n, width))
    i = -np.ones((n, width), dtype=np.int32)
    for a, entries in enumerate(comp):
        for b, (cf, ix) in enumerate(entries[:width]):
            c[a, b] = float(np.real(cf))
            i[a, b] = int(ix)
    return c, i


def gen_radial(arc):
    rng = np.random.default_rng(0)
    out = {"atoms": {}}
    for cls, npairs in (("Rubidium", 70), ("Caesium", 12), ("Potassium39", 12)):
        a = getattr(arc, cls)()
        dip, quad, wf = [], [], []
        fixed_d = [(60, 0, .5, 60, 1, 1.5), (60, 0, .5, 70, 1, 1.5), (28, 0, .5, 28, 1, .5),
                   (20, 0, .5, 120, 1, 1.5), (120, 19, 18.5, 120, 20, 19.5),
                   (60, 3, 3.5, 61, 4, 4.5), (60, 10, 10.5, 59, 11, 11.5),
                   (25, 2, 2.5, 26, 3, 3.5)]
        fixed_q = [(60, 0, .5, 59, 2, 2.5), (60, 1, 1.5, 60, 1, .5), (60, 2, 2.5, 62, 4, 4.5),
                   (100, 20, 20.5, 100, 20, 19.5), (40, 3, 3.5, 40, 5, 4.5)]
        pairs_d = list(fixed_d) if cls == "Rubidium" else []
        pairs_q = list(fixed_q) if cls == "Rubidium" else []
        while len(pairs_d) < npairs:
            n1, n2 = int(rng.integers(20, 121)), int(rng.integers(20, 121))
            l1 = int(rng.integers(0, min(20, n1 - 1) + 1))
            l2 = l1 + int(rng.choice([-1, 1]))
            if l2 < 0 or l2 > min(20, n2 - 1):
                continue
            j1 = l1 + float(rng.choice([-.5, .5]))
            j2 = l2 + float(rng.choice([-.5, .5]))
            if j1 < 0 or j2 < 0 or abs(j1 - j2) > 1.1:
                continue
            pairs_d.append((n1, l1, j1, n2, l2, j2))
        while len(pairs_q) < npairs // 2:
            n1, n2 = int(rng.integers(20, 121)), int(rng.integers(20, 121))
            l1 = int(rng.integers(0, min(20, n1 - 1) + 1))
            l2 = l1 + int(rng.choice([-2, 0, 2]))
            if l2 < 0 or l2 > min(20, n2 - 1):
                continue
            j1 = l1 + float(rng.choice([-.5, .5]))
            j2 = l2 + float(rng.choice([-.5, .5]))
            if j1 < 0 or j2 < 0 or abs(j1 - j2) > 2.1 or (n1, l1, j1) == (n2, l2, j2):
                continue
            pairs_q.append((n1, l1, j1, n2, l2, j2))
        t0 = time.time()
        for p in pairs_d:
            dip.append(list(p) + [float(a.getRadialMatrixElement(*p))])
        for p in pairs_q:
            quad.append(list(p) + [float(a.getQuadrupoleMatrixElement(*p))])
        # wavefunction-level pins: length, value checksums
        for (n, l, j) in [(20, 0, .5), (60, 1, 1.5), (60, 10, 10.5), (120, 20, 19.5), (35, 4, 4.5)]:
            E = a.getEnergy(n, l, j) / 27.211386245981
            r, psi = a.radialWavefunction(l, 0.5, j, E, a.alphaC ** (1 / 3.0),
                                          2.0 * n * (n + 15.0), 0.001)
            nz = int(np.argmax(psi != 0))
            idx = [nz, nz + 1, len(r) // 3, len(r) // 2, len(r) - 2, len(r) - 1]
            wf.append({"n": n, "l": l, "j": j, "E_au": E, "L": int(len(r)),
                       "first_nonzero": nz,
                       "idx": idx, "r": [float(r[i]) for i in idx],
                       "psi": [float(psi[i]) for i in idx],
                       "sum_abs_psi": float(np.abs(psi).sum()),
                       "sum_psi2_r": float((psi * psi * r).sum())})
        energies = []
        for n in (8, 9, 20, 28, 60, 120):
            for l in range(0, min(n, 8)):
                for j in (l - .5, l + .5):
                    if j > 0:
                        energies.append([n, l, j, float(a.getEnergy(n, l, j))])
        out["atoms"][cls] = {"dipole": dip, "quadrupole": quad, "wavefunctions": wf,
                             "energies_eV": energies}
        print(cls, "radial golden", len(dip), len(quad), "%.1fs" % (time.time() - t0))
    with open(os.path.join(GOLD, "radial_golden.json"), "w") as f:
        json.dump(out, f, indent=0)


def gen_angular(arc):
    from arc.wigner import CG, Wigner3j, Wigner6j, WignerDmatrix
    from arc.alkali_atom_functions import _EFieldCoupling
    rng = np.random.default_rng(1)
    out = {"CG": [], "w3j": [], "w6j": [], "efield": [], "zeeman": [], "wignerD": []}
    for _ in range(300):
        j1 = int(rng.integers(0, 12)) / 2
        j2 = int(rng.integers(0, 6)) / 2
        j3 = abs(j1 - j2) + int(rng.integers(0, int(2 * min(j1, j2)) + 1))
        m1 = -j1 + int(rng.integers(0, int(2 * j1) + 1))
        m2 = -j2 + int(rng.integers(0, int(2 * j2) + 1))
        m3 = m1 + m2
        if abs(m3) > j3:
            continue
        out["CG"].append([j1, m1, j2, m2, j3, m3, float(CG(j1, m1, j2, m2, j3, m3))])
        out["w3j"].append([j1, j2, j3, m1, m2, -m3, float(Wigner3j(j1, j2, j3, m1, m2, -m3))])
    for _ in range(200):
        l = int(rng.integers(0, 8)); s = .5
        j = l + float(rng.choice([-.5, .5]))
        c = int(rng.integers(1, 3))
        l1 = l + int(rng.integers(-c, c + 1))
        j1 = l1 + float(rng.choice([-.5, .5]))
        if j < 0 or j1 < 0 or l1 < 0 or abs(j - j1) > c:
            continue
        try:
            out["w6j"].append([l, s, j, j1, c, l1, float(Wigner6j(l, s, j, j1, c, l1))])
        except Exception:
            pass
    ef = _EFieldCoupling()
    for s in (0.5, 0, 1):
        for _ in range(60):
            l1 = int(rng.integers(0, 10))
            l2 = l1 + int(rng.choice([-1, 1]))
            if l2 < 0:
                continue
            j1 = l1 + float(rng.choice(np.arange(-s, s + .1, 1)))
            j2 = l2 + float(rng.choice(np.arange(-s, s + .1, 1)))
            if j1 < 0 or j2 < 0:
                continue
            mj = -min(j1, j2) + int(rng.integers(0, int(2 * min(j1, j2)) + 1))
            out["efield"].append([l1, j1, mj, l2, j2, mj, s,
                                  float(ef.getAngular(l1, j1, mj, l2, j2, mj, s=s))])
    a = arc.Rubidium()
    for _ in range(40):
        l = int(rng.integers(0, 6)); j = l + float(rng.choice([-.5, .5]))
        if j < 0:
            continue
        mj = -j + int(rng.integers(0, int(2 * j) + 1))
        out["zeeman"].append([l, j, mj, 1e-3, 0.5, float(a.getZeemanEnergyShift(l, j, mj, 1e-3))])
    for theta, phi in ((0.7, 0.0), (np.pi / 4, 0.0), (1.3, 0.4)):
        wgd = WignerDmatrix(theta, phi)
        for j in (0.5, 1.5, 2.5, 4.5, 1.0, 2.0):
            m = wgd.get(j)
            m = np.asarray(m.todense() if hasattr(m, "todense") else m)
            out["wignerD"].append({"theta": theta, "phi": phi, "j": j,
                                   "re": np.real(m).tolist(), "im": np.imag(m).tolist()})
    with open(os.path.join(GOLD, "angular_golden.json"), "w") as f:
        json.dump(out, f)
    print("angular golden", {k: len(v) for k, v in out.items()})


def _stark_case(arc, atom, args, fields, name, kwargs=None, pol=None):
    kwargs = kwargs or {}
    t0 = time.time()
    c = arc.StarkMap(atom)
    c.defineBasis(*args, **kwargs)
    t1 = time.time()
    c.diagonalise(fields)
    t2 = time.time()
    N = len(c.basisStates)
    cc, ci = [], []
    for fi in range(len(fields)):
        a, b = _comp_arrays(c.composition[fi], 3)
        cc.append(a); ci.append(b)
    nzr, nzc = np.nonzero(c.mat2)
    d = dict(args=np.array(args, dtype=float), s=kwargs.get("s", 0.5), Bz=kwargs.get("Bz", 0.0),
             basisStates=np.array(c.basisStates, dtype=float),
             indexOfCoupledState=c.indexOfCoupledState,
             mat1_diag=np.diag(c.mat1).copy(), mat2_rows=nzr.astype(np.int32),
             mat2_cols=nzc.astype(np.int32), mat2_vals=c.mat2[nzr, nzc],
             fields=np.array(fields, dtype=float), y=np.array(c.y),
             highlight=np.array(c.highlight), comp_c=np.array(cc), comp_i=np.array(ci),
             t_defineBasis=t1 - t0, t_diagonalise=t2 - t1)
    if pol is not None:
        d["polarizability"] = float(c.getPolarizability(**pol))
    np.savez_compressed(os.path.join(GOLD, name), **d)
    print(name, "N", N, "defineBasis %.1fs diagonalise %.1fs" % (t1 - t0, t2 - t1),
          d.get("polarizability"))


def gen_stark(arc):
    a = arc.Rubidium()
    # C1 (BASELINE.json configs[0]); 13 of the 600 fields kept (indices 0,50,...,550,599)
    F = np.linspace(0, 600, 600)
    _stark_case(arc, a, (28, 0, 0.5, 0.5, 23, 32, 20), F[list(range(0, 600, 50)) + [599]],
                "stark_c1_golden.npz")
    # reference test/test_single_atom.py:5-14  (853.2 +- 0.1)
    _stark_case(arc, a, (75, 0, 0.5, 0.5, 70, 80, 20), np.linspace(0.3, 20, 100)[::9],
                "stark_rb75s_golden.npz")
    # small cases: Bz != 0, |mj|=1.5
    _stark_case(arc, a, (30, 2, 2.5, 1.5, 28, 32, 8), np.linspace(0, 2000, 5),
                "stark_small_bz_golden.npz", kwargs={"Bz": 1e-3})


def gen_stark_full_pol(arc):
    """Full test_single_atom run for the polarizability pin (y of tracked state only)."""
    a = arc.Rubidium()
    c = arc.StarkMap(a)
    c.defineBasis(75, 0, 0.5, 0.5, 70, 80, 20)
    F = np.linspace(0.3, 20, 100)
    c.diagonalise(F)
    p = c.getPolarizability(minStateContribution=0.9)
    with open(os.path.join(GOLD, "stark_rb75s_pol.json"), "w") as f:
        json.dump({"polarizability": float(p), "reference_test_value": 853.2, "tol": 0.1}, f)
    print("Rb75S polarizability", p)


def _pair_case(arc, atom, ctor, basis, Rs, k, name, kwargs=None):
    kwargs = kwargs or {}
    t0 = time.time()
    p = arc.PairStateInteractions(atom, *ctor, **kwargs)
    p.defineBasis(*basis[:5], **({"Bz": basis[5]} if len(basis) > 5 else {}))
    t1 = time.time()
    N = len(p.basisStates)
    p.diagonalise(np.array(Rs, dtype=float), k)
    t2 = time.time()
    rng = np.random.default_rng(7)
    probes = rng.standard_normal((N, 3))
    d = dict(ctor=np.array(ctor, dtype=float), basis=np.array(basis, dtype=float),
             interactionsUpTo=kwargs.get("interactionsUpTo", 1),
             channel=np.array(p.channel, dtype=float),
             basisStates=np.array(p.basisStates, dtype=float),
             index=np.array(p.index, dtype=np.int32),
             originalPairStateIndex=p.originalPairStateIndex,
             matDiagonal=np.asarray(p.matDiagonal.diagonal()),
             probes=probes, r=np.array(p.r), y=np.array(p.y),
             highlight=np.array(p.highlight), k=k,
             t_defineBasis=t1 - t0, t_diagonalise=t2 - t1)
    for o, m in enumerate(p.matR):
        if np.iscomplexobj(m.data):
            raise RuntimeError("complex matR not supported in golden")
        d["matR%d_nnz" % o] = m.nnz
        d["matR%d_sumabs" % o] = float(abs(m).sum())
        d["matR%d_probe" % o] = m.dot(probes)
        if N <= 800:
            coo = m.tocoo()
            d["matR%d_rows" % o] = coo.row.astype(np.int32)
            d["matR%d_cols" % o] = coo.col.astype(np.int32)
            d["matR%d_vals" % o] = coo.data
    for o, c in enumerate(p.coupling):
        d["coupling%d_nnz" % o] = c.nnz
        d["coupling%d_sumabs" % o] = float(abs(c).sum())
    # dense cross-check of eigsh (SURVEY 8c iv): k nearest sigma=0 from numpy eigh
    yd = []
    for rval in p.r:
        m = p.matDiagonal
        rX = (rval * 1e-6) ** 3
        for mr in p.matR:
            m = m + mr / rX
            rX *= rval * 1e-6
        w = np.linalg.eigvalsh(m.toarray())
        sel = np.sort(np.argsort(np.abs(w), kind="stable")[:min(k, N - 1)])
        yd.append(w[sel])
    d["y_dense"] = np.array(yd)
    np.savez_compressed(os.path.join(GOLD, name), **d)
    print(name, "N", N, "channels", len(p.channel), "nnz", p.matR[0].nnz,
          "defineBasis %.1fs diagonalise %.1fs" % (t1 - t0, t2 - t1))


def gen_pair(arc, big=True):
    a = arc.Rubidium()
    # small cases first (fast): theta=0; theta!=0 with Bz; quadrupole terms
    _pair_case(arc, a, (60, 0, .5, 60, 0, .5, .5, .5), (0, 0, 2, 3, 8e9),
               [1.5, 3.0, 8.0], 40, "pair_small_theta0_golden.npz")
    _pair_case(arc, a, (40, 2, 2.5, 40, 2, 2.5, 2.5, 2.5), (np.pi / 4, 0, 1, 3, 6e9, 1e-4),
               [2.0, 4.0, 9.0], 60, "pair_small_theta_bz_golden.npz")
    _pair_case(arc, a, (50, 1, 1.5, 50, 1, 1.5, 1.5, .5), (0.3, 0, 1, 3, 5e9),
               [1.2, 5.0], 30, "pair_small_quad_golden.npz", kwargs={"interactionsUpTo": 2})
    if big:
        # C3 (BASELINE.json configs[2]); 3 of the 1000 R points + 3 more
        _pair_case(arc, a, (60, 0, .5, 60, 0, .5, .5, .5), (0, 0, 5, 5, 25e9),
                   [1.0, 2.5, 4.0, 5.5, 7.75, 10.0], 250, "pair_c3_golden.npz")


def gen_divalent(arc):
    a = arc.Strontium88()
    rng = np.random.default_rng(3)
    out = {"dipole": [], "quadrupole": [], "energies_eV": []}
    for s in (0, 1):
        cnt = 0
        while cnt < 40:
            n1, n2 = int(rng.integers(40, 75)), int(rng.integers(40, 75))
            l1 = int(rng.integers(0, 25))
            l2 = l1 + int(rng.choice([-1, 1]))
            if l2 < 0 or l1 >= n1 or l2 >= n2:
                continue
            j1 = l1 + int(rng.integers(-s, s + 1)); j2 = l2 + int(rng.integers(-s, s + 1))
            if j1 < 0 or j2 < 0 or (s == 1 and (l1 == 0 and j1 != 1 or l2 == 0 and j2 != 1)):
                continue
            try:
                v = float(a.getRadialMatrixElement(n1, l1, j1, n2, l2, j2, s=s))
            except Exception:
                continue
            out["dipole"].append([n1, l1, j1, n2, l2, j2, s, v]); cnt += 1
        cnt = 0
        while cnt < 20:
            n1, n2 = int(rng.integers(40, 75)), int(rng.integers(40, 75))
            l1 = int(rng.integers(0, 25))
            l2 = l1 + int(rng.choice([-2, 0, 2]))
            if l2 < 0 or l1 >= n1 or l2 >= n2:
                continue
            j1 = l1 + int(rng.integers(-s, s + 1)); j2 = l2 + int(rng.integers(-s, s + 1))
            if j1 < 0 or j2 < 0 or (s == 1 and (l1 == 0 and j1 != 1 or l2 == 0 and j2 != 1)):
                continue
            try:
                v = float(a.getQuadrupoleMatrixElement(n1, l1, j1, n2, l2, j2, s=s))
            except Exception:
                continue
            out["quadrupole"].append([n1, l1, j1, n2, l2, j2, s, v]); cnt += 1
        for n in (45, 60, 70):
            for l in range(0, 8):
                for j in range(max(l - s, 0), l + s + 1):
                    if s == 1 and l == 0 and j != 1:
                        continue
                    try:
                        out["energies_eV"].append([n, l, j, s, float(a.getEnergy(n, l, j, s=s))])
                    except Exception:
                        pass
    with open(os.path.join(GOLD, "divalent_golden.json"), "w") as f:
        json.dump(out, f)
    print("divalent golden", {k: len(v) for k, v in out.items()})
    # small Sr88 singlet Stark map
    _stark_case(arc, a, (55, 0, 0, 0, 53, 57, 12), np.linspace(0, 100, 4),
                "stark_sr88_singlet_golden.npz", kwargs={"s": 0})


def gen_lown(arc):
    """Stark basis that reaches below groundStateN: the reference keeps the states listed in
    atom.extraLevels (Rb 4D, 4F; K 3D) and drops the other n < groundStateN states."""
    _stark_case(arc, arc.Rubidium(), (6, 0, 0.5, 0.5, 4, 7, 3), np.linspace(0, 1e5, 3),
                "stark_rb_lown_golden.npz")
    _stark_case(arc, arc.Potassium(), (5, 0, 0.5, 0.5, 3, 6, 2), np.linspace(0, 1e5, 3),
                "stark_k_lown_golden.npz")


def numerovback_cases():
    """(name, inner, outer, kfun, step, init1, init2) -- analytic callables, shared by the golden
    generator and the tests (tests/test_radial_gpu.py imports this)."""
    from math import exp

    def hyd(l, E):
        def k(x):
            r = x * x
            return -3.0 / (4.0 * r) + 4.0 * r * (2.0 * (E + 1.0 / r) - l * (l + 1) / (r ** 2))
        return k

    def screened(x):          # scalar-only callable (math.exp): exercises the per-point fallback
        r = x * x
        return -3.0 / (4.0 * r) + 4.0 * r * (2.0 * (-0.01 + (1.0 + 5.0 * exp(-1.3 * r)) / r) - 2.0 / (r ** 2))
    return [("hyd_10s", 0.05, 2.0 * 10 * 25, hyd(0, -0.5 / 100.0), 0.001, 0.01, 0.01),
            ("hyd_6p", 0.01, 2.0 * 6 * 21, hyd(1, -0.5 / 36.0), 0.002, 0.01, 0.01),
            ("hyd_offres_l2_diverging", 0.01, 2.0 * 8 * 23, hyd(2, -0.5 / 6.3 ** 2), 0.001, 0.01, 0.01),
            ("hyd_offres_l1_diverging", 0.0025, 2.0 * 8 * 23, hyd(1, -0.5 / 5.5 ** 2), 0.001, 0.01, 0.01),
            ("screened_7p", 0.3, 2.0 * 7 * 22, screened, 0.001, 0.01, 0.005)]


def gen_numerovback(arc):
    """Outputs of the live reference's pure-Python NumerovBack (aaf:3939-4063) for analytic
    callables, and of radialWavefunction on the cpp_numerov=False path (aaf:606-623)."""
    from arc.alkali_atom_functions import NumerovBack
    d = {}
    for name, inner, outer, k, step, i1, i2 in numerovback_cases():
        r, psi = NumerovBack(inner, outer, k, step, i1, i2)
        d[name + "_r"], d[name + "_psi"] = r[::7].copy(), psi[::7].copy()
        d[name + "_head_r"], d[name + "_head_psi"] = r[:400].copy(), psi[:400].copy()
        d[name + "_L"] = len(r)
        d[name + "_sum"] = np.array([psi.sum(), np.abs(psi).sum(), r.sum(), float((psi == 0).sum())])
        print(name, len(r), int((psi == 0).sum()))
    a = arc.Rubidium(cpp_numerov=False)
    for (n, l, j) in ((9, 0, 0.5), (8, 2, 2.5)):
        E = a.getEnergy(n, l, j) / 27.211
        r, psi = a.radialWavefunction(l, 0.5, j, E, a.alphaC ** (1 / 3.0), 2.0 * n * (n + 15.0), 0.001)
        nm = "rb_%d_%d" % (n, l)
        d[nm + "_r"], d[nm + "_psi"], d[nm + "_L"] = r[::7].copy(), psi[::7].copy(), len(r)
        d[nm + "_args"] = np.array([l, 0.5, j, E, a.alphaC ** (1 / 3.0), 2.0 * n * (n + 15.0), 0.001])
        print(nm, len(r))
    np.savez_compressed(os.path.join(GOLD, "numerovback_golden.npz"), **d)


def anglepairs_cases():
    """(name, ctor args, anglePairs, nRange, energyDelta) shared by generator and tests."""
    return [("dd_same", (40, 2, 2.5, 40, 2, 2.5, 2.5, 2.5), [(0.0, 0.0), (np.pi / 4, 0.0), (np.pi / 2, 0.0), (1.0, 0.0)],
             3, 30e9),
            ("ss_diff_n", (40, 0, 0.5, 41, 0, 0.5, 0.5, -0.5), [(0.3, 0.0), (1.2, 0.0)], 3, 30e9),
            ("pp_phi", (38, 1, 1.5, 38, 1, 1.5, 1.5, 0.5), [(0.7, 1.1), (0.2, -0.4)], 2, 25e9),
            ("sd_diff_phi", (42, 0, 0.5, 41, 2, 1.5, 0.5, 1.5), [(0.9, 0.6)], 2, 25e9)]


def gen_anglepairs(arc):
    """getC6perturbatively_anglePairs (pair:2065-2287) of the live reference."""
    a = arc.Rubidium()
    d = {}
    for name, ctor, angles, nRange, dE in anglepairs_cases():
        calc = arc.PairStateInteractions(a, *ctor)
        c6, c6hop, mats = calc.getC6perturbatively_anglePairs(angles, nRange, dE, returnInteractionMatrix=True)
        d[name + "_c6"] = np.array(c6, dtype=float)
        d[name + "_c6hop"] = np.array(c6hop, dtype=float)
        d[name + "_mats"] = np.array(mats)
        ev, vecs, deg = calc.getC6perturbatively_anglePairs(angles, nRange, dE, degeneratePerturbation=True)
        d[name + "_ev"] = np.array(ev, dtype=float)
        d[name + "_vecs"] = np.array(vecs)
        d[name + "_deg"] = np.array(deg, dtype=float)
        print(name, np.array(c6), np.array(c6hop), np.array(ev).shape)
    np.savez_compressed(os.path.join(GOLD, "anglepairs_golden.npz"), **d)


def gen_c4(arc, which):
    """C4 (BASELINE.json configs[3]) at its stated size: Sr88 n=50-70, lmax=30, singlet N=651 /
    triplet N=1932; 8 of the 2000 fields kept."""
    a = arc.Strontium88()
    F = np.linspace(0, 200, 2000)[[0, 1, 250, 700, 1000, 1400, 1750, 1999]]
    if which == "singlet":
        _stark_case(arc, a, (60, 0, 0, 0, 50, 70, 30), F, "stark_c4_singlet_golden.npz", kwargs={"s": 0})
    else:
        _stark_case(arc, a, (60, 0, 1, 0, 50, 70, 30), F, "stark_c4_triplet_golden.npz", kwargs={"s": 1})


def gen_c5(arc):
    """C5 (BASELINE.json configs[4]) at its stated size: Rb 80D5/2+80D5/2, theta=pi/4, Bz=1e-4 T,
    nRange=3, lrange=4, 10 GHz -> N=10384; 2 of the 4096 R points."""
    a = arc.Rubidium()
    R = np.linspace(2, 20, 4096)[[600, 3500]]
    _pair_case(arc, a, (80, 2, 2.5, 80, 2, 2.5, 2.5, 2.5), (np.pi / 4, 0, 3, 4, 10e9, 1e-4),
               list(R), 250, "pair_c5_golden.npz")


def gen_driving(arc):
    """drivingFromState variants: highlight = laser-coupling weighted admixture
    (calculations_atom_single.py:887-962,1003-1021; calculations_atom_pairstate.py:3332-3413,3507-3520)."""
    a = arc.Rubidium()
    c = arc.StarkMap(a)
    c.defineBasis(30, 2, 2.5, 0.5, 28, 32, 6)
    F = np.linspace(0, 800, 4)
    drive = [5, 1, 1.5, 0.5, 0]
    c.diagonalise(F, drivingFromState=drive)
    np.savez_compressed(os.path.join(GOLD, "stark_driving_golden.npz"),
                        args=np.array([30, 2, 2.5, 0.5, 28, 32, 6], dtype=float), drive=np.array(drive, dtype=float),
                        fields=F, y=np.array(c.y), highlight=np.array(c.highlight), maxCoupling=c.maxCoupling)
    print("stark driving golden: N", len(c.basisStates), "maxCoupling", c.maxCoupling)
    p = arc.PairStateInteractions(a, 40, 0, .5, 40, 0, .5, .5, .5)
    p.defineBasis(0, 0, 2, 3, 25e9)
    drive = [39, 1, 1.5, 0.5, 0]
    R = np.array([1.5, 3.0, 6.0])
    k = 30
    p.diagonalise(R, k, drivingFromState=drive)
    np.savez_compressed(os.path.join(GOLD, "pair_driving_golden.npz"),
                        ctor=np.array([40, 0, .5, 40, 0, .5, .5, .5]), basis=np.array([0, 0, 2, 3, 25e9]),
                        drive=np.array(drive, dtype=float), r=np.array(p.r), k=k, y=np.array(p.y),
                        highlight=np.array(p.highlight), maxCoupling=p.maxCoupling,
                        maxCoupledStateIndex=p.maxCoupledStateIndex, N=len(p.basisStates))
    print("pair driving golden: N", len(p.basisStates), "maxCoupling", p.maxCoupling)


def gen_sorted(arc):
    """sortEigenvectors=True (adiabatic re-ordering, calculations_atom_pairstate.py:3452-3492)."""
    a = arc.Rubidium()
    p = arc.PairStateInteractions(a, 60, 0, .5, 60, 0, .5, .5, .5)
    p.defineBasis(0, 0, 2, 3, 8e9)
    R = np.linspace(1.2, 8.0, 24)
    k = 20
    p.diagonalise(R, k, sortEigenvectors=True)
    np.savez_compressed(os.path.join(GOLD, "pair_sorted_golden.npz"),
                        ctor=np.array([60, 0, .5, 60, 0, .5, .5, .5]), basis=np.array([0, 0, 2, 3, 8e9]),
                        r=np.array(p.r), k=k, y=np.array(p.y), highlight=np.array(p.highlight),
                        N=len(p.basisStates))
    print("pair sorted golden: N", len(p.basisStates), np.array(p.y).shape)


def gen_c6(arc):
    """getC6perturbatively known answers (calculations_atom_pairstate.py:1144-1440)."""
    out = []
    k = arc.Potassium()
    out.append({"atom": "Potassium39", "ctor": [66, 0, .5, 66, 0, .5, .5, .5], "args": [0, 0, 5, 30e9],
                "value": float(arc.PairStateInteractions(k, 66, 0, .5, 66, 0, .5, .5, .5).getC6perturbatively(0, 0, 5, 30e9)),
                "pin": "test/test_pairstate.py:4-10 (-265 +- 1)"})
    rb = arc.Rubidium()
    out.append({"atom": "Rubidium", "ctor": [60, 0, .5, 60, 0, .5, .5, .5], "args": [0, 0, 5, 25e9],
                "value": float(arc.PairStateInteractions(rb, 60, 0, .5, 60, 0, .5, .5, .5).getC6perturbatively(0, 0, 5, 25e9)),
                "pin": "doc/Rydberg_atoms_a_primer_notebook.ipynb:1710 (-138.88... sign per ARC)"})
    for th in (0.0, 0.7, 1.5):
        out.append({"atom": "Rubidium", "ctor": [60, 2, 2.5, 60, 2, 2.5, 2.5, 2.5], "args": [th, 0, 5, 25e9],
                    "value": float(arc.PairStateInteractions(rb, 60, 2, 2.5, 60, 2, 2.5, 2.5, 2.5).getC6perturbatively(th, 0, 5, 25e9)),
                    "pin": "live reference"})
    with open(os.path.join(GOLD, "c6_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("c6 golden", [(o["atom"], o["args"][0], o["value"]) for o in out])


def gen_shirley(arc):
    """ShirleyMethod (Floquet Stark maps, calculations_atom_single.py:3172-3982):
    small Rb bases, full manifold (edN=0) and dipole-limited (edN=2), fn = 1 and 2."""
    rb = arc.Rubidium()
    cases = [
        dict(name="a", basis=(30, 0, 0.5, 0.5), kw=dict(q=0, nMin=28, nMax=32, maxL=3, edN=0), fn=1,
             eFields=[0.0, 50.0, 200.0], freqs=[1.0e9, 6.0e9]),
        dict(name="b", basis=(30, 2, 2.5, 0.5), kw=dict(q=1, nMin=27, nMax=33, maxL=5, edN=2, Bz=1e-4), fn=2,
             eFields=[100.0], freqs=[2.0e9, 3.5e9, 10.0e9]),
    ]
    out = {}
    for c in cases:
        t0 = time.time()
        calc = arc.ShirleyMethod(rb)
        calc.defineBasis(*c["basis"], **c["kw"])
        calc.defineShirleyHamiltonian(fn=c["fn"])
        calc.diagonalise(c["eFields"], c["freqs"])
        k = c["name"]
        out[k + "_basis"] = np.array(c["basis"], dtype=float)
        out[k + "_kw"] = np.array([c["kw"].get("q", 0), c["kw"]["nMin"], c["kw"]["nMax"], c["kw"]["maxL"],
                                   c["kw"]["edN"], c["kw"].get("Bz", 0.0), c["fn"]], dtype=float)
        out[k + "_eFields"] = np.array(c["eFields"], dtype=float)
        out[k + "_freqs"] = np.array(c["freqs"], dtype=float)
        out[k + "_basisStates"] = np.array(calc.basisStates, dtype=float)
        out[k + "_index"] = calc.indexOfCoupledState
        out[k + "_bareEnergies"] = np.array(calc.bareEnergies)
        out[k + "_V"] = np.array(calc.V)
        out[k + "_eigs"] = np.array(calc.eigs)
        out[k + "_transProbs"] = np.array(calc.transProbs)
        out[k + "_targetShifts"] = np.array(calc.targetShifts)
        print("shirley", k, "dim0", len(calc.basisStates), "eigs", np.array(calc.eigs).shape,
              "%.1fs" % (time.time() - t0))
    np.savez_compressed(os.path.join(GOLD, "shirley_golden.npz"), **out)


def gen_resonances(arc):
    """StarkMapResonances.findResonances (calculations_atom_pairstate.py:4590-4790), small Rb
    case: pair 30S1/2 30S1/2 -> dipole-coupled P-state pairs, 4 fields."""
    from unittest.mock import MagicMock
    import arc.calculations_atom_pairstate as cap
    cap.plt.subplots.return_value = (MagicMock(), MagicMock())                            # fig, ax
    rb = arc.Rubidium()
    st1, st2 = [30, 0, 0.5, 0.5], [30, 0, 0.5, 0.5]
    calc = arc.StarkMapResonances(rb, list(st1), rb, list(st2))
    F = np.linspace(0, 600, 4)
    t0 = time.time()
    calc.findResonances(28, 32, 5, F, energyRange=[-20e9, 20e9])
    out = dict(state1=np.array(st1), state2=np.array(st2), args=np.array([28, 32, 5], dtype=float), fields=F,
               energyRange=np.array([-20e9, 20e9]), pairStateEnergy=calc.pairStateEnergy,
               originalStateY=np.array(calc.originalStateY),
               originalStateContribution=np.array(calc.originalStateContribution))
    for i in range(len(F)):
        out["y_%d" % i] = np.array(calc.y[i], dtype=float)
        out["comp_%d" % i] = np.array(calc.composition[i]).astype(str) if len(calc.composition[i]) else np.zeros((0, 2), dtype=str)
    np.savez_compressed(os.path.join(GOLD, "resonances_golden.npz"), **out)
    print("resonances", [len(calc.y[i]) for i in range(len(F))], "%.1fs" % (time.time() - t0))


def gen_levelfit(arc):
    """getC6fromLevelDiagram / getC3fromLevelDiagram / getVdwFromLevelDiagram
    (calculations_atom_pairstate.py:3978-4472) on small Rb pair bases."""
    import json as _json
    import arc.calculations_atom_pairstate as cap
    from unittest.mock import MagicMock
    cap.plt.subplots.return_value = (MagicMock(), MagicMock())
    rb = arc.Rubidium()
    out = {}
    # off-resonant S+S pair: van der Waals (C6) regime
    p = arc.PairStateInteractions(rb, 60, 0, 0.5, 60, 0, 0.5, 0.5, 0.5)
    p.defineBasis(0., 0., 2, 2, 15e9)
    R = np.linspace(3.0, 10.0, 36)
    p.diagonalise(R, 40)
    out["c6"] = {"ctor": [60, 0, .5, 60, 0, .5, .5, .5], "basis": [0., 0., 2, 2, 15e9], "R": [3.0, 10.0, 36], "k": 40,
                 "N": len(p.basisStates), "range": [4.0, 9.5],
                 "C6": float(p.getC6fromLevelDiagram(4.0, 9.5)),
                 "Rvdw": float(p.getVdwFromLevelDiagram(3.0, 9.5, minStateContribution=0.5))}
    # resonant S+P pair: C3 regime
    p = arc.PairStateInteractions(rb, 60, 0, 0.5, 60, 1, 1.5, 0.5, 0.5)
    p.defineBasis(0., 0., 1, 1, 10e9)
    R = np.linspace(2.0, 12.0, 30)
    p.diagonalise(R, 20)
    out["c3"] = {"ctor": [60, 0, .5, 60, 1, 1.5, .5, .5], "basis": [0., 0., 1, 1, 10e9], "R": [2.0, 12.0, 30], "k": 20,
                 "N": len(p.basisStates), "range": [3.0, 11.0],
                 "C3": float(p.getC3fromLevelDiagram(3.0, 11.0, resonantBranch=+1))}
    with open(os.path.join(GOLD, "levelfit_golden.json"), "w") as f:
        _json.dump(out, f, indent=1)
    print("levelfit", out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["radial", "angular", "stark", "pair", "divalent"]
    arc = refenv.load(home_name="home_golden_" + "_".join(what))
    if "radial" in what:
        gen_radial(arc)
    if "angular" in what:
        gen_angular(arc)
    if "stark" in what:
        gen_stark(arc)
    if "starkpol" in what:
        gen_stark_full_pol(arc)
    if "pairsmall" in what:
        gen_pair(arc, big=False)
    if "pair" in what:
        gen_pair(arc, big=True)
    if "divalent" in what:
        gen_divalent(arc)
    if "anglepairs" in what:
        gen_anglepairs(arc)
    if "numerovback" in what:
        gen_numerovback(arc)
    if "lown" in what:
        gen_lown(arc)
    if "c4singlet" in what:
        gen_c4(arc, "singlet")
    if "c4triplet" in what:
        gen_c4(arc, "triplet")
    if "c5" in what:
        gen_c5(arc)
    if "driving" in what:
        gen_driving(arc)
    if "sorted" in what:
        gen_sorted(arc)
    if "c6" in what:
        gen_c6(arc)
    if "shirley" in what:
        gen_shirley(arc)
    if "resonances" in what:
        gen_resonances(arc)
    if "levelfit" in what:
        gen_levelfit(arc)
