"""oracle/oracle.py -- CPU restatement of the reference hot path (the CHECKER).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
`cpu_baseline` / `--impl reference` legs may import this module.  The product
package never does (it fails loudly when its CUDA library is missing).

Parity status: PINNED.
  * Numerov: liboracle.so (numerov_oracle.c) is checked bit-for-bit against the
    reference's own arc_c_extensions.c compiled into oracle/_ref
    (tests/test_oracle.py::test_restatement_matches_reference_build) and against
    golden vectors frozen from the live reference (tests/golden/radial_golden.json).
  * eigensolves: the reference calls numpy.linalg.eigh (LAPACK dsyevd) and
    scipy.sparse.linalg.eigsh (ARPACK mode 3) -- third-party code that is present
    in this image (numpy 2.3 / scipy 1.18), so the oracle calls the
    very same functions; golden y/highlight come from the live reference
    (tests/golden/stark_c1_golden.npz, pair_c3_golden.npz).
  * semiclassical (divalent) radial elements: restates
    alkali_atom_functions.py:2683-2802 on top of mpmath.angerj (mpmath 1.3.0, the
    reference's own dependency); golden values in tests/golden/radial_golden.json.

Every function cites the reference file:line it follows
(aaf = arc/alkali_atom_functions.py, single = arc/calculations_atom_single.py,
 pair = arc/calculations_atom_pairstate.py, cext = arc/arc_c_extensions.c).
"""
import ctypes
import glob
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(ref=True):
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    if ref and os.path.exists("/root/reference/arc/arc_c_extensions.c"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        lib = ctypes.CDLL(path)
        d, i = ctypes.c_double, ctypes.c_int
        lib.arc_oracle_numerov_length.argtypes = [d, d, d]
        lib.arc_oracle_numerov_length.restype = i
        lib.arc_oracle_numerov.argtypes = [d, d, d, d, d, i, d, d, d, d, d, i,
                                           d, d, d, d, d, d,
                                           ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.POINTER(i)]
        lib.arc_oracle_numerov.restype = i
        _LIB = lib
    return _LIB


def have_ref():
    return bool(glob.glob(os.path.join(_HERE, "_ref", "arc_c_extensions*.so")))


def ref_module():
    """The reference's own C extension built into oracle/_ref (kind='reference')."""
    global _REF
    if _REF is None:
        so = glob.glob(os.path.join(_HERE, "_ref", "arc_c_extensions*.so"))
        if not so:
            raise RuntimeError("oracle/_ref not built (needs /root/reference)")
        spec = importlib.util.spec_from_file_location("arc_c_extensions", so[0])
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _REF = m
    return _REF


# --------------------------------------------------------------------------- #
# subsystem (1): Numerov + radial integrals
# --------------------------------------------------------------------------- #
def numerov(innerLimit, outerLimit, step, init1, init2, l, s, j, stateEnergy,
            alphaC, alpha, Z, a1, a2, a3, a4, rc, mu):
    """cext:108-285 via the C restatement. Returns (psi=R*r unnormalised, r, dp)."""
    lib = _lib()
    L = lib.arc_oracle_numerov_length(innerLimit, outerLimit, step)
    psi = np.empty(L + 2, dtype=np.float64)
    r = np.empty(L + 2, dtype=np.float64)
    dp = ctypes.c_int(0)
    rc_ = lib.arc_oracle_numerov(innerLimit, outerLimit, step, init1, init2, int(l),
                                 s, j, stateEnergy, alphaC, alpha, int(Z), a1, a2,
                                 a3, a4, rc, mu, psi.ctypes.data, r.ctypes.data,
                                 ctypes.byref(dp))
    if rc_ < 0:
        raise RuntimeError("oracle numerov failed (%d)" % rc_)
    return psi[:L], r[:L], dp.value


def ref_numerov(*args):
    """Same 18 arguments through the reference's own compiled C (oracle/_ref)."""
    d = ref_module().NumerovWavefunction(*args)
    L = d.shape[0]
    flat = np.frombuffer((ctypes.c_double * (2 * L)).from_address(d.ctypes.data))
    return flat[:L].copy(), flat[L:2 * L].copy()


def numerov_back(innerLimit, outerLimit, kfun, step, init1, init2):
    """Restatement of the pure-Python integrator NumerovBack (aaf:3939-4063): same grid
    (aaf:3984-3991), recurrence (aaf:3992-4019), checkPoint / divergence walk (aaf:4021-4047),
    tail rule (aaf:4049-4054) and output conversion (aaf:4056-4060).  Pinned by
    tests/golden/numerovback_golden.npz (outputs of the live reference)."""
    from math import sqrt
    br = round((sqrt(outerLimit) - sqrt(innerLimit)) / step)
    sol = np.zeros(br)
    rad = np.zeros(br)
    h2 = step ** 2

    def nxt(x, y1, y2):
        return ((2.0 * (1.0 - 5.0 / 12.0 * h2 * kfun(x)) * y1 - (1.0 + 1.0 / 12.0 * h2 * kfun(x + step)) * y2)
                / (1 + 1 / 12.0 * h2 * kfun(x - step)))
    br -= 1
    x = sqrt(innerLimit) + step * (br - 1)
    sol[br] = nxt(x, init1, init2)
    rad[br] = x
    x -= step
    br -= 1
    sol[br] = nxt(x, sol[br + 1], init1)
    rad[br] = x
    maxValue, checkPoint, fromLastMax = 0.0, 0, 0
    while br > checkPoint:
        br -= 1
        x -= step
        sol[br] = nxt(x, sol[br + 1], sol[br + 2])
        rad[br] = x
        if abs(sol[br] * sqrt(x)) > maxValue:
            maxValue = abs(sol[br] * sqrt(x))
        else:
            fromLastMax += 1
            if fromLastMax > 50:
                checkPoint = br
    divergencePoint = 0
    while br > 0 and divergencePoint == 0:
        br -= 1
        x -= step
        sol[br] = nxt(x, sol[br + 1], sol[br + 2])
        rad[br] = x
        if divergencePoint == 0 and abs(sol[br] * sqrt(x)) > maxValue:
            divergencePoint = br
            while abs(sol[divergencePoint]) > abs(sol[divergencePoint + 1]) and divergencePoint < checkPoint:
                divergencePoint += 1
            if divergencePoint > checkPoint:
                raise RuntimeError("Numerov error")
    br = divergencePoint
    while br > 0:
        rad[br] = rad[br + 1] - step
        sol[br] = 0
        br -= 1
    return rad * rad, sol * np.sqrt(rad), divergencePoint


def trapezoid(y, x):
    """scipy.integrate.trapezoid(y, x=x) (aaf:34): sum(dx * (y[1:]+y[:-1]) / 2)."""
    d = np.diff(x)
    return float(np.sum(d * (y[1:] + y[:-1]) / 2.0))


def radial_wavefunction(sp, step=0.001, use_ref=False):
    """aaf:503-625 with the call-site arguments of aaf:1068-1085.

    sp: dict(n,l,j,E_au,alphaC,alpha,Z,a1,a2,a3,a4,rc,mu); s=0.5 as at aaf:1070.
    innerLimit = max(4*step, alphaC**(1/3)); outerLimit = 2 n (n+15)."""
    inner = max(4.0 * step, sp["alphaC"] ** (1 / 3.0))
    outer = 2.0 * sp["n"] * (sp["n"] + 15.0)
    args = (inner, outer, step, 0.01, 0.01, int(sp["l"]), 0.5, float(sp["j"]),
            sp["E_au"], sp["alphaC"], sp["alpha"], int(sp["Z"]), sp["a1"], sp["a2"],
            sp["a3"], sp["a4"], sp["rc"], sp["mu"])
    if use_ref:
        psi, r = ref_numerov(*args)
    else:
        psi, r, _ = numerov(*args)
    suma = trapezoid(psi ** 2, r)                                   # aaf:605
    return r, psi / np.sqrt(suma)                                   # aaf:606


def radial_integral(sp1, sp2, power, use_ref=False):
    """aaf:1087-1096 (power=1, dipole) / aaf:1191-1201 (power=2, quadrupole).
    sp1 must be the LOWER-energy state (ordering done by the caller, aaf:1021)."""
    r1, p1 = radial_wavefunction(sp1, use_ref=use_ref)
    r2, p2 = radial_wavefunction(sp2, use_ref=use_ref)
    up = min(len(r1), len(r2))
    w = r1[:up] if power == 1 else r1[:up] * r1[:up]
    return trapezoid(p1[:up] * p2[:up] * w, r1[:up])


def semiclassical_dipole(nu, nu1, l1, l2):
    """aaf:2683-2735 given nu=sqrt(-Ry/E1), nu1=sqrt(-Ry/E2) (state 1 lower)."""
    from mpmath import angerj
    l_c = (l1 + l2 + 1.0) / 2.0
    nu_c = np.sqrt(nu * nu1)
    dn = nu - nu1
    gamma = ((l2 - l1) * l_c) / nu_c
    if dn == 0:
        g0, g1, g2, g3 = 1, 0, 0, 0
    else:
        am, ap = angerj(dn - 1.0, -dn), angerj(dn + 1, -dn)
        g0 = (1.0 / (3.0 * dn)) * (am - ap)
        g1 = -(1.0 / (3.0 * dn)) * (am + ap)
        g2 = g0 - np.sin(np.pi * dn) / (np.pi * dn)
        g3 = (dn / 2.0) * g0 + g1
    return float((3 / 2) * nu_c ** 2 * (1 - (l_c / nu_c) ** 2) ** 0.5
                 * (g0 + gamma * g1 + gamma ** 2 * g2 + gamma ** 3 * g3))


def semiclassical_quadrupole(nu, nu1, l1, l2):
    """aaf:2737-2802 given nu=n1-delta1, nu1=n2-delta2."""
    from mpmath import angerj
    dl = abs(l2 - l1)
    l_c = (l1 + l2 + 1.0) / 2.0
    nu_c = np.sqrt(nu * nu1)
    dn = nu - nu1
    gamma = ((l2 - l1) * l_c) / nu_c
    if dn == 0:
        q = np.array([1.0, 0, 0, 0])
    else:
        am, ap = angerj(dn - 1.0, -dn), angerj(dn + 1, -dn)
        g0 = float((1.0 / (3.0 * dn)) * (am - ap))
        g1 = float(-(1.0 / (3.0 * dn)) * (am + ap))
        q = np.zeros(4)
        q[0] = -(6.0 / (5.0 * dn)) * g1
        q[1] = -(6.0 / (5.0 * dn)) * g0 + (6.0 / 5.0) * np.sin(np.pi * dn) / (np.pi * dn ** 2)
        q[2] = -(3.0 / 4.0) * (6.0 / (5.0 * dn) * g1 + g0)
        q[3] = 0.5 * (dn * 0.5 * q[0] + q[1])
    if dl == 0:
        pre = (5.0 / 2.0) * nu_c ** 4 * (1.0 - (3.0 * l_c ** 2) / (5 * nu_c ** 2))
        return float(pre * (q[0] + gamma ** 2 * q[2]))
    if dl == 2:
        pre = ((5.0 / 2.0) * nu_c ** 4 * (1 - (l_c + 1) ** 2 / nu_c ** 2) ** 0.5
               * (1 - (l_c + 2) ** 2 / nu_c ** 2) ** 0.5)
        return float(pre * sum(gamma ** p * q[p] for p in range(4)))
    return 0.0


# --------------------------------------------------------------------------- #
# subsystems (2)+(3): assembly + eigensolve per sweep point
# --------------------------------------------------------------------------- #
def state_composition2(vec, upTo=4, totalContributionMax=0.95):
    """single:1395-1416."""
    contribution = np.absolute(vec)
    order = np.argsort(contribution, kind="heapsort")
    out, index, total = [], -1, 0
    if upTo == -1:
        for k in range(len(order)):
            i = order[-k - 1]
            out.append([vec[i], i])
        return out
    while index > -upTo and total < totalContributionMax:
        i = order[index]
        out.append([vec[i], i])
        total += contribution[i] ** 2
        index -= 1
    return out


def stark_diagonalise(mat1, mat2, fields, idx0, drive_w=None, upTo=4,
                      totalContributionMax=0.95, want_composition=True):
    """single:976-1021: per field m = mat1 + mat2*F; numpy.linalg.eigh."""
    ys, hls, comps = [], [], []
    for F in fields:
        ev, vec = np.linalg.eigh(mat1 + mat2 * F)                  # single:984-986
        ys.append(ev)
        if drive_w is None:
            hls.append(np.abs(vec[idx0, :]) ** 2)                  # single:993
        else:
            hls.append((np.abs(drive_w)[:, None] * np.abs(vec ** 2)).sum(axis=0))  # single:1006-1011
        if want_composition:
            comps.append([state_composition2(vec[:, i], upTo, totalContributionMax)
                          for i in range(len(ev))])
    return np.array(ys), np.array(hls), comps


def pair_matrix(mat_diag, matR, rval):
    """pair:3435-3440: m = matDiagonal + sum_o matR[o]/(R*1e-6)^(3+o)."""
    m = mat_diag
    rX = (rval * 1.0e-6) ** 3
    for mr in matR:
        m = m + mr / rX
        rX *= rval * 1.0e-6
    return m


def pair_diagonalise(mat_diag, matR, rangeR, k, opi, sigma_hz=0.0, method="eigsh",
                     drive_w=None):
    """pair:3307, 3425-3520.  method='eigsh' is the reference call (ARPACK
    shift-invert, tol=1e-6); method='dense' = numpy.linalg.eigh + the k
    eigenvalues nearest sigma, ascending (agrees with eigsh to 2e-14 GHz,
    SURVEY.md 8c rule iv)."""
    from scipy.sparse.linalg import eigsh
    r = np.sort(np.asarray(rangeR, dtype=float))[::-1]              # pair:3307
    ys, hls, vecs = [], [], []
    for rval in r:
        m = pair_matrix(mat_diag, matR, rval)
        if method == "eigsh":
            ev, vec = eigsh(m, k, sigma=sigma_hz * 1e-9, which="LM", tol=1e-6)
        else:
            md = m.toarray() if hasattr(m, "toarray") else np.asarray(m)
            w, v = np.linalg.eigh(md)
            sel = np.sort(np.argsort(np.abs(w - sigma_hz * 1e-9), kind="stable")[:k])
            ev, vec = w[sel], v[:, sel]
        ys.append(ev)
        if drive_w is None:
            hls.append(np.abs(vec[opi, :]) ** 2)                   # pair:3502
        else:
            hls.append((np.abs(drive_w)[:, None] * np.abs(vec) ** 2).sum(axis=0))
        vecs.append(vec)
    return r, np.array(ys), np.array(hls), vecs
