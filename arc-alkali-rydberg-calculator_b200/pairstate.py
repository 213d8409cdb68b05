"""PairStateInteractions -- drop-in for the defineBasis/diagonalise path of
arc.calculations_atom_pairstate.PairStateInteractions (calculations_atom_pairstate.py:97-3522).

Same constructor and method signatures and the same result attributes (`channel`,
`coupling`, `basisStates`, `index`, `matDiagonal`, `matR`, `originalPairStateIndex`, `r`,
`y`, `highlight`, `composition`).  What is different underneath:
  * every radial dipole/quadrupole element of both atoms comes from ONE batched CUDA
    Numerov run (atoms.precompute_radial) instead of a per-element Python->C->sqlite chain;
  * the channel-coupling predicate (`__isCoupled`, :625-710) is evaluated with numpy
    broadcasting over all channel pairs;
  * the m-resolved angular blocks use the closed form
        matR[o][rows of j, cols of i] = rho * Re[(D(j1')xD(j2')) d (D(j1)xD(j2))^H]
    (SURVEY.md 8a/a16'), with d = elem * sum_p f_p kron(A_p, B_p) built from two small
    single-atom Clebsch-Gordan matrices instead of a (2j+1)^4 Python loop;
  * diagonalise runs the dense batched CUDA eigensolver (sweep.pair_sweep) and returns
    the k eigenpairs nearest the requested detuning, ascending -- the semantics of the
    reference's shift-invert ARPACK call (:3444-3450).
No CPU fallback.
"""
from math import factorial, floor, pi, sqrt

import numpy as np
from scipy.constants import epsilon_0, physical_constants, h as C_h, e as C_e
from scipy.sparse import csr_matrix

from .wigner import CG, Wigner6j, WignerDmatrix
from . import sweep as _sweep


def printStateLetter(l):
    return "SPDFGHIKLMN"[l] if l <= 10 else " l=%d" % l


def printStateStringLatex(n, l, j, s=None):
    """alkali_atom_functions.py:4343-4369."""
    if s is None:
        return str(n) + printStateLetter(l) + ("_{%.0d/2}" % (j * 2))
    if abs(floor(j) - j) < 0.1:
        subscript = "_{%.0d}" % (j)
    else:
        subscript = "_{%.0d/2}" % (j * 2)
    return str(n) + (" ^{%d}" % (round(2 * s + 1))) + printStateLetter(l) + subscript


def _coupling_order(l, j, l1, j1, upTo):
    """1 = dipole, 2 = quadrupole, 0 = none (calculations_atom_pairstate.py:682-707)."""
    dl, dj = abs(l - l1), abs(j - j1)
    if dl == 1 and dj < 1.1:
        return 1
    if dl in (0, 1, 2) and dj < 2.1 and 2 <= upTo:
        return 2
    return 0

def _unembed_hermitian(w2, v2, n):
    """Eigenpairs of an n x n complex-Hermitian matrix from those of its 2n x 2n real symmetric
    embedding [[Re, -Im], [Im, Re]] (w2 ascending, v2[:, i] eigenvectors): every eigenvalue of the
    Hermitian matrix appears twice; inside each cluster of equal eigenvalues the embedded vectors
    (x; y) -> x + i y span the complex eigenspace, an orthonormal basis of which is taken from
    their SVD.  Returns (w[n] ascending, vecs[n, n] with vecs[i] the i-th eigenvector)."""
    z = v2[:n, :] + 1j * v2[n:, :]
    scale = max(np.abs(w2).max(), 1e-300)
    w, vecs = [], []
    a = 0
    while a < 2 * n:
        b = a + 1
        while b < 2 * n and abs(w2[b] - w2[a]) <= 1e-10 * scale:
            b += 1
        d = (b - a) // 2
        u, _s, _vh = np.linalg.svd(z[:, a:b], full_matrices=False)
        for q in range(d):
            w.append(float(np.mean(w2[a:b])))
            vecs.append(u[:, q])
        a = b
    if len(w) != n:          # a cluster was split by the tolerance: plain pairing of the doubled spectrum
        w = [float(0.5 * (w2[2 * i] + w2[2 * i + 1])) for i in range(n)]
        q, _r = np.linalg.qr(z[:, ::2])
        vecs = list(q.T)
    return np.array(w), np.array(vecs)


class PairCompositionView:
    """Lazy `composition[r][i]` -> LaTeX string of the <= 4 dominant basis states, as
    built by the reference's `_stateComposition` (calculations_atom_pairstate.py:3695-3727)."""

    def __init__(self, calc, comp_c, comp_i):
        self._calc, self.comp_c, self.comp_i = calc, comp_c, comp_i

    def __len__(self):
        return self.comp_c.shape[0]

    def entries(self, r, i):
        return [(self.comp_c[r, i, t], int(self.comp_i[r, i, t]))
                for t in range(self.comp_c.shape[2]) if self.comp_i[r, i, t] >= 0]

    def _string(self, r, i):
        value, total = "$", 0.0
        for t, (c, idx) in enumerate(self.entries(r, i)):
            if t != 0 and c > 0:
                value += "+"
            value += ("%.2f" % c) + self._calc._addState(*self._calc.basisStates[idx])
            total += c * c
        if total < 0.999:
            value += "+\\ldots"
        return value + "$"

    def __getitem__(self, r):
        if isinstance(r, slice):
            return [self[k] for k in range(*r.indices(len(self)))]
        return [self._string(r, i) for i in range(self.comp_c.shape[1])]

    def __iter__(self):
        for r in range(len(self)):
            yield self[r]


def _kron(A, B):
    """numpy.kron for two matrices (same products, without its generic-rank bookkeeping)."""
    return (A[:, None, :, None] * B[None, :, None, :]).reshape(A.shape[0] * B.shape[0], A.shape[1] * B.shape[1])


class PairStateInteractions:
    def __init__(self, atom, n, l, j, nn, ll, jj, m1, m2, interactionsUpTo=1, s=0.5,
                 s2=None, atom2=None, reference_6j_table_quirk=True):
        self.atom1 = atom
        self.atom2 = atom if atom2 is None else atom2
        self.n, self.l, self.j = n, l, j
        self.nn, self.ll, self.jj = nn, ll, jj
        self.m1, self.m2 = m1, m2
        self.interactionsUpTo = interactionsUpTo
        if getattr(atom, "divalent", False) and not (s == 0 or s == 1):
            raise ValueError("total angular spin s has to be defined explicitly for "
                             "calculations, and value has to be 0 or 1 for singlet and "
                             "tripplet states respectively.")
        self.s1 = s
        self.s2 = s if s2 is None else s2
        for a, sp, nm in ((self.atom1, self.s1, "atom1"), (self.atom2, self.s2, "atom2")):
            if getattr(a, "divalent", False):
                if abs(sp) > 0.1 and abs(sp - 1) > 0.1:
                    raise ValueError("%s is DivalentAtom and its spin has to be s=0 or s=1" % nm)
            elif abs(sp - 0.5) > 0.1:
                raise ValueError("%s is AlkaliAtom and its spin has to be s=0.5" % nm)
        if abs((self.s1 - self.m1) % 1) > 0.1:
            raise ValueError("atom1 with spin s = %.1d cannot have m1 = %.1d" % (self.s1, self.m1))
        if abs((self.s2 - self.m2) % 1) > 0.1:
            raise ValueError("atom2 with spin s = %.1d cannot have m2 = %.1d" % (self.s2, self.m2))
        # the reference's shipped Wigner-6j table returns 0 whenever the coupled state has
        # l1 = l + 2 (arc/data/precalculated6j.npy, lookup at wigner.py:307-325); keep that
        # behaviour by default so matR is identical to the reference's.
        self.reference_6j_table_quirk = reference_6j_table_quirk
        self.coupling, self.channel = [], []
        self.basisStates, self.matrixElement = [], []
        self.matDiagonal, self.matR = [], []
        self.originalPairStateIndex = 0
        self.r, self.y, self.highlight, self.composition = [], [], [], []
        self.fig = self.ax = 0
        self.maxCoupling = 0.0
        self.drivingFromState = [0, 0, 0, 0, 0]
        x = self.interactionsUpTo
        # factorial term of the angular matrix (calculations_atom_pairstate.py:389-406)
        self.fcp = np.zeros((x + 1, x + 1, 2 * x + 1))
        for c1 in range(1, x + 1):
            for c2 in range(1, x + 1):
                for p in range(-min(c1, c2), min(c1, c2) + 1):
                    self.fcp[c1, c2, p + x] = (factorial(c1 + c2)
                                               / (factorial(c1 + p) * factorial(c1 - p)
                                                  * factorial(c2 + p) * factorial(c2 - p)) ** 0.5)
        self._angcache = {}
        self._cgmat = {}
        self._ops = None
        self.gpu_kernel_launches = 0
        self.last_sweep = None

    # ------------------------------------------------------------------ helpers
    def _E1(self, n, l, j):
        return self.atom1.getEnergy(n, l, j, s=self.s1)

    def _E2(self, n, l, j):
        return self.atom2.getEnergy(n, l, j, s=self.s2)

    def _getEnergyDefect(self, n, l, j, nn, ll, jj, n1, l1, j1, n2, l2, j2):
        """calculations_atom_pairstate.py:712-741 (J)."""
        return C_e * (self._E1(n1, l1, j1) + self._E2(n2, l2, j2)
                      - self._E1(n, l, j) - self._E2(nn, ll, jj))

    def _w6j(self, l, s, j, j1, c, l1):
        # the table is only consulted for l, l1 <= wignerPrecalJmax = 23 (wigner.py:306-315);
        # beyond it the reference evaluates the true (non-zero) 6j
        if self.reference_6j_table_quirk and (l1 - l) == 2 and l <= 23 and l1 <= 23:
            return 0.0
        return Wigner6j(l, s, j, j1, c, l1)

    def _cg_matrix(self, j, c, p, j1):
        """A_p[m1' , m] = CG(j, m, c, p, j1, m1'), rows m1' = -j1..j1, cols m = -j..j."""
        key = (j, c, p, j1)
        if key not in self._cgmat:
            A = np.zeros((round(2 * j1 + 1), round(2 * j + 1)))
            for b in range(A.shape[1]):
                m = -j + b
                m1 = m + p
                if abs(m1) <= j1 + 0.1:
                    A[round(m1 + j1), b] = CG(j, m, c, p, j1, m1)
            self._cgmat[key] = A
        return self._cgmat[key]

    def _angular_factors(self, l, j, ll, jj, l1, j1, l2, j2):
        """elem and the per-polarisation single-atom matrices of
        `__getAngularMatrix_M` (calculations_atom_pairstate.py:410-520):
            d = elem * sum_p fcp[c1,c2,p] kron(A_p, B_p)."""
        key = (l, j, ll, jj, l1, j1, l2, j2)
        if key in self._angcache:
            return self._angcache[key]
        dl, dj = abs(l - l1), abs(j - j1)
        c1 = 1 if (dl == 1 and dj < 1.1) else (2 if dl in (0, 1, 2) else None)
        dl, dj = abs(ll - l2), abs(jj - j2)
        c2 = 1 if (dl == 1 and dj < 1.1) else (2 if dl in (0, 1, 2) else None)
        if c1 is None or c2 is None:
            raise ValueError("error in __getAngularMatrix_M")
        out = None
        if not (c1 > self.interactionsUpTo or c2 > self.interactionsUpTo):
            elem = ((-1.0) ** (j + jj + self.s1 + self.s2 + l1 + l2)
                    * CG(l, 0, c1, 0, l1, 0) * CG(ll, 0, c2, 0, l2, 0))
            elem = elem * sqrt((2.0 * l + 1.0) * (2.0 * ll + 1.0)) \
                * sqrt((2.0 * j + 1.0) * (2.0 * jj + 1.0))
            elem = elem * self._w6j(l, self.s1, j, j1, c1, l1) * self._w6j(ll, self.s2, jj, j2, c2, l2)
            terms = []
            lim = min(c1, c2)
            for p in range(-lim, lim + 1):
                f = self.fcp[c1, c2, p + self.interactionsUpTo]
                terms.append((f, self._cg_matrix(j, c1, p, j1), self._cg_matrix(jj, c2, -p, j2)))
            out = (elem, terms)
        self._angcache[key] = out
        return out

    def _getAngularMatrix_M(self, l, j, ll, jj, l1, j1, l2, j2):
        """Dense angular block, same layout as the reference's (rows: coupled state)."""
        af = self._angular_factors(l, j, ll, jj, l1, j1, l2, j2)
        shape = (round((2 * j1 + 1) * (2 * j2 + 1)), round((2 * j + 1) * (2 * jj + 1)))
        if af is None or af[0] == 0.0:
            return np.zeros(shape)
        elem, terms = af
        return elem * sum(f * _kron(A, B) for f, A, B in terms)

    def _isCoupled(self, n, l, j, nn, ll, jj, n1, l1, j1, n2, l2, j2, limit):
        """Scalar form of calculations_atom_pairstate.py:625-710 (returns c1+c2 or False)."""
        if not (abs(self._getEnergyDefect(n, l, j, nn, ll, jj, n1, l1, j1, n2, l2, j2)) / C_h < limit):
            return False
        if n == n1 and nn == n2 and l == l1 and ll == l2 and j == j1 and jj == j2:
            return False

        def forb(dl, ja, jb):
            return dl != 1 and ((abs(ja - 0.5) < 0.1 and abs(jb - 0.5) < 0.1)
                                or (abs(ja) < 0.1 and abs(jb - 1) < 0.1)
                                or (abs(ja - 1) < 0.1 and abs(jb) < 0.1))
        if forb(abs(l1 - l), j, j1) or forb(abs(l2 - ll), jj, j2):
            return False
        if (abs(j) < 0.1 and abs(j1) < 0.1) or (abs(jj) < 0.1 and abs(j2) < 0.1):
            return False
        if (abs(l) < 0.1 and abs(l1) < 0.1) or (abs(ll) < 0.1 and abs(l2) < 0.1):
            return False
        c1 = _coupling_order(l, j, l1, j1, self.interactionsUpTo)
        c2 = _coupling_order(ll, jj, l2, j2, self.interactionsUpTo)
        if c1 == 0 or c2 == 0:
            return False
        return c1 + c2

    def getC6perturbatively(self, theta, phi, nRange, energyDelta, degeneratePerturbation=False):
        """C6 from second-order perturbation theory, GHz um^6
        (calculations_atom_pairstate.py:1144-1440).  The radial elements of all intermediate
        pair states come from one batched GPU run; the final (2j+1)(2jj+1) eigenproblem runs on
        the GPU eigensolver."""
        if abs(phi) >= 1e-9:
            raise NotImplementedError("phi != 0 makes the interaction matrix complex-Hermitian")
        wgd = WignerDmatrix(theta, phi)
        Lmod2 = (self.l + self.ll) % 2
        lmin1 = self.l - 1 if self.l - 1 >= 0 else 1
        lmin2 = self.ll - 1 if self.ll - 1 >= 0 else 1
        dim = round((2 * self.j + 1) * (2 * self.jj + 1))
        inter = np.zeros((dim, dim))
        cand = []
        for n1 in range(max(self.n - nRange, 1), self.n + nRange + 1):
            for n2 in range(max(self.nn - nRange, 1), self.nn + nRange + 1):
                for l1 in range(lmin1, min(self.l + 2, n1), 2):
                    for l2 in range(lmin2, min(self.ll + 2, n2), 2):
                        if (l1 + l2) % 2 != Lmod2:
                            continue
                        j1 = l1 - self.s1
                        while j1 < -0.1:
                            j1 += 2 * self.s1
                        while j1 <= l1 + self.s1 + 0.1:
                            j2 = l2 - self.s2
                            while j2 < -0.1:
                                j2 += 2 * self.s2
                            while j2 <= l2 + self.s2 + 0.1:
                                coupled = self._isCoupled(self.n, self.l, self.j, self.nn, self.ll,
                                                          self.jj, n1, l1, j1, n2, l2, j2, energyDelta)
                                if (coupled and (not (self.interactionsUpTo == 1) or (Lmod2 == ((l1 + l2) % 2)))
                                        and (n1 >= self.atom1.groundStateN or [n1, l1, j1] in self.atom1.extraLevels)
                                        and (n2 >= self.atom2.groundStateN or [n2, l2, j2] in self.atom2.extraLevels)):
                                    cand.append((n1, l1, j1, n2, l2, j2, coupled))
                                j2 = j2 + 1.0
                            j1 = j1 + 1.0
        # one GPU batch per atom for all intermediate states
        for atom, spin, mine, col in ((self.atom1, self.s1, (self.n, self.l, self.j), 0),
                                      (self.atom2, self.s2, (self.nn, self.ll, self.jj), 3)):
            dip, quad = [], []
            for c in cand:
                other = (c[col], c[col + 1], c[col + 2])
                order = _coupling_order(mine[1], mine[2], other[1], other[2], self.interactionsUpTo)
                (dip if order == 1 else quad).append(mine + other)
            before = atom.gpu_kernel_launches
            if getattr(atom, "divalent", False):
                atom.precompute_radial(dipole_pairs=dip, quadrupole_pairs=quad, s=spin)
            else:
                atom.precompute_radial(dipole_pairs=dip, quadrupole_pairs=quad)
            self.gpu_kernel_launches += atom.gpu_kernel_launches - before
        a0 = physical_constants["Bohr radius"][0]
        for (n1, l1, j1, n2, l2, j2, coupled) in cand:
            defect = self._getEnergyDefect(self.n, self.l, self.j, self.nn, self.ll, self.jj,
                                           n1, l1, j1, n2, l2, j2) / C_h * 1.0e-9          # GHz
            if abs(defect) < 1e-10:
                raise ValueError("The requested pair-state is dipole coupled resonatly (energy "
                                 "defect = 0) to other pair-states. Aborting pertubative calculation.")
            r1 = self.atom1.getRadialCoupling(self.n, self.l, self.j, n1, l1, j1, s=self.s1)
            r2 = self.atom2.getRadialCoupling(self.nn, self.ll, self.jj, n2, l2, j2, s=self.s2)
            strength = (C_e ** 2 / (4.0 * pi * epsilon_0) * r1 * r2 * a0 ** coupled) \
                * (1.0e-9 * (1.0e6) ** 3 / C_h)                                          # GHz um^3
            d = self._getAngularMatrix_M(self.l, self.j, self.ll, self.jj, l1, j1, l2, j2)
            inter += d.T.dot(d) * abs(strength) ** 2 / defect
        rot = np.kron(wgd.get(self.j), wgd.get(self.jj))
        inter = (rot.dot(inter.dot(rot.conj().T))).real
        inter = 0.5 * (inter + inter.T)
        value, vecs = _sweep.syevd_batched(inter, device=getattr(self.atom1, "device", None))
        self.gpu_kernel_launches += 1
        value, vectors = value[0], vecs[0].T
        state = np.zeros(dim)
        state[round(self.j + self.m1) * round(2 * self.jj + 1) + round(self.jj + self.m2)] = 1.0
        if not degeneratePerturbation:
            for i, v in enumerate(vectors):
                if abs(np.vdot(v, state)) > 1 - 1e-9:
                    return value[i]
            return float(state.dot(inter.dot(state)))
        return np.real(value), vectors

    # ------------------------------------------------------------------ C6 on (theta, phi) grids
    def _ljCoupledCheck(self, l, j, l1, j1, s):
        """Is (l, j) -> (l1, j1) dipole (or, with interactionsUpTo = 2, quadrupole) coupled?
        Selection rules of calculations_atom_pairstate.py:1490-1547, including its use of
        Python's round-half-even (`round(j - 0.5)` is 0 for j = 1/2 AND for j = 1)."""
        if (l == l1 and j == j1) or (abs(j) < 0.1 and abs(j1) < 0.1) or (abs(l) < 0.1 and abs(l1) < 0.1):
            return False
        if not (abs(round(l - j)) < s + 0.1 and abs(round(l1 - j1)) < s + 0.1):
            return False
        dl, dj = abs(round(l - l1)), abs(round(j - j1))
        if dl == 1 and dj in (0, 1):
            return True
        if self.interactionsUpTo != 2 or dl not in (0, 2) or dj not in (0, 1, 2):
            return False

        def is0(x):
            return abs(x) < 0.1
        forbidden = ((is0(j) and is0(round(j1 - 1))) or (is0(round(j - 1)) and is0(j1))
                     or (is0(round(j - 0.5)) and is0(round(j1 - 0.5)))
                     or (is0(l) and is0(round(l1 - 1))) or (is0(round(l - 1)) and is0(l1)))
        return not forbidden

    def _findAllCoupledAngularMomentumStates(self, l, j, s1, ll, jj, s2, stateHopping=False):
        """All second-order paths (l,j; ll,jj) -> (l1,j1; l2,j2) -> final, final = initial or,
        for state hopping, the swapped pair (calculations_atom_pairstate.py:1549-1639); tuples
        (l, j, ll, jj, l1, j1, l2, j2, l', j', ll', jj') in the reference's loop order."""
        up = self.interactionsUpTo
        legs1 = [(l1, float(j1)) for l1 in range(max(0, l - up), l + up + 1)
                 for j1 in np.arange(max(s1, j - up), j + up + 1, 1) if self._ljCoupledCheck(l, j, l1, j1, s1)]
        legs2 = [(l2, float(j2)) for l2 in range(max(0, ll - up), ll + up + 1)
                 for j2 in np.arange(max(s2, jj - up), jj + up + 1, 1) if self._ljCoupledCheck(ll, jj, l2, j2, s2)]
        out = []
        for l1, j1 in legs1:
            for l2, j2 in legs2:
                if not stateHopping:
                    out.append((l, j, ll, jj, l1, j1, l2, j2, l, j, ll, jj))
                elif self._ljCoupledCheck(l1, j1, ll, jj, s1) and self._ljCoupledCheck(l2, j2, l, j, s2):
                    out.append((l, j, ll, jj, l1, j1, l2, j2, ll, jj, l, j))
        return out

    def _getC6contributions_lj(self, nRange, energyDelta, stateHopping=False):
        """Per angular path the radial sum V_lj = sum_{n', n''} |V1 V2| / defect in GHz um^6
        (calculations_atom_pairstate.py:1641-1789).  All radial elements of all paths are
        computed in ONE batched GPU run per atom before the sums are formed."""
        paths = self._findAllCoupledAngularMomentumStates(self.l, self.j, self.s1, self.ll, self.jj, self.s2,
                                                          stateHopping=stateHopping)
        n1s = range(max(self.n - nRange, 1), self.n + nRange + 1)
        n2s = range(max(self.nn - nRange, 1), self.nn + nRange + 1)
        a1, a2 = self.atom1, self.atom2
        ok_hop = (not stateHopping) or (self.nn >= a1.groundStateN and self.n >= a2.groundStateN)
        terms = []                     # (path index, n', n'', defect)
        for pi_, (l, j, ll, jj, l2, j2, ll2, jj2, l3, j3, ll3, jj3) in enumerate(paths):
            for n2 in n1s:
                for nn2 in n2s:
                    # like the reference, `[n, l, j] in extraLevels` compares a list with tuples
                    nCheck = ((n2 >= a1.groundStateN or [n2, l2, j2] in a1.extraLevels)
                              and (nn2 >= a2.groundStateN or [nn2, ll2, jj2] in a2.extraLevels) and ok_hop)
                    defect = self._getEnergyDefect(self.n, l, j, self.nn, ll, jj, n2, l2, j2, nn2, ll2, jj2) \
                        / C_h * 1e-9
                    if abs(defect) < 1e-10:
                        raise ValueError("The requested pair-state is dipole coupled resonatly (energy "
                                         "defect = 0) to other pair-states. Aborting pertubative calculation.")
                    if abs(defect) < energyDelta * 10 ** -9 and nCheck:
                        terms.append((pi_, n2, nn2, defect))
        # batch seam: every leg of every term, per atom
        want = {id(a1): ([], [], a1, self.s1), id(a2): ([], [], a2, self.s2)}

        def need(atom, sa, sb):
            order = 1 if (abs(sa[1] - sb[1]) == 1 and abs(sa[2] - sb[2]) < 1.1) else 2
            want[id(atom)][order - 1].append(tuple(sa) + tuple(sb))
        for pi_, n2, nn2, _d in terms:
            l, j, ll, jj, l2, j2, ll2, jj2, l3, j3, ll3, jj3 = paths[pi_]
            need(a1, (self.n, l, j), (n2, l2, j2))
            need(a2, (self.nn, ll, jj), (nn2, ll2, jj2))
            if stateHopping:
                need(a1, (n2, l2, j2), (self.nn, l3, j3))
                need(a2, (nn2, ll2, jj2), (self.n, ll3, jj3))
        for dip, quad, atom, spin in want.values():
            before = atom.gpu_kernel_launches
            if getattr(atom, "divalent", False):
                atom.precompute_radial(dipole_pairs=dip, quadrupole_pairs=quad, s=spin)
            else:
                atom.precompute_radial(dipole_pairs=dip, quadrupole_pairs=quad)
            self.gpu_kernel_launches += atom.gpu_kernel_launches - before
        a0 = physical_constants["Bohr radius"][0]
        unit = 1.0e-9 * (1.0e6) ** 3 / C_h                                    # J m^3 -> GHz um^3

        def strength(sa1, sb1, sa2, sb2):
            """_atomLightAtomCoupling (alkali_atom_functions.py:4066-4130) in GHz um^3."""
            c = 0
            for (x, y) in ((sa1, sb1), (sa2, sb2)):
                c += 1 if (abs(x[1] - y[1]) == 1 and abs(x[2] - y[2]) < 1.1) else 2
            r1 = a1.getRadialCoupling(*sa1, *sb1, s=self.s1)
            r2 = a2.getRadialCoupling(*sa2, *sb2, s=self.s2)
            return C_e ** 2 / (4.0 * pi * epsilon_0) * r1 * r2 * a0 ** c * unit
        V = [0.0] * len(paths)
        for pi_, n2, nn2, defect in terms:
            l, j, ll, jj, l2, j2, ll2, jj2, l3, j3, ll3, jj3 = paths[pi_]
            v1 = strength((self.n, l, j), (n2, l2, j2), (self.nn, ll, jj), (nn2, ll2, jj2))
            v2 = v1 if not stateHopping else strength((n2, l2, j2), (self.nn, l3, j3),
                                                      (nn2, ll2, jj2), (self.n, ll3, jj3))
            V[pi_] += abs(v1 * v2) / defect
        return [[*pth, float(v)] for pth, v in zip(paths, V)]

    def _getPerturbativeC6Matrix_lj(self, ljInteractions):
        """m_j-resolved interaction matrix sum_paths V_lj d2 d1 (calculations_atom_pairstate.py:1791-1818)."""
        Imat = 0
        for vals in ljInteractions:
            d1 = self._getAngularMatrix_M(*vals[0:8])
            d2 = self._getAngularMatrix_M(*vals[4:12])
            Imat = Imat + vals[-1] * d2.dot(d1)
        return np.array(Imat)

    @staticmethod
    def _getAngularBasisRotationMatrix(j1, j2):
        """Permutation between the |m_j1> x |m_j2> and |m_j2> x |m_j1> product bases
        (calculations_atom_pairstate.py:1462-1488)."""
        d1, d2 = round(2 * j1 + 1), round(2 * j2 + 1)
        P = np.zeros((d1 * d2, d1 * d2), dtype=np.complex128)
        for i in range(d2):
            for jx in range(d1):
                P[i * d1 + jx, jx * d2 + i] = 1.0
        return P

    def getC6perturbatively_anglePairs(self, anglePairs, nRange, energyDelta, degeneratePerturbation=False,
                                       returnInteractionMatrix=False):
        """Second-order C6 (GHz um^6) for a list of (theta, phi) orientations
        (calculations_atom_pairstate.py:2065-2287): the radial sums and the m_j-resolved interaction
        matrices are formed once, rotated per orientation with Wigner D matrices, and -- for
        degeneratePerturbation=True -- ALL orientations are diagonalised in one batched GPU
        eigensolve.  phi != 0 gives complex-Hermitian matrices; they are solved through the real
        symmetric embedding [[Re, -Im], [Im, Re]] (each eigenvalue appears twice; one complex
        eigenvector is rebuilt per eigenvalue).  Return values as in the reference."""
        state1 = [self.n, self.l, self.j, self.s1]
        state2 = [self.nn, self.ll, self.jj, self.s2]
        same = state1 == state2
        C6, C6hop, eigenvectors, matrices = [], [], [], []
        degenerateStates = [[self.n, self.l, self.j, self.nn, self.ll, self.jj],
                            [self.nn, self.ll, self.jj, self.n, self.l, self.j]]
        Imat11 = self._getPerturbativeC6Matrix_lj(self._getC6contributions_lj(nRange, energyDelta, False))
        if not same:
            Imat21 = self._getPerturbativeC6Matrix_lj(self._getC6contributions_lj(nRange, energyDelta, True))
            B12 = self._getAngularBasisRotationMatrix(self.jj, self.j)
        d2 = round(2 * self.jj + 1)
        i1 = round(self.j + self.m1) * d2 + round(self.jj + self.m2)                  # |j m1> x |jj m2>
        i2 = round(self.jj + self.m2) * round(2 * self.j + 1) + round(self.j + self.m1)
        full = []
        for angle1, angle2 in anglePairs:
            wgd = WignerDmatrix(angle1, angle2)
            R1 = np.kron(wgd.get(self.j), wgd.get(self.jj))
            I11 = R1.dot(Imat11.dot(R1.conj().T))
            if not same:
                R2 = np.kron(wgd.get(self.jj), wgd.get(self.j))
                I21 = R2.dot(Imat21.dot(R1.conj().T))
            if not degeneratePerturbation:
                C6.append(float(np.real(I11[i1, i1])))
                if not same:
                    C6hop.append(float(np.real(I21[i2, i1])))
            if degeneratePerturbation or returnInteractionMatrix:
                Irot = I11 if same else np.block([[I11, B12.dot(I21.dot(B12))],
                                                  [I21, B12.T.dot(I11.dot(B12))]])
                full.append(np.array(Irot))
                if returnInteractionMatrix:
                    matrices.append(np.array(Irot))
        if degeneratePerturbation:
            # numpy.linalg.eigh reads the lower triangle only: make that explicit, then one batch
            herm = [np.tril(M) + np.tril(M, -1).conj().T for M in full]
            cplx = any(np.abs(M.imag).max() > 0 for M in herm)
            if not cplx:
                w, v = _sweep.syevd_batched(np.stack([M.real for M in herm]),
                                            device=getattr(self.atom1, "device", None))
                for t in range(len(herm)):
                    C6.append(np.array(w[t]))
                    eigenvectors.append(np.array(v[t].T, dtype=np.complex128))
            else:
                emb = np.stack([np.block([[M.real, -M.imag], [M.imag, M.real]]) for M in herm])
                w, v = _sweep.syevd_batched(emb, device=getattr(self.atom1, "device", None))
                for t, M in enumerate(herm):
                    ev, vecs = _unembed_hermitian(w[t], v[t], M.shape[0])
                    C6.append(ev)
                    eigenvectors.append(vecs)
            self.gpu_kernel_launches += 1
        if not degeneratePerturbation and not returnInteractionMatrix:
            return C6, C6hop
        if not degeneratePerturbation:
            return C6, C6hop, matrices
        if not returnInteractionMatrix:
            return C6, eigenvectors, degenerateStates
        return C6, eigenvectors, matrices, degenerateStates

    # ------------------------------------------------------------------ channels
    def _makeRawMatrix2(self, n, l, j, nn, ll, jj, k, lrange, limit, limitBasisToMj,
                        progressOutput=False, debugOutput=False):
        """Channel list + radial couplings (calculations_atom_pairstate.py:743-1048);
        same enumeration order, same predicates."""
        up = self.interactionsUpTo
        Lmod2 = (l + ll) % 2
        l1start, l2start = max(l - up, 0), max(ll - up, 0)
        states, opi = [], 0
        E0 = None
        for n1 in range(max(n - k, 1), n + k + 1):
            for n2 in range(max(nn - k, 1), nn + k + 1):
                l1max = min(max(l + up, lrange) + 1, n1 - 1)
                for l1 in range(l1start, l1max):
                    l2max = min(max(ll + up, lrange) + 1, n2 - 1)
                    for l2 in range(l2start, l2max):
                        j1 = l1 - self.s1
                        while j1 < -0.1:
                            j1 += 2 * self.s1
                        while j1 <= l1 + self.s1 + 0.1:
                            j2 = l2 - self.s2
                            while j2 < -0.1:
                                j2 += 2 * self.s2
                            while j2 <= l2 + self.s2 + 0.1:
                                ed = self._getEnergyDefect(n, l, j, nn, ll, jj,
                                                           n1, l1, j1, n2, l2, j2) / C_h
                                if (abs(ed) < limit
                                        and (not (up == 1) or (Lmod2 == ((l1 + l2) % 2)))
                                        and ((not limitBasisToMj)
                                             or (j1 + j2 + 0.1 > self.m1 + self.m2))
                                        and (n1 >= self.atom1.groundStateN
                                             or [n1, l1, j1] in self.atom1.extraLevels)
                                        and (n2 >= self.atom2.groundStateN
                                             or [n2, l2, j2] in self.atom2.extraLevels)):
                                    if (n == n1 and nn == n2 and l == l1 and ll == l2
                                            and j == j1 and jj == j2):
                                        opi = len(states)
                                    states.append([n1, l1, j1, n2, l2, j2])
                                j2 = j2 + 1.0
                            j1 = j1 + 1.0
        dim = len(states)
        opZeemanShift = ((self.atom1.getZeemanEnergyShift(self.l, self.j, self.m1, self.Bz, s=self.s1)
                          + self.atom2.getZeemanEnergyShift(self.ll, self.jj, self.m2, self.Bz, s=self.s2))
                         / C_h * 1.0e-9)
        st = np.array(states, dtype=float).reshape(dim, 6)
        E1 = np.array([self._E1(int(s_[0]), int(s_[1]), s_[2]) for s_ in st])
        E2 = np.array([self._E2(int(s_[3]), int(s_[4]), s_[5]) for s_ in st])
        for i in range(dim):
            # __getEnergyDefect(opi -> i), same summation order as the reference
            ed = C_e * (E1[i] + E2[i] - E1[opi] - E2[opi]) / C_h * 1.0e-9 - opZeemanShift
            states[i].append(ed)

        # ---- vectorised __isCoupled over all i < j (calculations_atom_pairstate.py:625-710)
        nA, lA, jA, nB, lB, jB = (st[:, c] for c in range(6))
        I, J = np.triu_indices(dim, 1)
        defect = C_e * (E1[J] + E2[J] - E1[I] - E2[I])
        ok = np.abs(defect) / C_h < limit
        dl1, dl2 = np.abs(lA[J] - lA[I]), np.abs(lB[J] - lB[I])
        dj1, dj2 = np.abs(jA[J] - jA[I]), np.abs(jB[J] - jB[I])

        def forbidden(dl, ja, jb):
            return (dl != 1) & (((np.abs(ja - 0.5) < 0.1) & (np.abs(jb - 0.5) < 0.1))
                                | ((np.abs(ja) < 0.1) & (np.abs(jb - 1) < 0.1))
                                | ((np.abs(ja - 1) < 0.1) & (np.abs(jb) < 0.1)))
        ok &= ~(forbidden(dl1, jA[I], jA[J]) | forbidden(dl2, jB[I], jB[J]))
        ok &= ~((np.abs(jA[I]) < 0.1) & (np.abs(jA[J]) < 0.1))
        ok &= ~((np.abs(jB[I]) < 0.1) & (np.abs(jB[J]) < 0.1))
        ok &= ~((np.abs(lA[I]) < 0.1) & (np.abs(lA[J]) < 0.1))
        ok &= ~((np.abs(lB[I]) < 0.1) & (np.abs(lB[J]) < 0.1))

        def order(dl, dj):
            c = np.zeros(len(dl), dtype=int)
            dip = (dl == 1) & (dj < 1.1)
            quad = (~dip) & ((dl == 0) | (dl == 1) | (dl == 2)) & (dj < 2.1) & (2 <= up)
            c[dip] = 1
            c[quad] = 2
            return c
        c1, c2 = order(dl1, dj1), order(dl2, dj2)
        ok &= (c1 > 0) & (c2 > 0)
        ok &= (np.abs(nA[I] - nA[J]) <= k) & (np.abs(nB[I] - nB[J]) <= k)
        I, J, c1, c2 = I[ok], J[ok], c1[ok], c2[ok]

        # ---- all radial elements of both atoms in one GPU batch per atom (vector form: the memo is
        # filled by one precompute_radial call per atom, then read back for all couplings at once)
        def radial(atom, nq, lq, jq, cs, spin):
            P = np.stack([nq[I], lq[I], jq[I], nq[J], lq[J], jq[J]], axis=1)
            dip, quad = cs == 1, cs == 2
            before = atom.gpu_kernel_launches
            if getattr(atom, "divalent", False):
                atom.precompute_radial(dipole_pairs=P[dip], quadrupole_pairs=P[quad], s=spin)
            else:
                atom.precompute_radial(dipole_pairs=P[dip], quadrupole_pairs=P[quad])
            self.gpu_kernel_launches += atom.gpu_kernel_launches - before
            S = np.stack([nq, lq, jq], axis=1)
            r = np.zeros(len(I))
            r[dip] = atom.radial_elements_indexed(S, I[dip], J[dip], 1, s=spin)
            r[quad] = atom.radial_elements_indexed(S, I[quad], J[quad], 2, s=spin)
            return r
        r1 = radial(self.atom1, nA, lA, jA, c1, self.s1)
        r2 = radial(self.atom2, nB, lB, jB, c2, self.s2)

        # _atomLightAtomCoupling (alkali_atom_functions.py:4123-4129), GHz m^(c1+c2+1)
        a0 = physical_constants["Bohr radius"][0]
        a0pow = np.array([a0 ** c for c in range(5)])
        val = (C_e ** 2 / (4.0 * pi * epsilon_0) * r1 * r2 * a0pow[c1 + c2]) / C_h * 1.0e-9
        cons = []
        for o in range(2 * up - 1):
            m = (c1 + c2 - 2) == o
            cons.append((val[m], I[m], J[m]))
        coupling = [csr_matrix((c[0], (c[1], c[2])), shape=(dim, dim)) for c in cons]
        return states, coupling

    # ------------------------------------------------------------------ defineBasis
    def defineBasis(self, theta, phi, nRange, lrange, energyDelta, Bz=0, progressOutput=False,
                    debugOutput=False):
        """calculations_atom_pairstate.py:2920-3244."""
        if abs(phi) >= 1e-9:
            raise NotImplementedError(
                "phi != 0 gives a complex-Hermitian interaction matrix "
                "(calculations_atom_pairstate.py:3201-3204); the B200 eigensolver is "
                "real-symmetric only")
        self.theta, self.phi = theta, phi
        self.nRange, self.lrange, self.energyDelta, self.Bz = nRange, lrange, energyDelta, Bz
        self.basisStates, self.matrixElement = [], []
        self._ops = None
        wgd = WignerDmatrix(theta, phi)
        limitBasisToMj = theta < 0.001
        originalMj = self.m1 + self.m2
        self.channel, self.coupling = self._makeRawMatrix2(
            self.n, self.l, self.j, self.nn, self.ll, self.jj, nRange, lrange, energyDelta,
            limitBasisToMj, progressOutput=progressOutput, debugOutput=debugOutput)
        nch = len(self.channel)
        opi = 0
        index = np.zeros(nch + 1, dtype=np.int64)   # cast to the reference's int16 after the size check
        prod_idx = []           # per channel: index of each basis state in the full m1 x m2 product
        mcache = {}

        def mgrid(jq):          # m = j, j-1, ..., -j exactly as the reference's np.linspace(j, -j, 2j+1)
            if jq not in mcache:
                mcache[jq] = list(np.linspace(jq, -jq, round(1 + 2 * jq)))
            return mcache[jq]
        for i, ch in enumerate(self.channel):
            index[i] = len(self.basisStates)
            pidx = []
            ja, jb = ch[2], ch[5]
            for m1c in mgrid(ja):
                for m2c in mgrid(jb):
                    if (not limitBasisToMj) or (abs(originalMj - m1c - m2c) < 0.1):
                        self.basisStates.append([ch[0], ch[1], ch[2], m1c, ch[3], ch[4], ch[5], m2c])
                        self.matrixElement.append(i)
                        pidx.append(round(m1c + ja) * round(2 * jb + 1) + round(m2c + jb))
                        if (abs(ch[0] - self.n) < 0.1 and abs(ch[1] - self.l) < 0.1
                                and abs(ch[2] - self.j) < 0.1 and abs(m1c - self.m1) < 0.1
                                and abs(ch[3] - self.nn) < 0.1 and abs(ch[4] - self.ll) < 0.1
                                and abs(ch[5] - self.jj) < 0.1 and abs(m2c - self.m2) < 0.1):
                            opi = len(self.basisStates) - 1
            prod_idx.append(np.array(pidx, dtype=np.int64))
        index[-1] = len(self.basisStates)
        dimension = len(self.basisStates)
        if dimension > 32767:
            raise ValueError("basis too large (the reference's index array is int16)")
        self.index = index.astype(np.int16)

        # ---- matDiagonal (calculations_atom_pairstate.py:3133-3159)
        diag = np.zeros(dimension)
        zcache = {}
        for ii, ch in enumerate(self.channel):
            off = 0.00000001
            for i in range(self.index[ii], self.index[ii + 1]):
                b = self.basisStates[i]
                kz = (b[1], b[2], b[3], b[5], b[6], b[7])
                if kz not in zcache:
                    zcache[kz] = ((self.atom1.getZeemanEnergyShift(b[1], b[2], b[3], self.Bz, s=self.s1)
                                   + self.atom2.getZeemanEnergyShift(b[5], b[6], b[7], self.Bz, s=self.s2))
                                  / C_h * 1.0e-9)
                diag[i] = ch[6] + zcache[kz] + off
                off += 0.00000001
        idx = np.arange(dimension)
        self.matDiagonal = csr_matrix((diag, (idx, idx)), shape=(dimension, dimension))

        # ---- matR (closed form of calculations_atom_pairstate.py:3116-3217)
        trivial = wgd.trivial
        self.matR = []
        self._matR_coo = []
        index64 = index                                     # int64 copy of self.index
        nstates_ch = np.diff(index64)
        # the angular block and its sparsity pattern depend on the eight (l, j) numbers only (the m
        # sub-levels kept per channel follow from j1, j2): channels are classed by (l1, j1, l2, j2),
        # one block per pair of classes, and all coupled channel pairs of a class pair are expanded
        # into matrix entries in one numpy step
        lj = np.array([[ch[1], ch[2], ch[4], ch[5]] for ch in self.channel], dtype=np.float64).reshape(nch, 4)
        ljc, cls = (np.unique(lj, axis=0, return_inverse=True) if nch else (lj, np.zeros(0, dtype=np.int64)))
        cls = np.asarray(cls).ravel()
        first_of = np.zeros(len(ljc), dtype=np.int64)
        first_of[cls[::-1]] = np.arange(nch)[::-1]          # a representative channel per class
        blkcache = {}

        def block(ci, cj):
            """(rows within channel j, columns within channel i, values) of the angular block between the
            channel classes ci -> cj, entries with |value| <= 1e-5 dropped (pair:3198-3217); () if empty."""
            ii, jjc = int(first_of[ci]), int(first_of[cj])
            chi, chj = self.channel[ii], self.channel[jjc]
            af = self._angular_factors(chi[1], chi[2], chi[4], chi[5], chj[1], chj[2], chj[4], chj[5])
            if af is None or af[0] == 0.0:
                return ()
            elem, terms = af
            if trivial:
                B = elem * sum(f * _kron(A, Bm) for f, A, Bm in terms)
            else:
                D1, D2 = wgd.get(chi[2]), wgd.get(chi[5])
                D3, D4 = wgd.get(chj[2]), wgd.get(chj[5])
                B = elem * sum(f * _kron(D3 @ A @ D1.conj().T, D4 @ Bm @ D2.conj().T)
                               for f, A, Bm in terms)
                B = B.real
            blk = B[np.ix_(prod_idx[jjc], prod_idx[ii])]     # rows: j states, cols: i states
            jr, ic = np.nonzero(np.abs(blk) > 1.0e-5)
            return (jr, ic, blk[jr, ic]) if len(jr) else ()

        for c in self.coupling:
            rows, cols, vals = [], [], []
            c = c.tocsr()
            ci_all = np.repeat(np.arange(nch), np.diff(c.indptr))
            cj_all = np.asarray(c.indices, dtype=np.int64)
            keep = (nstates_ch[ci_all] > 0) & (nstates_ch[cj_all] > 0)
            ci_all, cj_all, rho_all = ci_all[keep], cj_all[keep], np.asarray(c.data)[keep]
            code = cls[ci_all] * len(ljc) + cls[cj_all]
            order = np.argsort(code, kind="stable")
            bounds = np.flatnonzero(np.diff(code[order])) + 1
            for grp in np.split(order, bounds) if len(order) else []:
                k0 = int(code[grp[0]])
                hit = blkcache.get(k0)
                if hit is None:
                    hit = blkcache[k0] = block(k0 // len(ljc), k0 % len(ljc))
                if not hit:
                    continue
                jr, ic, bv = hit
                v = (rho_all[grp][:, None] * bv[None, :]).ravel()
                gi = (index64[ci_all[grp]][:, None] + ic[None, :]).ravel()
                gj = (index64[cj_all[grp]][:, None] + jr[None, :]).ravel()
                rows.append(gi); cols.append(gj); vals.append(v)
                rows.append(gj); cols.append(gi); vals.append(v)
            if rows:
                rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
            else:
                rows, cols, vals = np.zeros(0, int), np.zeros(0, int), np.zeros(0)
            self.matR.append(csr_matrix((vals, (rows, cols)), shape=(dimension, dimension)))
            self._matR_coo.append((rows.astype(np.int32), cols.astype(np.int32), vals))
        self.originalPairStateIndex = opi

    # ------------------------------------------------------------------ diagonalise
    def _addState(self, n1, l1, j1, mj1, n2, l2, j2, mj2):
        """calculations_atom_pairstate.py:3729-3756."""
        if abs(self.s1 - 0.5) < 0.1:
            s = "|%s %d/2" % (printStateStringLatex(n1, l1, j1, s=self.s1), round(2 * mj1))
        else:
            s = "|%s %d" % (printStateStringLatex(n1, l1, j1, s=self.s1), round(mj1))
        if abs(self.s2 - 0.5) < 0.1:
            s += ",%s %d/2\\rangle" % (printStateStringLatex(n2, l2, j2, s=self.s2), round(2 * mj2))
        else:
            s += ",%s %d\\rangle" % (printStateStringLatex(n2, l2, j2, s=self.s2), round(mj2))
        return s

    @staticmethod
    def _greedy_assignment(S):
        """perm[c] = raw index assigned to position c: entries of S taken in decreasing order, a row and a
        column used once each (the reference's argsort + rowPicked / columnPicked scan,
        calculations_atom_pairstate.py:3460-3480).  Fast path: when every position has an overlap above
        1/2 with exactly one raw state (and vice versa), the scan would take exactly those k entries --
        it reaches them before any smaller one and none of them share a row or a column -- so they are
        assigned in one numpy pass.  Any other point (avoided crossings, states entering or leaving the
        window) runs the full scan, ties included, exactly as before."""
        k = S.shape[0]
        big_r, big_c = np.nonzero(S > 0.5)
        if (len(big_r) == k and np.bincount(big_r, minlength=k).max() == 1
                and np.bincount(big_c, minlength=k).max() == 1):
            perm = np.empty(k, dtype=np.int64)
            perm[big_c] = big_r
            return perm
        order_flat = np.argsort(S.ravel())
        rows_, cols_ = np.unravel_index(order_flat, (k, k))
        rowPicked = np.zeros(k, dtype=bool)
        colPicked = np.zeros(k, dtype=bool)
        perm = np.zeros(k, dtype=np.int64)
        j, done = k * k - 1, 0
        while done < k:
            while rowPicked[rows_[j]] or colPicked[cols_[j]]:
                j -= 1
            rowPicked[rows_[j]] = True
            colPicked[cols_[j]] = True
            perm[cols_[j]] = rows_[j]
            done += 1
        return perm

    @staticmethod
    def _adiabatic_sort(res):
        """Re-order every point's eigenpairs so that index i follows the state adiabatically
        (calculations_atom_pairstate.py:3452-3492): greedy assignment by decreasing squared
        overlap with the previous point's (already re-ordered) eigenvectors.  The overlaps of
        the raw eigenvector sets come from the GPU (res.overlap); the permutations compose
        sequentially on the host."""
        nR, k = res.y.shape
        prev = np.arange(k)                       # position -> raw index at the previous point
        for t in range(nR):
            S = res.overlap[t][:, prev]           # S[i, c] = overlap(raw i at t, position c at t-1)
            perm = PairStateInteractions._greedy_assignment(S)
            res.y[t] = res.y[t][perm]
            res.hl[t] = res.hl[t][perm]
            res.comp_c[t] = res.comp_c[t][perm]
            res.comp_i[t] = res.comp_i[t][perm]
            prev = perm

    def device_operators(self):
        """matDiagonal and matR staged on the GPU (dense), built once per defineBasis."""
        if self._ops is None:
            self._ops = _sweep.PairOperators(self.matDiagonal.diagonal(), self._matR_coo,
                                             len(self.basisStates),
                                             device=getattr(self.atom1, "device", None))
            self.gpu_kernel_launches += len(self._matR_coo)
        return self._ops

    def diagonalise(self, rangeR, noOfEigenvectors, drivingFromState=[0, 0, 0, 0, 0],
                    eigenstateDetuning=0.0, sortEigenvectors=False, progressOutput=False,
                    debugOutput=False, batch=None, group=None):
        """calculations_atom_pairstate.py:3246-3522.  `group` (extension): None = all R points
        on this GPU; sweep.WORLD or a torch.distributed group = points sharded over its ranks
        (collective call, every rank gets the full result)."""
        self.r = np.sort(rangeR)[::-1]
        dimension = len(self.basisStates)
        self.noOfEigenvectors = noOfEigenvectors
        self.y, self.highlight, self.composition = [], [], []
        if noOfEigenvectors >= dimension - 1:
            noOfEigenvectors = dimension - 1
            print("Warning: Requested number of eigenvectors >=dimension-1\n "
                  "ARPACK can only find up to dimension-1 eigenvectors, where "
                  "dimension is matrix dimension.\n")
            if noOfEigenvectors < 1:
                return
        drive_w = None
        self.maxCoupling = 0.0
        self.maxCoupledStateIndex = 0
        if drivingFromState[0] != 0:
            self.drivingFromState = drivingFromState
            n1, l1 = round(drivingFromState[0]), round(drivingFromState[1])
            j1, m1, q = drivingFromState[2], drivingFromState[3], drivingFromState[4]
            o = self.basisStates[self.originalPairStateIndex]
            coupling = []
            for i in range(dimension):
                b = self.basisStates[i]
                c = 0.0
                if (round(abs(b[5] - l1)) == 1 and abs(b[0] - o[0]) < 0.1 and abs(b[1] - o[1]) < 0.1
                        and abs(b[2] - o[2]) < 0.1 and abs(b[3] - o[3]) < 0.1):
                    c += self.atom2.getDipoleMatrixElement(n1, l1, j1, m1, round(b[4]), round(b[5]),
                                                           b[6], b[7], q, s=self.s2)
                c = abs(c) ** 2
                if c > self.maxCoupling:
                    self.maxCoupling, self.maxCoupledStateIndex = c, i
                coupling.append(c)
            drive_w = np.abs(np.array(coupling) / self.maxCoupling)
        ops = self.device_operators()
        res = _sweep.pair_sweep(ops, self.r, noOfEigenvectors, eigenstateDetuning * 1.0e-9,
                                self.originalPairStateIndex, drive_w=drive_w, topk=4,
                                contrib_max=0.95, batch=batch, want_overlap=bool(sortEigenvectors),
                                group=group)
        self.last_sweep = res
        self.gpu_kernel_launches += res.kernel_launches
        if sortEigenvectors:
            self._adiabatic_sort(res)
        self.y = [res.y[i] for i in range(len(self.r))]
        self.highlight = [res.hl[i] for i in range(len(self.r))]
        self.composition = PairCompositionView(self, res.comp_c, res.comp_i)

    # ------------------------------------------------------------------ fit helpers
    def _fit_range(self, rStart, rStop):
        """calculations_atom_pairstate.py:4026-4043 (r is sorted descending by diagonalise)."""
        fromRindex = -1
        toRindex = -1
        for br in range(len(self.r)):
            if (fromRindex == -1) and (self.r[br] >= rStart):
                fromRindex = br
            if self.r[br] > rStop:
                toRindex = br - 1
                break
        if (fromRindex != -1) and (toRindex == -1):
            toRindex = len(self.r) - 1
        if fromRindex == -1:
            print("\nERROR: could not find data for energy levels for interatomic")
            print("distances between %2.f and %.2f mu m.\n\n" % (rStart, rStop))
        return fromRindex, toRindex

    def getC6fromLevelDiagram(self, rStart, rStop, showPlot=False, minStateContribution=0.0):
        """C6 (GHz um^6) from a fit of |E(R)| of the state that carries most of the original
        pair state to log(C6/R^6 + offset) (calculations_atom_pairstate.py:3978-4112; plotting
        is out of scope).  The argmax over `highlight` per distance is vectorised."""
        from scipy.optimize import curve_fit
        fromRindex, toRindex = self._fit_range(rStart, rStop)
        if fromRindex == -1:
            return 0
        initialStateDetuning, initialStateDetuningX = [], []
        for br in range(fromRindex, toRindex + 1):
            hl = np.abs(np.asarray(self.highlight[br]))
            index = int(np.argmax(hl)) if len(hl) else -1
            if index != -1 and hl[index] > minStateContribution:
                initialStateDetuning.append(abs(self.y[br][index]))
                initialStateDetuningX.append(self.r[br])
        initialStateDetuning = np.log(np.array(initialStateDetuning))
        initialStateDetuningX = np.array(initialStateDetuningX)

        def c6fit(r, c6, offset):
            return np.log(c6 / r ** 6 + offset)
        try:
            popt, pcov = curve_fit(c6fit, initialStateDetuningX, initialStateDetuning, [1, 0])
        except Exception as ex:
            print(ex)
            print("ERROR: unable to find a fit for C6.")
            return False
        print("c6 = ", popt[0], " GHz /R^6 (mu m)^6")
        print("offset = ", popt[1])
        self.fitX = initialStateDetuningX
        self.fitY = initialStateDetuning
        self.fittedCurveY = np.array([c6fit(v, popt[0], popt[1]) for v in initialStateDetuningX])
        return popt[0]

    def _adiabatic_branch(self, fromRindex, toRindex, minStateContribution, selectBranch):
        """Track the original state from large to small R, stopping at a slope discontinuity
        (calculations_atom_pairstate.py:4189-4216, 4356-4376)."""
        det, detX = [], []
        discontinuityDetected = False
        for br in range(toRindex, fromRindex - 1, -1):
            index = -1
            maxPortion = minStateContribution
            for br2 in range(len(self.y[br])):
                if (abs(self.highlight[br][br2]) > maxPortion) and (
                        not selectBranch or (self.y[br][br2] * selectBranch > 0.0)):
                    index = br2
                    maxPortion = abs(self.highlight[br][br2])
            if len(detX) > 2:
                slope1 = (det[-1] - det[-2]) / (detX[-1] - detX[-2])
                slope2 = (abs(self.y[br][index]) - det[-1]) / (self.r[br] - detX[-1])
                if abs(slope2) > 3.0 * abs(slope1):
                    discontinuityDetected = True
            if (index != -1) and (not discontinuityDetected):
                det.append(abs(self.y[br][index]))
                detX.append(self.r[br])
        return det, detX

    def getC3fromLevelDiagram(self, rStart, rStop, showPlot=False, minStateContribution=0.0,
                              resonantBranch=+1):
        """C3 (GHz um^3), fit to log(C3/R^3 + offset) (calculations_atom_pairstate.py:4114-4291;
        like the reference the branch test multiplies by the boolean `selectBranch`)."""
        from scipy.optimize import curve_fit
        selectBranch = False
        if abs(self.l - self.ll) == 1:
            selectBranch = True
            resonantBranch = float(resonantBranch)
        fromRindex, toRindex = self._fit_range(rStart, rStop)
        if fromRindex == -1:
            return False
        det, detX = self._adiabatic_branch(fromRindex, toRindex, minStateContribution, selectBranch)
        det = np.log(np.array(det))
        detX = np.array(detX)

        def c3fit(r, c3, offset):
            return np.log(c3 / r ** 3 + offset)
        try:
            popt, pcov = curve_fit(c3fit, detX, det, [1, 0])
        except Exception as ex:
            print(ex)
            print("ERROR: unable to find a fit for C3.")
            return False
        print("c3 = ", popt[0], " GHz /R^3 (mu m)^3")
        print("offset = ", popt[1])
        self.fitX, self.fitY = detX, det
        self.fittedCurveY = np.array([c3fit(v, popt[0], popt[1]) for v in detX])
        return popt[0]

    def getVdwFromLevelDiagram(self, rStart, rStop, showPlot=False, minStateContribution=0.0):
        """van der Waals radius (um) (calculations_atom_pairstate.py:4293-4472)."""
        from scipy.optimize import curve_fit
        fromRindex, toRindex = self._fit_range(rStart, rStop)
        if fromRindex == -1:
            return False
        det, detX = self._adiabatic_branch(fromRindex, toRindex, minStateContribution, False)
        det = np.log(abs(np.array(det)))
        detX = np.array(detX)

        def vdwFit(r, offset, scale, vdw):
            return np.log(abs(offset + scale * (1.0 - np.sqrt(1.0 + (vdw / r) ** 6))
                              / (1.0 - np.sqrt(1 + vdw ** 6))))
        noOfPoints = len(detX)
        print("Data points to fit = ", noOfPoints)
        try:
            popt, pcov = curve_fit(vdwFit, detX, det, [0, det[noOfPoints // 2], detX[noOfPoints // 2]])
        except Exception as ex:
            print(ex)
            print("ERROR: unable to find a fit for van der Waals distance.")
            return False
        if (detX[0] < popt[2]) or (popt[2] < detX[-1]):
            print("WARNING: vdw radius seems to be outside the fitting range!")
            print("It's estimated to be around %.2f mu m from the current fit." % popt[2])
        print("Rvdw =  ", popt[2], " mu m")
        print("offset = ", popt[0], "\n scale = ", popt[1])
        self.fitX, self.fitY = detX, det
        self.fittedCurveY = np.array([vdwFit(v, popt[0], popt[1], popt[2]) for v in detX])
        return popt[2]
