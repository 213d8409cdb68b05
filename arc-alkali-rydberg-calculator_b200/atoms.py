"""Atom model for the hot path: energies, radial matrix elements, Zeeman shifts.

Host-side mirror of the parts of the reference's `AlkaliAtom`
(arc/alkali_atom_functions.py) and `DivalentAtom` (arc/divalent_atom_functions.py)
that the hot path calls, with the same method names, argument meaning and return
units.  Element INPUT DATA (model-potential parameters, quantum defects, NIST
levels, literature dipole elements) is read from data/atoms.json, exported from
the reference's element classes by tools/export_atom_data.py.

Radial matrix elements are computed on the GPU (radial.RadialBatch): a single
`getRadialMatrixElement` call is a batch of one; `precompute_radial` evaluates a
whole table in two kernel launches and fills the in-memory memo that stands in for
the reference's per-element sqlite memo (alkali_atom_functions.py:1057-1065).
"""
import json
import os
from math import floor, sqrt

import numpy as np
from scipy.constants import physical_constants, m_e as C_m_e

from .wigner import CG
from . import radial as _radial

_DATA = None
_HARTREE_EV = 27.211_386_245_981      # alkali_atom_functions.py:1072


def _data():
    global _DATA
    if _DATA is None:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "atoms.json")
        with open(p) as f:
            _DATA = json.load(f)
    return _DATA


def printStateLetter(l):
    return "SPDFGHIKLMNOQRTUVWXYZ"[l] if l < 21 else " l=%d" % l


def _unique_bounded(code):
    """np.unique(code, return_index=True, return_inverse=True) for non-negative int64 codes; when
    the codes are dense enough a lookup table replaces the sort (O(m) instead of O(m log m)).  The
    representative index is that of the LAST occurrence (any occurrence serves the callers)."""
    code = np.asarray(code, dtype=np.int64)
    if not len(code):
        z = np.zeros(0, dtype=np.int64)
        return z, z, z
    R = int(code.max()) + 1
    if R > max(1 << 22, 8 * len(code)):
        return np.unique(code, return_index=True, return_inverse=True)
    rep = np.full(R, -1, dtype=np.int64)
    rep[code] = np.arange(len(code))
    uc = np.flatnonzero(rep >= 0)
    rank = np.empty(R, dtype=np.int64)
    rank[uc] = np.arange(len(uc))
    return uc, rep[uc], rank[code]


class AlkaliAtom:
    """Data-backed atom.  `AlkaliAtom("Rubidium85")` or the named constructors below."""

    divalent = False

    def __init__(self, className, preferQuantumDefects=True, cpp_numerov=True, device=None):
        d = _data()
        className = d["aliases"].get(className, className)
        if className not in d["atoms"]:
            raise ValueError("unknown element class %r" % className)
        a = d["atoms"][className]
        self.className = className
        self.elementName = a["elementName"]
        self.Z, self.I, self.mass = a["Z"], a["I"], a["mass"]
        self.alpha, self.alphaC = a["alpha"], a["alphaC"]
        self.alpha_d_eff, self.alpha_q_eff = a["alpha_d_eff"], a["alpha_q_eff"]
        self.a1, self.a2, self.a3, self.a4, self.rc = a["a1"], a["a2"], a["a3"], a["a4"], a["rc"]
        self.gS, self.gL = a["gS"], a["gL"]
        self.groundStateN = a["groundStateN"]
        # tuples, exactly like the reference (alkali_atom_data.py:286-295).  The reference's
        # consumers test `[n, l, j] in atom.extraLevels` with a LIST (calculations_atom_single.py:
        # 749-751, calculations_atom_pairstate.py:824-829), which never equals a tuple, so states
        # below groundStateN are always dropped from Stark / pair bases.  Parity keeps that
        # (pinned by tests/golden/stark_*_lown_golden.npz from the live reference).
        self.extraLevels = [tuple(x) for x in a["extraLevels"]]
        self.minQuantumDefectN = a["minQuantumDefectN"]
        self.NISTdataLevels = a["NISTdataLevels"]
        self.scaledRydbergConstant = a["scaledRydbergConstant"]
        self.ionisationEnergy = a["ionisationEnergy"]
        self.quantumDefect = a["quantumDefect"]
        self.preferQuantumDefects = preferQuantumDefects
        self.cpp_numerov = cpp_numerov
        self.device = device
        self._raw = a
        if not self.divalent:
            self._saved = {(r[0], r[1]): r[2] for r in a.get("sEnergy", [])}
            self._lit = {}
            for r in a.get("literatureDME", []):
                key = tuple(r[:6])
                if key not in self._lit or r[7] < self._lit[key][1]:
                    self._lit[key] = (r[6], r[7])
        self._dipole = {}       # memo: (n1,l1,2j1,n2,l2,2j2[,s]) -> value, lower-energy state first
        self._quadrupole = {}
        self._energy = {}
        self.gpu_kernel_launches = 0
        # fast_radial=True routes batches of >= gram_threshold new elements through the tensor-core
        # Gram kernel (radial.gram_refined_device): 2-3x faster on large tables, but it uses one
        # common r grid instead of the lower-energy state's own grid, which moves strongly
        # cancelling elements (|value| < 1e-4 sqrt(<r^p>_a <r^p>_b)) by up to 3e-12 of that scale.
        # Default off: the one-warp-per-pair kernel reproduces the reference's formula exactly.
        self.fast_radial = False
        # radial_group (opt-in, collective): tables of >= 1024 new elements are sharded over the
        # ranks of this torch.distributed group (sweep.WORLD = the default group) and all-gathered
        # (radial.sharded_table); None = every process computes its own table, no collective
        self.radial_group = None
        self.gram_threshold = 64

    # pickling support (saveCalculation, alkali_atom_functions.py:4136-4278): plain data only
    def __getstate__(self):
        return self.__dict__.copy()

    # ------------------------------------------------------------------ energies
    def _getSavedEnergy(self, n, l, j, s=0.5):
        """alkali_atom_functions.py:880-892."""
        if abs(j - (l - 0.5)) < 0.001:
            return self._saved.get((n, l), 0.0)
        elif abs(j - (l + 0.5)) < 0.001:
            return self._saved.get((l, n), 0.0)
        raise ValueError("j (=%.1f) is not equal to l+1/2 nor l-1/2 (l=%d)" % (j, l))

    def getQuantumDefect(self, n, l, j, s=0.5):
        """alkali_atom_functions.py:894-976."""
        defect = 0.0
        if l < 5:
            c = self.quantumDefect[round(floor(s) + s + j - l)][l]
            if l < 3 and abs(c[0]) < 1e-9 and self.Z != 1:
                raise ValueError("Quantum defects for requested state "
                                 "(n = %d, l = %d, j = %.1f, s=%.1f) are uknown. "
                                 "Aborting calculation." % (n, l, j, s))
            defect = (c[0] + c[1] / ((n - c[0]) ** 2) + c[2] / ((n - c[0]) ** 4)
                      + c[3] / ((n - c[0]) ** 6) + c[4] / ((n - c[0]) ** 8)
                      + c[5] / ((n - c[0]) ** 10))
        else:
            n = int(n)
            l = int(l)
            top_r4 = 3 * pow(n, 2) - (l * (l + 1))
            bottom_r4 = 2 * pow(n, 5) * (l + 1.5) * (l + 1) * (l + 0.5) * l * (l - 0.5)
            r4 = top_r4 / bottom_r4
            top_r6 = (35 * pow(n, 4) - 5 * pow(n, 2) * (6 * l * (l + 1) - 5)
                      + 3 * (l + 2) * (l + 1) * l * (l - 1))
            bottom_r6 = (8 * pow(n, 7) * (l + 2.5) * (l + 2) * (l + 1.5) * (l + 1)
                         * (l + 0.5) * l * (l - 0.5) * (l - 1) * (l - 1.5))
            r6 = top_r6 / bottom_r6
            defect = 0.5 * (self.alpha_d_eff * r4 + self.alpha_q_eff * r6) * pow(n, 3)
        return defect

    def _getHydrogenicCorrection(self, n, l, j, s=0.5):
        """alkali_atom_functions.py:848-878."""
        C_alpha = physical_constants["fine-structure constant"][0]
        spinOrbit = (-self.scaledRydbergConstant * pow(C_alpha, 2)
                     * (j * (j + 1) - l * (l + 1) - s * (s + 1))
                     / (2 * l * (l + 0.5) * (l + 1) * pow(n, 3)))
        relCorr = (self.scaledRydbergConstant / pow(n, 2)
                   * (pow(C_alpha / n, 2) * ((n / (l + 0.5)) - 0.75)))
        return spinOrbit + relCorr

    def getEnergy(self, n, l, j, s=0.5):
        """Level energy relative to the ionisation limit in eV
        (alkali_atom_functions.py:794-846)."""
        key = (n, l, j, s)
        if key in self._energy:
            return self._energy[key]
        if l >= n:
            raise ValueError("Requested energy for state l=%d >= n=%d !" % (l, n))
        if ((not self.preferQuantumDefects or n < self.minQuantumDefectN)
                and (n <= self.NISTdataLevels)
                and (abs(self._getSavedEnergy(n, l, j, s=s)) > 1e-8)):
            e = self._getSavedEnergy(n, l, j, s=s)
        else:
            defect = self.getQuantumDefect(n, l, j, s=s)
            if l <= 4:
                e = -self.scaledRydbergConstant / ((n - defect) ** 2)
            else:
                e = (-self.scaledRydbergConstant / ((n - defect) ** 2)
                     - self._getHydrogenicCorrection(n, l, j, s=s))
        self._energy[key] = e
        return e

    def getZeemanEnergyShift(self, l, j, mj, magneticFieldBz, s=0.5):
        """Linear Zeeman shift in J (alkali_atom_functions.py:2647-2681)."""
        prefactor = physical_constants["Bohr magneton"][0] * magneticFieldBz
        gs = -physical_constants["electron g factor"][0]
        sumOverMl = 0
        for ml in np.linspace(mj - s, mj + s, round(2 * s + 1)):
            if abs(ml) <= l + 0.1:
                ms = mj - ml
                sumOverMl += (ml + gs * ms) * abs(CG(l, ml, s, ms, j, mj)) ** 2
        return prefactor * sumOverMl

    # --------------------------------------------------------- radial wavefunctions
    def _state_record(self, n, l, j, step=0.001):
        """arcb200_state_t with the call-site arguments of
        alkali_atom_functions.py:1068-1085 and 555-601 (s = 0.5 as there)."""
        inner = max(4.0 * step, self.alphaC ** (1 / 3.0))
        outer = 2.0 * n * (n + 15.0)
        E = self.getEnergy(n, l, j) / _HARTREE_EV
        mu = (self.mass - C_m_e) / self.mass
        if l < 4:
            p = (self.a1[l], self.a2[l], self.a3[l], self.a4[l], self.rc[l])
        else:
            p = (0.0, 0.0, 0.0, 0.0, 0.0)
        return _radial.make_state(inner, outer, step, 0.01, 0.01, l, 0.5, j, E,
                                  self.alphaC, self.alpha, self.Z, p[0], p[1], p[2],
                                  p[3], p[4], mu)

    # model potential of the valence electron (alkali_atom_functions.py:431-501); numpy
    # expressions, so they accept scalars and grids alike
    def effectiveCharge(self, l, r):
        return (1.0 + (self.Z - 1) * np.exp(-self.a1[l] * r)
                - r * (self.a3[l] + self.a4[l] * r) * np.exp(-self.a2[l] * r))

    def corePotential(self, l, r):
        return -self.effectiveCharge(l, r) / r - self.alphaC / (2 * r ** 4) * (
            1 - np.exp(-((r / self.rc[l]) ** 6)))

    def potential(self, l, s, j, r):
        so = self.alpha ** 2 / (2.0 * r ** 3) * (j * (j + 1.0) - l * (l + 1.0) - s * (s + 1)) / 2.0
        if l < 4:
            return self.corePotential(l, r) + so
        return -1.0 / r + so

    def radialWavefunction(self, l, s, j, stateEnergy, innerLimit, outerLimit, step):
        """(r, R(r)*r normalised) -- alkali_atom_functions.py:503-625, computed by the
        batched CUDA Numerov kernel (batch of one).  With cpp_numerov=False the reference's
        pure-Python route (aaf:606-623: NumerovBack on a callable k(x)) is taken instead: k(x) is
        evaluated on the host grid, the recurrence runs in numerov_coeff_kernel."""
        innerLimit = max(4.0 * step, innerLimit)
        mu = (self.mass - C_m_e) / self.mass
        if not self.cpp_numerov:
            from .numerov import NumerovBack

            def potential(x):
                r = x * x
                return -3.0 / (4.0 * r) + 4.0 * r * (
                    2.0 * mu * (stateEnergy - self.potential(l, s, j, r)) - l * (l + 1) / (r ** 2))
            r, psi_r = NumerovBack(innerLimit, outerLimit, potential, step, 0.01, 0.01)
            self.gpu_kernel_launches += 1
            suma = np.trapezoid(psi_r ** 2, x=r)
            return r, psi_r / np.sqrt(suma)
        if l < 4:
            p = (self.a1[l], self.a2[l], self.a3[l], self.a4[l], self.rc[l])
        else:
            p = (0.0, 0.0, 0.0, 0.0, 0.0)
        rec = _radial.make_state(innerLimit, outerLimit, step, 0.01, 0.01, l, s, j,
                                 stateEnergy, self.alphaC, self.alpha, self.Z, *p, mu)
        rb = _radial.RadialBatch(np.array([rec]), normalise=True, device=self.device)
        self.gpu_kernel_launches += rb.kernel_launches
        rb.raise_on_failure()
        return rb.wavefunction(0)

    # --------------------------------------------------------- radial matrix elements
    @staticmethod
    def _dipole_allowed(l1, j1, l2, j2):
        return abs(l1 - l2) == 1 and abs(j1 - j2) < 1.1       # aaf:1016-1019

    @staticmethod
    def _quadrupole_allowed(l1, j1, l2, j2):
        dl = abs(l1 - l2)
        return dl in (0, 1, 2) and abs(j1 - j2) < 2.1           # aaf:1134-1137

    def _ordered_key(self, n1, l1, j1, n2, l2, j2, s=0.5):
        """lower-energy state first (aaf:1021-1037)."""
        if self.getEnergy(n1, l1, j1, s=s) > self.getEnergy(n2, l2, j2, s=s):
            n1, l1, j1, n2, l2, j2 = n2, l2, j2, n1, l1, j1
        return (round(n1), round(l1), round(2 * j1), round(n2), round(l2), round(2 * j2))

    # ---- vector forms: whole tables at numpy speed (defineBasis of a 2000-state Stark basis asks
    # for 1e5 elements; one Python call chain per element was most of its wall time)
    @staticmethod
    def _as_pairs(pairs):
        """(n1, l1, j1, n2, l2, j2) rows as a contiguous float64 [m, 6] array (lists of tuples accepted)."""
        if not isinstance(pairs, np.ndarray):
            pairs = list(pairs)
            if not pairs:
                return np.zeros((0, 6))
        return np.ascontiguousarray(pairs, dtype=np.float64).reshape(-1, 6)

    @staticmethod
    def _index_states(P):
        """Rows (n1,l1,j1,n2,l2,j2) -> (S, ia, ib): the distinct (n, l, j) states as S[ns, 3] and,
        per row, the indices of its two states."""
        A = P.reshape(-1, 3)                              # rows alternate state 1, state 2
        if not len(A):
            return np.zeros((0, 3)), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
        n, l, j2 = (np.rint(c).astype(np.int64) for c in (A[:, 0], A[:, 1], 2 * A[:, 2]))
        code = ((n - n.min()) * (l.max() + 1) + l) * (j2.max() + 1) + j2
        _uc, first, inv = _unique_bounded(code)
        return A[first], inv[0::2], inv[1::2]

    @staticmethod
    def _allowed_cols(l1, j1, l2, j2, power):
        """Selection rules of getRadialMatrixElement (aaf:1016-1019) / getQuadrupoleMatrixElement
        (aaf:1134-1137) on columns."""
        dl, dj = np.abs(l1 - l2), np.abs(j1 - j2)
        if power == 1:
            return (dl == 1) & (dj < 1.1)
        return ((dl == 0) | (dl == 1) | (dl == 2)) & (dj < 2.1)

    def _keys_from(self, S, lo, hi, s):
        """Memo keys (n, l, 2j, n', l', 2j') of the elements S[lo] -> S[hi], as `_ordered_key` builds them."""
        C = np.empty((len(S), 3), dtype=np.int64)
        C[:, 0], C[:, 1], C[:, 2] = np.rint(S[:, 0]), np.rint(S[:, 1]), np.rint(2 * S[:, 2])
        return list(map(tuple, np.hstack([C[lo], C[hi]]).tolist()))

    def _unique_elements(self, S, ia, ib, power, s):
        """Distinct allowed elements among S[ia] <-> S[ib]: (ok mask over the rows, memo keys of the
        distinct elements with the lower-energy state first (aaf:1021-1037), index of every allowed
        row into them)."""
        ok = self._allowed_cols(S[ia, 1], S[ia, 2], S[ib, 1], S[ib, 2], power)
        if not ok.any():
            return ok, [], np.zeros(0, dtype=np.int64)
        ia, ib = ia[ok], ib[ok]
        es = 0.5 if s is None else s
        E = np.full(len(S), np.nan)
        for q in np.union1d(ia, ib):       # like the per-element path, only states of allowed pairs are looked up
            E[q] = self.getEnergy(int(round(S[q, 0])), int(round(S[q, 1])), float(S[q, 2]), s=es)
        swap = E[ia] > E[ib]
        lo, hi = np.where(swap, ib, ia), np.where(swap, ia, ib)
        ucode, _first, inv = _unique_bounded(lo * len(S) + hi)
        return ok, self._keys_from(S, ucode // len(S), ucode % len(S), s), inv

    def _compute_elements(self, keys, powers, s=0.5):
        """Values of the elements `keys` (memo keys, lower-energy state first; powers[i] = 1 dipole,
        2 quadrupole): ONE batched Numerov run over the distinct states, one integral launch."""
        K = np.array(keys, dtype=np.int64).reshape(-1, 6)
        S, ia, ib = self._index_states(K.astype(np.float64) * np.array([1, 1, .5, 1, 1, .5]))
        states = [self._state_record(int(t[0]), int(t[1]), float(t[2])) for t in S]
        pair_rows = np.empty((len(keys), 3), dtype=np.int32)
        pair_rows[:, 0], pair_rows[:, 1], pair_rows[:, 2] = ia, ib, powers
        rb = _radial.RadialBatch(np.array(states), normalise=True, device=self.device)
        rb.raise_on_failure()          # a state without a usable wavefunction: error, not a silent bad table
        if getattr(self, "radial_group", None) is not None and len(keys) >= 1024:
            # collective: every rank of the group builds the same table
            vals = rb.integrals_sharded(pair_rows, self.radial_group).cpu().numpy()
        elif not self.fast_radial or len(keys) < self.gram_threshold:
            # sorted by (power, lower-energy state): the warps of a CTA share psi_a / r_a through L1
            order = np.lexsort((pair_rows[:, 1], pair_rows[:, 0], pair_rows[:, 2]))
            vals = np.empty(len(keys))
            vals[order] = rb.integrals(pair_rows[order])
        else:
            # tensor-core path: one Gram block per ((l,j) of state a, (l,j) of state b, power)
            groups = {}
            for t, (key, power) in enumerate(zip(keys, powers)):
                groups.setdefault((key[1], key[2], key[4], key[5], power), []).append(t)
            blocks, where = [], [None] * len(keys)
            for g, (gk, ts) in enumerate(groups.items()):
                ra = sorted(set(int(pair_rows[t, 0]) for t in ts))
                cb = sorted(set(int(pair_rows[t, 1]) for t in ts))
                ia_ = {s_: q for q, s_ in enumerate(ra)}
                ib_ = {s_: q for q, s_ in enumerate(cb)}
                blocks.append((ra, cb, gk[4]))
                for t in ts:
                    where[t] = (g, ia_[int(pair_rows[t, 0])] * len(cb) + ib_[int(pair_rows[t, 1])])
            out, offs, _nref = rb.gram_refined_device(blocks, thresh=_radial.GRAM_REFINE_THRESH)
            out = out.cpu().numpy()
            vals = np.array([out[offs[g] + e] for g, e in where])
        self.gpu_kernel_launches += rb.kernel_launches
        return np.asarray(vals, dtype=np.float64).tolist()

    def radial_elements_indexed(self, S, ia, ib, power=1, s=0.5, useLiterature=True):
        """Vector form of getRadialMatrixElement (power 1) / getQuadrupoleMatrixElement (power 2)
        for the state pairs S[ia[k]] <-> S[ib[k]] (S = [ns, 3] rows (n, l, j)): one value per k,
        0 where the selection rules forbid the element.  Elements missing from the memo are
        computed in ONE GPU batch."""
        S = np.asarray(S, dtype=np.float64).reshape(-1, 3)
        ia, ib = np.asarray(ia, dtype=np.int64), np.asarray(ib, dtype=np.int64)
        out = np.zeros(len(ia))
        ok, keys, inv = self._unique_elements(S, ia, ib, power, s)
        if not keys:
            return out
        memo = self._dipole if power == 1 else self._quadrupole
        lit = self._lit if (power == 1 and useLiterature) else {}
        missing = [k for k in dict.fromkeys(keys) if k not in memo and k not in lit]   # S may repeat a state
        if missing:
            memo.update(zip(missing, self._compute_elements(missing, [power] * len(missing), s)))
        if lit:
            vals = np.array([lit[k][0] if k in lit else memo[k] for k in keys])
        else:
            vals = np.array(list(map(memo.__getitem__, keys)))
        out[ok] = vals[inv]
        return out

    def radial_elements(self, pairs, power=1, s=0.5, useLiterature=True):
        """radial_elements_indexed for explicit (n1, l1, j1, n2, l2, j2) rows."""
        S, ia, ib = self._index_states(self._as_pairs(pairs))
        return self.radial_elements_indexed(S, ia, ib, power, s=s, useLiterature=useLiterature)

    def precompute_radial(self, dipole_pairs=(), quadrupole_pairs=(), s=0.5,
                          useLiterature=True):
        """Evaluate all listed (n1,l1,j1,n2,l2,j2) elements (lists of tuples or [m, 6] arrays) in
        one GPU batch and memoise them.  This is the batch seam that replaces the reference's
        sqlite memo (seam B2).  Returns the number of newly computed elements."""
        fresh = []
        for pairs, memo, power in ((dipole_pairs, self._dipole, 1), (quadrupole_pairs, self._quadrupole, 2)):
            P = self._as_pairs(pairs)
            if not len(P):
                fresh.append([])
                continue
            _ok, keys, _inv = self._unique_elements(*self._index_states(P), power, s)
            lit = self._lit if (power == 1 and useLiterature) else {}
            fresh.append([k for k in dict.fromkeys(keys) if k not in memo and k not in lit])
        nd, nq = len(fresh[0]), len(fresh[1])
        if nd + nq == 0:
            return 0
        vals = self._compute_elements(fresh[0] + fresh[1], [1] * nd + [2] * nq, s)
        self._dipole.update(zip(fresh[0], vals[:nd]))
        self._quadrupole.update(zip(fresh[1], vals[nd:]))
        return nd + nq

    def getRadialMatrixElement(self, n1, l1, j1, n2, l2, j2, s=0.5, useLiterature=True):
        """Radial dipole element in a0*e (alkali_atom_functions.py:978-1105)."""
        if not self._dipole_allowed(l1, j1, l2, j2):
            return 0
        key = self._ordered_key(n1, l1, j1, n2, l2, j2, s=s)
        if useLiterature and key in self._lit:
            return self._lit[key][0]
        if key not in self._dipole:
            self.precompute_radial(dipole_pairs=[(n1, l1, j1, n2, l2, j2)],
                                   useLiterature=False)
        return self._dipole[key]

    def getQuadrupoleMatrixElement(self, n1, l1, j1, n2, l2, j2, s=0.5):
        """Radial quadrupole element in a0^2*e (alkali_atom_functions.py:1107-1210)."""
        if not self._quadrupole_allowed(l1, j1, l2, j2):
            return 0
        key = self._ordered_key(n1, l1, j1, n2, l2, j2, s=s)
        if key not in self._quadrupole:
            self.precompute_radial(quadrupole_pairs=[(n1, l1, j1, n2, l2, j2)])
        return self._quadrupole[key]

    def getRadialCoupling(self, n, l, j, n1, l1, j1, s=0.5):
        """alkali_atom_functions.py:2260-2298."""
        dl = abs(l - l1)
        if dl == 1 and abs(j - j1) < 1.1:
            return self.getRadialMatrixElement(n, l, j, n1, l1, j1, s=s)
        elif (dl == 0 or dl == 1 or dl == 2) and (abs(j - j1) < 2.1):
            return self.getQuadrupoleMatrixElement(n, l, j, n1, l1, j1, s=s)
        return 0

    def updateDipoleMatrixElementsFile(self):
        """No-op: the memo lives in memory (the reference dumps sqlite -> .npy here,
        alkali_atom_functions.py:1860-1906); see saveMatrixElementTables for the same files."""
        return None

    # -------------------------------------------------- persistence in the reference's formats
    def _memo_rows(self, memo):
        """Rows of the reference's memo tables: alkali (n1,l1,j1_x2,n2,l2,j2_x2,value)
        (alkali_atom_functions.py:275-283); divalent (n1,l1,j1,n2,l2,j2,s,value)
        (divalent_atom_functions.py:164-171).  Lower-energy state first in both."""
        return [tuple(float(x) for x in k) + (float(v),) for k, v in sorted(memo.items())]

    def saveMatrixElementTables(self, dipole_path, quadrupole_path):
        """Write the radial dipole / quadrupole memo as the `.npy` files stock ARC reads at start
        (`*_dipole_matrix_elements.npy`, alkali_atom_functions.py:246-265, written by its
        updateDipoleMatrixElementsFile, 1860-1906): float64[n, 7] (alkali) or [n, 8] (divalent)."""
        width = 8 if getattr(self, "divalent", False) else 7
        for path, memo in ((dipole_path, self._dipole), (quadrupole_path, self._quadrupole)):
            rows = self._memo_rows(memo)
            np.save(path, np.array(rows, dtype=np.float64).reshape(-1, width))

    def loadMatrixElementTables(self, dipole_path=None, quadrupole_path=None):
        """Read `.npy` tables in the reference's format into the memo (elements present are
        never recomputed, like rows of the reference's sqlite memo).  Returns rows loaded."""
        n = 0
        for path, memo in ((dipole_path, self._dipole), (quadrupole_path, self._quadrupole)):
            if path is None:
                continue
            data = np.load(path, allow_pickle=True, encoding="latin1")
            for row in data:
                memo[tuple(int(round(float(x))) for x in row[:-1])] = float(row[-1])
                n += 1
        return n

    def exportToArcDatabase(self, db_path):
        """Bulk-insert the memo into a SQLite file with the schema of the reference's
        `precalculated` database (tables `dipoleME`, `quadrupoleME`; SURVEY.md 8b seam B2):
        stock ARC pointed at it (`atom.conn`) answers these elements without running Numerov
        (alkali_atom_functions.py:1057-1065).  Returns (dipole rows, quadrupole rows)."""
        import sqlite3
        div = getattr(self, "divalent", False)
        cols = ("n1 TINYINT UNSIGNED, l1 TINYINT UNSIGNED, j1 TINYINT UNSIGNED, n2 TINYINT UNSIGNED, "
                "l2 TINYINT UNSIGNED, j2 TINYINT UNSIGNED, s TINYINT UNSIGNED" if div else
                "n1 TINYINT UNSIGNED, l1 TINYINT UNSIGNED, j1_x2 TINYINT UNSIGNED, n2 TINYINT UNSIGNED, "
                "l2 TINYINT UNSIGNED, j2_x2 TINYINT UNSIGNED")
        pk = "n1,l1,j1,n2,l2,j2,s" if div else "n1,l1,j1_x2,n2,l2,j2_x2"
        conn = sqlite3.connect(db_path)
        c = conn.cursor()
        counts = []
        for table, val, memo in (("dipoleME", "dme", self._dipole), ("quadrupoleME", "qme", self._quadrupole)):
            c.execute("CREATE TABLE IF NOT EXISTS %s (%s, %s DOUBLE, PRIMARY KEY (%s))" % (table, cols, val, pk))
            rows = [tuple(int(x) for x in k) + (float(v),) for k, v in memo.items()]
            c.executemany("INSERT OR REPLACE INTO %s VALUES (%s)" % (table, ",".join("?" * (len(pk.split(",")) + 1))), rows)
            counts.append(len(rows))
        conn.commit()
        conn.close()
        return tuple(counts)

    def importFromArcDatabase(self, db_path):
        """Read `dipoleME` / `quadrupoleME` of a reference database into the memo."""
        import sqlite3
        conn = sqlite3.connect(db_path)
        c = conn.cursor()
        n = 0
        for table, memo in (("dipoleME", self._dipole), ("quadrupoleME", self._quadrupole)):
            c.execute("SELECT COUNT(*) FROM sqlite_master WHERE type='table' AND name=?", (table,))
            if c.fetchone()[0] == 0:
                continue
            for row in c.execute("SELECT * FROM %s" % table):
                memo[tuple(int(x) for x in row[:-1])] = float(row[-1])
                n += 1
        conn.close()
        return n

    # -------------------------------------------------- angular parts of dipole elements
    def getReducedMatrixElementL(self, n1, l1, j1, n2, l2, j2, s=0.5):
        """<l||er||l'> (alkali_atom_functions.py:1277-1309)."""
        r = self.getRadialMatrixElement(n1, l1, j1, n2, l2, j2, s=s)
        return (-1) ** l1 * sqrt((2.0 * l1 + 1.0) * (2.0 * l2 + 1.0)) \
            * _w3j(l1, 1, l2, 0, 0, 0) * r

    def getReducedMatrixElementJ(self, n1, l1, j1, n2, l2, j2, s=0.5):
        """<j||er||j'> (alkali_atom_functions.py:1311-1345)."""
        from .wigner import Wigner6j
        return ((-1) ** (round(l1 + s + j2 + 1.0)) * sqrt((2.0 * j1 + 1.0) * (2.0 * j2 + 1.0))
                * Wigner6j(j1, 1.0, j2, l2, s, l1)
                * self.getReducedMatrixElementL(n1, l1, j1, n2, l2, j2, s=s))

    def getDipoleMatrixElement(self, n1, l1, j1, mj1, n2, l2, j2, mj2, q, s=0.5):
        """<n1 l1 j1 mj1| e r_q |n2 l2 j2 mj2> in a0*e (alkali_atom_functions.py:1347-1401, 2871)."""
        if abs(q) > 1.1:
            return 0
        return ((-1) ** (round(j1 - mj1)) * _w3j(j1, 1, j2, -mj1, -q, mj2)
                * self.getReducedMatrixElementJ(n1, l1, j1, n2, l2, j2, s=s))


def _w3j(*a):
    from .wigner import Wigner3j
    try:
        return Wigner3j(*a)
    except ValueError:
        return 0.0


class DivalentAtom(AlkaliAtom):
    """Singlet/triplet series of a divalent atom; radial elements are semiclassical
    (divalent_atom_functions.py:369-464, 713-783)."""

    divalent = True

    def __init__(self, className, preferQuantumDefects=True, device=None):
        super().__init__(className, preferQuantumDefects=preferQuantumDefects, device=device)
        a = self._raw
        self.defectFittingRange = a["defectFittingRange"]
        self._savedE = {(r[0], r[1], r[2], r[3]): r[4] for r in a.get("savedEnergy", [])}
        self._lit = {}
        for r in a.get("literatureDME", []):
            key = (r[0], r[1], r[2], r[3], r[4], r[5], r[6])
            if key not in self._lit or r[8] < self._lit[key][1]:
                self._lit[key] = (r[7], r[8])
        self.energyLevelsExtrapolated = False

    def _getSavedEnergy(self, n, l, j, s=0):
        return self._savedE.get((n, l, float(j), float(s)), 0)

    def getEnergy(self, n, l, j, s=None):
        """divalent_atom_functions.py:322-353."""
        if s is None:
            raise ValueError("Spin state for DivalentAtom has to be explicitly defined "
                             "as a keyword argument s=0 or s=1")
        key = (n, l, j, s)
        if key in self._energy:
            return self._energy[key]
        if l >= n:
            raise ValueError("Requested energy for state l=%d >= n=%d !" % (l, n))
        stateLabel = "%d%s%d" % (round(2 * s + 1), printStateLetter(l), j)
        minQuantumDefectN, maxQuantumDefectN = 100000, 0
        if stateLabel in self.defectFittingRange:
            minQuantumDefectN, maxQuantumDefectN = self.defectFittingRange[stateLabel]
        e = None
        if not self.preferQuantumDefects or n < minQuantumDefectN:
            savedEnergy = self._getSavedEnergy(n, l, j, s=s)
            if abs(savedEnergy) > 1e-8:
                e = savedEnergy
            elif n < minQuantumDefectN or n > maxQuantumDefectN:
                self.energyLevelsExtrapolated = True
        if e is None:
            defect = self.getQuantumDefect(n, l, j, s=s)
            e = -self.scaledRydbergConstant / ((n - defect) ** 2)
        self._energy[key] = e
        return e

    def radialWavefunction(self, *a, **k):
        raise NotImplementedError("radialWavefunction calculation for DivalentAtom "
                                  "has not been implemented yet.")   # div:785-794

    def _ordered(self, n1, l1, j1, n2, l2, j2, s):
        if self.getEnergy(n1, l1, j1, s=s) > self.getEnergy(n2, l2, j2, s=s):
            n1, l1, j1, n2, l2, j2 = n2, l2, j2, n1, l1, j1
        return (round(n1), round(l1), j1, round(n2), round(l2), j2, s)

    def _semiclassical_params(self, keys, kind, s):
        """(nu, nu', l, l') per memo key: effective quantum numbers from the energies for the dipole
        formula (aaf:2693-2699), from the quantum defects for the quadrupole one (aaf:2749-2750);
        every distinct state is evaluated once."""
        K = np.array([k[:6] for k in keys], dtype=np.float64).reshape(-1, 6)
        S, ia, ib = self._index_states(K)
        if kind == 1:
            e = np.array([self.getEnergy(int(t[0]), int(t[1]), float(t[2]), s=s) for t in S])
            nu = np.sqrt(-self.scaledRydbergConstant / e)
        else:
            nu = np.array([int(t[0]) - self.getQuantumDefect(int(t[0]), int(t[1]), float(t[2]), s=s) for t in S])
        return np.stack([nu[ia], nu[ib], K[:, 1], K[:, 4]], axis=1)

    @staticmethod
    def _allowed_cols(l1, j1, l2, j2, power):
        dl = np.abs(l1 - l2)
        if power == 1:                                    # div:404-407 (its dj test is j2 - j2 == 0)
            return dl == 1
        return ((dl == 0) | (dl == 1) | (dl == 2)) & (np.abs(j1 - j2) < 2.1)

    def _keys_from(self, S, lo, hi, s):
        """Memo keys (n, l, j, n', l', j', s), as `_ordered` builds them."""
        n, l = np.rint(S[:, 0]).astype(np.int64), np.rint(S[:, 1]).astype(np.int64)
        return list(zip(n[lo].tolist(), l[lo].tolist(), S[lo, 2].tolist(),
                        n[hi].tolist(), l[hi].tolist(), S[hi, 2].tolist(), [s] * len(lo)))

    def _compute_elements(self, keys, powers, s=None):
        """Semiclassical elements (aaf:2683-2802) of the memo keys, one launch per kind."""
        powers = np.asarray(powers)
        vals = np.zeros(len(keys))
        for kind in (1, 2):
            sel = np.flatnonzero(powers == kind)
            if len(sel):
                params = self._semiclassical_params([keys[i] for i in sel], kind, s)
                vals[sel] = _radial.semiclassical_table(params, kind, device=self.device)
                self.gpu_kernel_launches += 1
        return vals.tolist()

    def radial_elements_indexed(self, S, ia, ib, power=1, s=None, useLiterature=True):
        if s is None:
            raise ValueError("You must specify total angular momentum s explicitly.")
        return super().radial_elements_indexed(S, ia, ib, power, s=s, useLiterature=useLiterature)

    def radial_elements(self, pairs, power=1, s=None, useLiterature=True):
        if s is None:
            raise ValueError("You must specify total angular momentum s explicitly.")
        return super().radial_elements(pairs, power, s=s, useLiterature=useLiterature)

    def precompute_radial(self, dipole_pairs=(), quadrupole_pairs=(), s=None,
                          useLiterature=True):
        if s is None:
            raise ValueError("You must specify total angular momentum s explicitly.")
        return super().precompute_radial(dipole_pairs, quadrupole_pairs, s=s, useLiterature=useLiterature)

    def getRadialMatrixElement(self, n1, l1, j1, n2, l2, j2, s=None, useLiterature=True):
        """divalent_atom_functions.py:369-464."""
        if s is None:
            raise ValueError("You must specify total angular momentum s explicitly.")
        if abs(l1 - l2) != 1:
            return 0
        key = self._ordered(n1, l1, j1, n2, l2, j2, s)
        if useLiterature and key in self._lit:
            return self._lit[key][0]
        if key not in self._dipole:
            self.precompute_radial(dipole_pairs=[(n1, l1, j1, n2, l2, j2)], s=s,
                                   useLiterature=False)
        return self._dipole[key]

    def getQuadrupoleMatrixElement(self, n1, l1, j1, n2, l2, j2, s=None):
        """divalent_atom_functions.py:713-783."""
        if s is None:
            raise ValueError("You must specify total angular momentum s explicitly.")
        dl = abs(l1 - l2)
        if not (dl in (0, 1, 2) and abs(j1 - j2) < 2.1):
            return 0
        key = self._ordered(n1, l1, j1, n2, l2, j2, s)
        if key not in self._quadrupole:
            self.precompute_radial(quadrupole_pairs=[(n1, l1, j1, n2, l2, j2)], s=s)
        return self._quadrupole[key]


def _alkali(name):
    def ctor(preferQuantumDefects=True, cpp_numerov=True, device=None):
        return AlkaliAtom(name, preferQuantumDefects=preferQuantumDefects,
                          cpp_numerov=cpp_numerov, device=device)
    ctor.__name__ = name
    return ctor


def _divalent(name):
    def ctor(preferQuantumDefects=True, device=None):
        return DivalentAtom(name, preferQuantumDefects=preferQuantumDefects, device=device)
    ctor.__name__ = name
    return ctor


Hydrogen = _alkali("Hydrogen")
Lithium6 = _alkali("Lithium6")
Lithium7 = _alkali("Lithium7")
Sodium = _alkali("Sodium")
Potassium39 = _alkali("Potassium39")
Potassium = Potassium39
Potassium40 = _alkali("Potassium40")
Potassium41 = _alkali("Potassium41")
Rubidium85 = _alkali("Rubidium85")
Rubidium = Rubidium85
Rubidium87 = _alkali("Rubidium87")
Caesium = _alkali("Caesium")
Cesium = Caesium
Strontium88 = _divalent("Strontium88")
Calcium40 = _divalent("Calcium40")
Ytterbium174 = _divalent("Ytterbium174")
