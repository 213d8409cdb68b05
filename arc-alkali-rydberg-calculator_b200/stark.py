"""StarkMap -- drop-in for arc.calculations_atom_single.StarkMap on the hot path.

Same constructor, `defineBasis` and `diagonalise` signatures and result attributes
as the reference (calculations_atom_single.py:494-1025); the radial table comes from
the batched CUDA Numerov kernels and the per-field eigensolves from the batched CUDA
eigensolver (sweep.stark_sweep).  No CPU fallback.
"""
import numpy as np
from scipy.constants import physical_constants, h as C_h, e as C_e

from .wigner import CG
from .atoms import _unique_bounded
from . import sweep as _sweep


class CompositionView:
    """Lazy list-of-lists view of the top-k composition arrays.

    `composition[f][i]` -> list of [coefficient, basis index] pairs, the layout of
    the reference's `_stateComposition2` (calculations_atom_single.py:1395-1416),
    materialised on access so a 600 x 410 x 3 sweep does not cost 700k Python lists.
    """

    def __init__(self, comp_c, comp_i):
        self.comp_c, self.comp_i = comp_c, comp_i

    def __len__(self):
        return self.comp_c.shape[0]

    def __getitem__(self, f):
        if isinstance(f, slice):
            return [self[k] for k in range(*f.indices(len(self)))]
        c, i = self.comp_c[f], self.comp_i[f]
        return [[[c[a, b], int(i[a, b])] for b in range(c.shape[1]) if i[a, b] >= 0]
                for a in range(c.shape[0])]

    def __iter__(self):
        for f in range(len(self)):
            yield self[f]


def efield_angular(l1, j1, mj1, l2, j2, mj2, s=0.5):
    """sum_ml CG CG sqrt((l>^2-ml^2)/((2l>+1)(2l>-1)))  (PRA 20, 2251 eq. 10;
    alkali_atom_functions.py:4513-4553 `_EFieldCoupling.getAngular`)."""
    sumPart = 0.0
    for ml in np.linspace(mj1 - s, mj1 + s, round(2 * s + 1)):
        if (abs(ml) - 0.1 < l1) and (abs(ml) - 0.1 < l2):
            angularPart = 0.0
            if abs(l1 - l2 - 1) < 0.1:
                angularPart = ((l1 ** 2 - ml ** 2) / ((2.0 * l1 + 1.0) * (2.0 * l1 - 1.0))) ** 0.5
            elif abs(l1 - l2 + 1) < 0.1:
                angularPart = ((l2 ** 2 - ml ** 2) / ((2.0 * l2 + 1.0) * (2.0 * l2 - 1.0))) ** 0.5
            sumPart += (CG(l1, ml, s, mj1 - ml, j1, mj1) * CG(l2, ml, s, mj1 - ml, j2, mj2)
                        * angularPart)
    return sumPart


def stark_couplings(atom, states, mj, s):
    """Every |dl| = 1 pair of the basis `states` ([n, l, j, ...] rows) in vector form: index arrays
    (ii, jj) with ii < jj, the radial dipole elements (ONE GPU batch for those not memoised yet) and
    the angular factors `efield_angular(l_ii, j_ii, mj, l_jj, j_jj, mj)` -- evaluated once per distinct
    (l, j, l', j').  Shared by StarkMap.defineBasis and StarkBasisGenerator (Shirley)."""
    dimension = len(states)
    S = np.array([st[:3] for st in states], dtype=np.float64).reshape(dimension, 3)
    by_l = {}
    for idx, st in enumerate(states):
        by_l.setdefault(st[1], []).append(idx)
    pa, pb = [], []
    for l1, idxs in by_l.items():
        up = by_l.get(l1 + 1)
        if up:
            A, B = np.meshgrid(np.array(idxs), np.array(up), indexing="ij")
            pa.append(A.ravel())
            pb.append(B.ravel())
    if not pa:
        z = np.zeros(0, dtype=np.int64)
        return z, z, np.zeros(0), np.zeros(0)
    pa, pb = np.concatenate(pa), np.concatenate(pb)
    ii, jj = np.minimum(pa, pb), np.maximum(pa, pb)
    radial = atom.radial_elements_indexed(S, ii, jj, 1, s=s)
    lj, cls = np.unique(S[:, 1:3], axis=0, return_inverse=True)           # (l, j) classes of the states
    cls = cls.ravel()
    ukey, _first, inv = _unique_bounded(cls[ii] * len(lj) + cls[jj])
    ang = np.array([efield_angular(int(lj[k // len(lj), 0]), lj[k // len(lj), 1], mj,
                                   int(lj[k % len(lj), 0]), lj[k % len(lj), 1], mj, s=s)
                    for k in ukey])
    return ii, jj, radial, ang[inv]


class StarkMap:
    """Stark maps: basis |n l j mj>, H(F) = mat1 + F * mat2, eigensolve per field."""

    def __init__(self, atom):
        self.atom = atom
        self.basisStates = []
        self.mat1 = []
        self.mat2 = []
        self.indexOfCoupledState = 0
        self.eFieldList = []
        self.y = []
        self.highlight = []
        self.composition = []
        self.fig = 0
        self.ax = 0
        self.fitX, self.fitY, self.fittedCurveY = [], [], []
        self.drivingFromState = [0, 0, 0, 0, 0]
        self.maxCoupling = 0.0
        self.eFieldCouplingSaved = False
        self.s = 0.5
        self.gpu_kernel_launches = 0
        self.last_sweep = None

    def _eFieldCouplingDivE(self, n1, l1, j1, mj1, n2, l2, j2, mj2, s=0.5):
        """calculations_atom_single.py:650-666."""
        if (abs(mj1 - mj2) > 0.1) or (abs(l1 - l2) != 1):
            return 0
        result = (self.atom.getRadialMatrixElement(n1, l1, j1, n2, l2, j2, s=s)
                  * physical_constants["Bohr radius"][0] * C_e)
        return result * efield_angular(l1, j1, mj1, l2, j2, mj2, s=s)

    def defineBasis(self, n, l, j, mj, nMin, nMax, maxL, Bz=0, progressOutput=False,
                    debugOutput=False, s=0.5):
        """calculations_atom_single.py:674-844 -- same basis order, `mat1` (GHz) and
        `mat2` (GHz/(V/m)) dense float64; all radial elements in one GPU batch."""
        self.n, self.l, self.j, self.mj = n, l, j, mj
        self.nMin, self.nMax, self.maxL, self.Bz, self.s = nMin, nMax, maxL, Bz, s
        states = []
        for tn in range(nMin, nMax + 1):
            for tl in range(min(maxL + 1, tn)):
                for tj in np.linspace(tl - s, tl + s, round(2 * s + 1)):
                    if (abs(mj) - 0.1 <= tj) and (tn >= self.atom.groundStateN
                                                  or [tn, tl, tj] in self.atom.extraLevels):
                        states.append([tn, tl, tj, mj])
        dimension = len(states)
        if progressOutput:
            print("Found ", dimension, " states.")
        indexOfCoupledState = 0
        for index, st in enumerate(states):
            if (st[0] == n) and (abs(st[1] - l) < 0.1) and (abs(st[2] - j) < 0.1) \
                    and (abs(st[3] - mj) < 0.1):
                indexOfCoupledState = index
        self.basisStates = states
        self.indexOfCoupledState = indexOfCoupledState

        # --- batch seam: every |dl|=1 pair goes to the GPU radial table at once (index arrays
        # instead of one Python call chain per matrix element)
        before = self.atom.gpu_kernel_launches
        ii, jj, radial, ang = stark_couplings(self.atom, states, mj, s)
        self.gpu_kernel_launches += self.atom.gpu_kernel_launches - before

        mat1 = np.zeros((dimension, dimension), dtype=np.double)
        mat2 = np.zeros((dimension, dimension), dtype=np.double)
        zeeman = {}                                   # depends on (l, j, mj) only
        for i in range(dimension):
            kz = (states[i][1], states[i][2], states[i][3])
            if kz not in zeeman:
                zeeman[kz] = self.atom.getZeemanEnergyShift(kz[0], kz[1], kz[2], self.Bz, s=self.s) / C_h * 1.0e-9
            mat1[i][i] = (self.atom.getEnergy(states[i][0], states[i][1], states[i][2],
                                              s=self.s) * C_e / C_h * 1e-9 + zeeman[kz])
        coupling = (radial * physical_constants["Bohr radius"][0] * C_e * ang) * 1.0e-9 / C_h
        mat2[jj, ii] = coupling
        mat2[ii, jj] = coupling
        self.mat1, self.mat2 = mat1, mat2
        return 0

    def diagonalise(self, eFieldList, drivingFromState=[0, 0, 0, 0, 0], progressOutput=False,
                    debugOutput=False, upTo=4, totalContributionMax=0.95, batch=None, group=None):
        """calculations_atom_single.py:846-1025.  Results: `y` (list of ascending
        eigenvalue arrays, GHz), `highlight`, `composition` -- same indexing as the
        reference; eigenvectors never leave the GPU.  `group` (extension): None = all fields
        on this GPU; sweep.WORLD or a torch.distributed group = fields sharded over its ranks
        (collective call, every rank gets the full result)."""
        dimension = len(self.basisStates)
        self.maxCoupling = 0.0
        self.drivingFromState = drivingFromState
        drive_w = None
        if self.drivingFromState[0] != 0:
            n1, l1 = round(drivingFromState[0]), round(drivingFromState[1])
            j1, m1, q = drivingFromState[2], drivingFromState[3], drivingFromState[4]
            coupling = []
            for i in range(dimension):
                thisCoupling = 0.0
                b = self.basisStates[i]
                if (round(abs(b[1] - l1)) == 1) and (round(abs(b[2] - j1)) <= 1) \
                        and (round(abs(b[3] - m1 - q)) == 0):
                    thisCoupling += self.atom.getDipoleMatrixElement(
                        n1, l1, j1, m1, round(b[0]), round(b[1]), b[2], b[3], q, s=self.s)
                thisCoupling = abs(thisCoupling) ** 2
                self.maxCoupling = max(self.maxCoupling, thisCoupling)
                coupling.append(thisCoupling)
            if self.maxCoupling < 0.00000001:
                raise Exception(
                    "State that you specified in drivingFromState, for a given laser "
                    "polarization, is uncoupled from the specified Stark manifold.")
            drive_w = np.abs(np.array(coupling) / self.maxCoupling)   # single:1009-1011
        self.eFieldList = eFieldList
        fields = np.asarray(eFieldList, dtype=np.float64)
        topk = dimension if upTo == -1 else max(upTo - 1, 0)
        cmax = 2.0 if upTo == -1 else totalContributionMax
        res = _sweep.stark_sweep(np.diag(self.mat1).copy(), self.mat2, fields,
                                 self.indexOfCoupledState, drive_w=drive_w, topk=topk,
                                 contrib_max=cmax, device=getattr(self.atom, "device", None),
                                 batch=batch, group=group)
        self.last_sweep = res
        self.gpu_kernel_launches += res.kernel_launches
        self.y = [res.y[f] for f in range(len(fields))]
        self.highlight = [res.hl[f] for f in range(len(fields))]
        self.composition = CompositionView(res.comp_c, res.comp_i)
        return

    def getState(self, state, electricField, minN, maxN, maxL, accountForAmplitude=0.95,
                 debugOutput=False):
        """calculations_atom_single.py:1577-1668: basis states and coefficients of the eigenstate
        (at the given field) that carries most of `state`, sorted by declining contribution until
        `accountForAmplitude` of the norm is covered, and its energy in eV.  One GPU eigensolve with
        the full sorted composition (upTo = -1) instead of the host `eigh` + argsort; the overall
        sign of the coefficient vector is the eigensolver's (arbitrary in the reference as well)."""
        self.defineBasis(state[0], state[1], state[2], state[3], minN, maxN, maxL)
        N = len(self.basisStates)
        res = _sweep.stark_sweep(np.diag(self.mat1).copy(), self.mat2, np.array([float(electricField)]),
                                 self.indexOfCoupledState, topk=N, contrib_max=2.0,
                                 device=getattr(self.atom, "device", None))
        self.gpu_kernel_launches += res.kernel_launches
        # strongest contribution of the requested state: first maximum, like the strict `>` scan
        eigenvectorIndex = int(np.argmax(res.hl[0]))
        energy = res.y[0][eigenvectorIndex] * 1e9 * C_h / C_e
        if debugOutput:
            print("Max overlap = %.3f" % res.hl[0][eigenvectorIndex])
            print("Eigen energy (state index %d) = %.2f eV" % (eigenvectorIndex, energy))
        c, ix = res.comp_c[0][eigenvectorIndex], res.comp_i[0][eigenvectorIndex]
        i = 0
        coef, contributingStates = [], []
        while accountForAmplitude > 0 and i < N:
            coef.append(c[i])
            accountForAmplitude -= abs(coef[-1]) ** 2
            contributingStates.append(self.basisStates[int(ix[i])])
            i += 1
        return contributingStates, coef, energy

    def _addState(self, n1, l1, j1, mj1):
        """calculations_atom_single.py:1418-1430."""
        from .pairstate import printStateStringLatex
        if abs(self.s - 0.5) < 0.1:
            return "|%s m_j=%d/2\\rangle" % (printStateStringLatex(n1, l1, j1), round(2 * mj1))
        return "|%s m_j=%d\\rangle" % (printStateStringLatex(n1, l1, j1, s=self.s), round(mj1))

    def _stateComposition(self, stateVector):
        """LaTeX label of a [[coefficient, basis index], ...] list
        (calculations_atom_single.py:1376-1393)."""
        i = 0
        totalContribution = 0
        value = "$"
        while (i < len(stateVector)) and (totalContribution < 0.95):
            if i != 0 and stateVector[i][0] > 0:
                value += "+"
            value = value + ("%.2f" % stateVector[i][0]) + self._addState(*self.basisStates[stateVector[i][1]])
            totalContribution += abs(stateVector[i][0]) ** 2
            i += 1
        if totalContribution < 0.999:
            value += "+\\ldots"
        return value + "$"

    def getPolarizability(self, maxField=1.0e10, showPlot=False, debugOutput=False,
                          minStateContribution=0.0):
        """calculations_atom_single.py:1432-1575 (fit of the tracked level to
        offset - alpha F^2 / 2); returns MHz cm^2 / V^2."""
        from scipy.optimize import curve_fit
        if self.drivingFromState[0] != 0:
            raise Exception("Program can only find Polarizability of the original state "
                            "if you highlight original state.")
        eFieldList = np.asarray(self.eFieldList)
        hl = np.asarray(self.highlight)
        y = np.asarray(self.y)
        st = self.basisStates[self.indexOfCoupledState]
        e0 = self.atom.getEnergy(st[0], st[1], st[2], s=self.s) * C_e / C_h * 1e-9
        stop = 0
        while stop < len(eFieldList) - 1 and eFieldList[stop] < maxField:
            stop += 1
        xs, ys = [], []
        for ii in range(stop):
            jj = int(np.argmax(hl[ii]))
            if minStateContribution < hl[ii][jj]:
                xs.append(eFieldList[ii])
                ys.append(y[ii][jj] - e0)
        xs = np.array(xs) / 100.0
        ys = np.array(ys)

        def polarizabilityFit(eField, offset, alpha):
            return offset - 0.5 * alpha * eField ** 2
        try:
            popt, _ = curve_fit(polarizabilityFit, xs, ys, [0, 0])
        except Exception as ex:
            print(ex)
            return 0
        self.fitX, self.fitY = xs, ys
        self.fittedCurveY = np.array([polarizabilityFit(v, popt[0], popt[1]) for v in xs])
        return popt[1] * 1.0e3
