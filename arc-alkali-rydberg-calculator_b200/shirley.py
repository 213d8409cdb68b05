"""StarkBasisGenerator / ShirleyMethod -- drop-in for the Floquet Stark maps of
arc.calculations_atom_single (3172-3982).

Same constructor, `defineBasis`, `defineShirleyHamiltonian` and `diagonalise` signatures and
result attributes as the reference.  The basis and the bare matrices (`bareEnergies`, `V`) are
built on the host from the batched GPU radial table; the Floquet Hamiltonians
    Hf = H0 + dT*freq + B*eField,   dimension dim0 * (2 fn + 1)
are assembled and diagonalised on the GPU (arcb200_floquet_sweep: the same assembly kernel and
batched eigensolver as StarkMap, plus a kernel that folds the Floquet blocks into the
long-time-averaged transition probabilities), one batched sweep over the fields per frequency.
No CPU fallback.
"""
import warnings

import numpy as np
from scipy.constants import physical_constants, h as C_h, e as C_e

from . import sweep as _sweep
from .stark import efield_angular, stark_couplings  # noqa: F401


class StarkBasisGenerator:
    """Basis |n l j mj> and the bare Stark matrices (calculations_atom_single.py:3172-3668)."""

    def __init__(self, atom):
        self.atom = atom
        self.n = self.l = self.j = self.mj = self.s = None
        self.basisStates = []
        self.indexOfCoupledState = None
        self.targetState = []
        self.targetEnergy = None
        self.bareEnergies = []
        self.nMin = self.nMax = self.maxL = self.Bz = self.q = None
        self.H = []
        self.V = []
        self.eFieldCouplingSaved = False
        self.gpu_kernel_launches = 0

    def _eFieldCouplingDivE(self, n1, l1, j1, mj1, n2, l2, j2, mj2, s=0.5):
        """calculations_atom_single.py:3294-3311."""
        if (abs(mj1 - mj2) > 0.1) or (abs(l1 - l2) != 1):
            return 0
        result = (self.atom.getRadialMatrixElement(n1, l1, j1, n2, l2, j2, s=s)
                  * physical_constants["Bohr radius"][0] * C_e)
        return result * efield_angular(l1, j1, mj1, l2, j2, mj2, s=s)

    def _eFieldCoupling(self, n1, l1, j1, mj1, n2, l2, j2, mj2, eField, s=0.5):
        return self._eFieldCouplingDivE(n1, l1, j1, mj1, n2, l2, j2, mj2, s=s) * eField

    # ---- photon-order selection rules (calculations_atom_single.py:3318-3394), as one table:
    # a state |t> is reached from |s> with `order` photons of polarisation q when dl = |ls - lt|
    # equals `order` with dj following dl (stretched), or -- with no change of l -- j flips at fixed l
    # (order 1, j = l +- s on either side) / nothing changes but m_j (order 2)
    @staticmethod
    def _photonCoupling(order, ns, ls, js, mjs, nt, lt, jt, mjt, q, s=0.5):
        if (ns, ls, js, mjs) == (nt, lt, jt, mjt):
            return False
        if (mjt - mjs) / order != q:
            return False
        dl, stretched = ls - lt, (ls - lt == js - jt)
        if order == 1:
            return abs(dl) == 1 and (stretched or (js == jt and (js == ls + s or jt == lt + s)))
        return (abs(dl) == 2 and stretched) or (dl == 0 and js == jt)

    def _onePhotonCoupling(self, ns, ls, js, mjs, nt, lt, jt, mjt, q, s=0.5):
        return self._photonCoupling(1, ns, ls, js, mjs, nt, lt, jt, mjt, q, s)

    def _twoPhotonCoupling(self, ns, ls, js, mjs, nt, lt, jt, mjt, q, s=0.5):
        return self._photonCoupling(2, ns, ls, js, mjs, nt, lt, jt, mjt, q, s)

    def defineBasis(self, n, l, j, mj, q=0, nMin=None, nMax=None, maxL=None, Bz=0, edN=0,
                    basisStates=None, progressOutput=False, debugOutput=False, s=0.5):
        """Same call signature and attributes as calculations_atom_single.py:3396-3487: either an
        explicit list of basis states, or the window (nMin, nMax, maxL) filtered by photon order edN."""
        self.n, self.l, self.j, self.mj, self.Bz, self.s = n, l, j, mj, Bz, s
        target = [n, l, j, mj]
        if basisStates is not None:
            if target not in basisStates:
                raise ValueError(f"Target state {target} not in states list")
            self.basisStates = basisStates
            self.indexOfCoupledState = basisStates.index(target)
        else:
            if None in (nMin, nMax, maxL):
                raise ValueError("Input arguments are not complete. "
                                 + "Either specify nMin, nMax, maxL or basisStates")
            if edN not in (0, 1, 2):
                raise ValueError("edN must be 0, 1, or 2")
            self.q, self.edN = q, edN
            self.nMin, self.nMax, self.maxL = nMin, nMax, maxL
            self.basisStates = self._enumerateBasis()
            # the reference appends the target where it meets it in the (n, l, j) scan
            self.indexOfCoupledState = self.basisStates.index(target)
            if progressOutput:
                print("Found ", len(self.basisStates), " states.")
        self.targetState = self.basisStates[self.indexOfCoupledState]
        self._buildHamiltonian(progressOutput, debugOutput)

    def _admitted_mj(self, tn, tl, tj):
        """m_j with which (tn, tl, tj) enters the basis, or None (calculations_atom_single.py:3529-3574:
        the target itself; every state above the ground state for edN = 0; else the states one -- and,
        for edN = 2, two -- photons of polarisation q away from the target)."""
        n, l, j, mj, q, s = self.n, self.l, self.j, self.mj, self.q, self.s
        if (tn, tl, tj) == (n, l, j):
            return mj
        if self.edN == 0:
            # (`[tn, tl, tj] in extraLevels` is a list-in-list-of-tuples test in the reference: never true)
            return mj + q if (tn >= self.atom.groundStateN or [tn, tl, tj] in self.atom.extraLevels) else None
        if self._photonCoupling(1, n, l, j, mj, tn, tl, tj, mj + q, q, s):
            return mj + q
        if self.edN == 2 and self._photonCoupling(2, n, l, j, mj, tn, tl, tj, mj + 2 * q, q, s):
            return mj + 2 * q
        return None

    def _enumerateBasis(self):
        s = self.s
        scan = ((tn, tl, tj) for tn in range(self.nMin, self.nMax + 1) for tl in range(min(self.maxL + 1, tn))
                for tj in np.linspace(tl - s, tl + s, round(2 * s + 1)) if not abs(self.mj + self.q) - 0.1 > tj)
        return [[tn, tl, tj, m] for tn, tl, tj in scan for m in [self._admitted_mj(tn, tl, tj)] if m is not None]

    def _buildHamiltonian(self, progressOutput=False, debugOutput=False):
        """calculations_atom_single.py:3584-3668: `bareEnergies` (Hz) and `V` (Hz/(V/m),
        half the coupling; the angular factor is taken at `self.mj` for both states like the
        reference).  All radial elements come from one GPU batch."""
        states = self.basisStates
        dimension = len(states)
        before = self.atom.gpu_kernel_launches
        ii, jj, radial, ang = stark_couplings(self.atom, states, self.mj, self.s)
        self.gpu_kernel_launches += self.atom.gpu_kernel_launches - before

        self.bareEnergies = np.zeros((dimension), dtype=np.double)
        self.V = np.zeros((dimension, dimension), dtype=np.double)
        for i in range(dimension):
            self.bareEnergies[i] = (
                self.atom.getEnergy(states[i][0], states[i][1], states[i][2], s=self.s)
                * C_e / C_h
                + self.atom.getZeemanEnergyShift(states[i][1], states[i][2], states[i][3],
                                                 self.Bz, s=self.s) / C_h)
        coupling = 0.5 * (radial * physical_constants["Bohr radius"][0] * C_e * ang) / C_h
        self.V[jj, ii] = coupling
        self.V[ii, jj] = coupling
        self.H = np.diag(self.bareEnergies)
        self.targetEnergy = self.bareEnergies[self.indexOfCoupledState]


class ShirleyMethod(StarkBasisGenerator):
    """Shirley's time-independent Floquet Hamiltonian (calculations_atom_single.py:3671-3982)."""

    def __init__(self, atom):
        super().__init__(atom)
        self.fn = None
        self.H0 = self.B = self.dT = []
        self.eFields = self.freqs = None
        self.eigs, self.eigVectors, self.transProbs, self.targetShifts = [], [], [], []

    def defineShirleyHamiltonian(self, fn, debugOutput=False):
        """calculations_atom_single.py:3795-3852: `H0`, `dT`, `B` as scipy CSR like the
        reference (the GPU path reads their diagonals / dense form)."""
        import scipy.sparse as sp
        self.fn = fn
        if not fn >= 1:
            raise ValueError("Floquet expansion must be greater than 1."
                             + " Rank of 0 is equivalent to rotating wave approximation"
                             + " solution and is not covered by this method.")
        dimension = len(self.bareEnergies)
        self.H0 = sp.diags(np.tile(self.bareEnergies, 2 * fn + 1)).tocsr()
        self.dT = sp.block_diag([sp.diags([float(i)], 0, shape=(dimension, dimension))
                                 for i in range(-fn, fn + 1, 1)], dtype=np.double).tocsr()
        self.B = sp.bmat([[self.V if abs(i - j) == 1 else None for i in range(-fn, fn + 1, 1)]
                          for j in range(-fn, fn + 1, 1)], dtype=np.double).tocsr()

    def diagonalise(self, eFields, freqs, progressOutput=False, debugOutput=False,
                    keepEigenvectors=False, batch=None):
        """calculations_atom_single.py:3854-3982.  Results `eigs`, `transProbs`,
        `targetShifts` with the reference's shapes (outer product of the inputs, singleton
        dimensions squeezed).  `eigVectors` (dim x dim per point -- 107 MB per point at the
        documented 2583 basis) is only filled with `keepEigenvectors=True`; the transition
        probabilities and shifts are computed on the GPU without it."""
        dim0 = len(self.basisStates)
        dim1 = 2 * self.fn + 1
        N = dim0 * dim1
        refInd = self.fn * dim0
        tarInd = self.indexOfCoupledState + refInd
        self.eFields = np.array(eFields, ndmin=1)
        self.freqs = np.array(freqs, ndmin=1)
        fields = np.asarray(self.eFields, dtype=np.float64).reshape(-1)
        fr = np.asarray(self.freqs, dtype=np.float64).reshape(-1)
        eig = np.zeros((len(fields), len(fr), N), dtype=np.double)
        transProbs = np.zeros((len(fields), len(fr), dim0), dtype=np.double)
        targetShifts = np.zeros((len(fields), len(fr)), dtype=np.double)
        eigVec = np.zeros((len(fields), len(fr), N, N), dtype=np.complex128) if keepEigenvectors else []
        h0 = np.tile(np.asarray(self.bareEnergies, dtype=np.float64), dim1)
        kblock = np.repeat(np.arange(-self.fn, self.fn + 1, dtype=np.float64), dim0)
        Bd = self.B.toarray()
        import torch
        device = getattr(self.atom, "device", None)
        from . import _lib
        dev = torch.device("cuda", _lib.require_device(device))
        d_B = torch.as_tensor(np.ascontiguousarray(Bd)).to(dev)
        for fi, freq in enumerate(fr):
            hdiag = h0 + kblock * freq                        # H0 + dT * freq
            y, hl, tp, launches = _sweep.floquet_sweep(hdiag, None, fields, tarInd, dim0,
                                                       device=device, batch=batch, d_B=d_B)
            self.gpu_kernel_launches += launches
            eig[:, fi, :] = y
            transProbs[:, fi, :] = tp
            for ei in range(len(fields)):
                evInd = int(np.argmax(hl[ei]))
                if np.count_nonzero(y[ei] == y[ei][evInd]) > 1:
                    warnings.warn("Multiple states have same overlap with target. Only saving first one.")
                targetShifts[ei, fi] = self.targetEnergy - y[ei][evInd]
            if keepEigenvectors:
                Hs = np.stack([np.diag(hdiag) + Bd * f for f in fields])
                _w, v = _sweep.syevd_batched(Hs, device=device)
                eigVec[:, fi] = v
        shp = (*self.eFields.shape, *self.freqs.shape)
        self.eigs = eig.reshape(*shp, N).squeeze()
        self.transProbs = transProbs.reshape(*shp, dim0).squeeze()
        self.targetShifts = targetShifts.reshape(shp).squeeze()
        self.eigVectors = eigVec.reshape(*shp, N, N).squeeze() if keepEigenvectors else []
