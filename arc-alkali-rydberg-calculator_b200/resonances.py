"""StarkMapResonances -- drop-in for arc.calculations_atom_pairstate.StarkMapResonances
(4478-4820): near-resonant dipole-coupled pair states in an electric field.

Same constructor and `findResonances` signature and result attributes (`eFieldList`,
`originalStateY`, `originalStateContribution`, `y`, `composition`) as the reference.  The
single-atom Stark maps come from the GPU StarkMap; the O(N1 N2) pair-energy scan per field,
a pure-Python double loop in the reference (4753-4789), is one CUDA kernel
(arcb200_resonance_scan) whose output order equals the reference's loop order.  Plotting
(`showPlot`, the matplotlib figure the reference builds inside findResonances) is out of scope.
"""
import numpy as np
from scipy.constants import h as C_h, e as C_e

from . import _lib
from .stark import StarkMap


class StarkMapResonances:
    def __init__(self, atom1, state1, atom2, state2):
        """calculations_atom_pairstate.py:4539-4588."""
        self.atom1 = atom1
        if getattr(atom1, "divalent", False) and (len(state1) != 5 or (state1[4] != 0 and state1[4] != 1)):
            raise ValueError("For divalent atoms state specification has to include total spin "
                             "angular momentum s as the last number in the state specification "
                             "[n,l,j,m_j,s].")
        self.state1 = state1
        if len(self.state1) == 4:
            self.state1.append(0.5)
        self.atom2 = atom2
        if getattr(atom2, "divalent", False) and (len(state1) != 5 or (state1[4] != 0 and state1[4] != 1)):
            raise ValueError("For divalent atoms state specification has to include total spin "
                             "angular momentum s as the last numbre in the state specification "
                             "[n,l,j,m_j,s].")
        self.state2 = state2
        if len(self.state2) == 4:
            self.state2.append(0.5)
        self.pairStateEnergy = ((atom1.getEnergy(self.state1[0], self.state1[1], self.state1[2],
                                                 s=self.state1[4])
                                 + atom2.getEnergy(self.state2[0], self.state2[1], self.state2[2],
                                                   s=self.state2[4])) * C_e / C_h * 1e-9)
        self.eFieldList = []
        self.originalStateY, self.originalStateContribution = [], []
        self.r, self.y, self.composition = [], [], []
        self.gpu_kernel_launches = 0

    def _scan(self, sm1, sm2, eMin, eMax):
        """All (j, jj) of every field with eMin < y1 + y2 - E0 < eMax whose leading basis
        states are dipole coupled to the original states; reference loop order."""
        import torch
        lib = _lib.load()
        device = _lib.require_device(getattr(self.atom1, "device", None))
        dev = torch.device("cuda", device)
        y1, y2 = np.asarray(sm1.y), np.asarray(sm2.y)
        nF, N1, N2 = y1.shape[0], y1.shape[1], y2.shape[1]
        l1 = np.array([b[1] for b in sm1.basisStates])
        l2 = np.array([b[1] for b in sm2.basisStates])
        lead1 = sm1.composition.comp_i[:, :, 0]
        lead2 = sm2.composition.comp_i[:, :, 0]
        ok1 = (np.abs(l1[lead1] - self.state1[1]) == 1).astype(np.uint8)
        ok2 = (np.abs(l2[lead2] - self.state2[1]) == 1).astype(np.uint8)
        d_y1, d_y2 = torch.as_tensor(y1.copy()).to(dev), torch.as_tensor(y2.copy()).to(dev)
        d_ok1, d_ok2 = torch.as_tensor(ok1).to(dev), torch.as_tensor(ok2).to(dev)
        cnt = torch.zeros(nF * N1, dtype=torch.int32, device=dev)
        args = (device, _lib.stream_ptr(device), nF, N1, N2, _lib.ptr(d_y1), _lib.ptr(d_y2),
                _lib.ptr(d_ok1), _lib.ptr(d_ok2), float(self.pairStateEnergy), float(eMin), float(eMax))
        _lib.check(lib.arcb200_resonance_scan(*args, None, _lib.ptr(cnt), None, None, None), "resonance count")
        off = torch.cumsum(cnt.to(torch.int64), 0) - cnt.to(torch.int64)
        total = int(cnt.sum().item())
        oe = torch.empty(max(total, 1), dtype=torch.float64, device=dev)
        oj = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        ojj = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        _lib.check(lib.arcb200_resonance_scan(*args, _lib.ptr(off), _lib.ptr(cnt), _lib.ptr(oe),
                                              _lib.ptr(oj), _lib.ptr(ojj)), "resonance fill")
        self.gpu_kernel_launches += 2
        fo = torch.cat([off[::N1], torch.tensor([total], dtype=torch.int64, device=dev)]).cpu().numpy()
        return oe[:total].cpu().numpy(), oj[:total].cpu().numpy(), ojj[:total].cpu().numpy(), fo

    def findResonances(self, nMin, nMax, maxL, eFieldList, energyRange=[-5.0e9, +5.0e9], Bz=0,
                       progressOutput=False):
        """calculations_atom_pairstate.py:4590-4790 (without the figure)."""
        self.eFieldList = eFieldList
        self.Bz = Bz
        eMin = energyRange[0] * 1.0e-9
        eMax = energyRange[1] * 1.0e-9
        sm1 = StarkMap(self.atom1)
        sm1.defineBasis(self.state1[0], self.state1[1], self.state1[2], self.state1[3], nMin, nMax,
                        maxL, Bz=self.Bz, progressOutput=progressOutput, s=self.state1[4])
        sm1.diagonalise(eFieldList, progressOutput=progressOutput)
        if ((self.atom2 is self.atom1) and (self.state1[0] == self.state2[0])
                and (self.state1[1] == self.state2[1]) and (abs(self.state1[2] - self.state2[2]) < 0.1)
                and (abs(self.state1[3] - self.state2[3]) < 0.1)
                and (abs(self.state1[4] - self.state2[4]) < 0.1)):
            sm2 = sm1
        else:
            sm2 = StarkMap(self.atom2)
            sm2.defineBasis(self.state2[0], self.state2[1], self.state2[2], self.state2[3], nMin,
                            nMax, maxL, Bz=self.Bz, progressOutput=progressOutput, s=self.state2[4])
            sm2.diagonalise(eFieldList, progressOutput=progressOutput)
        self.gpu_kernel_launches += sm1.gpu_kernel_launches + (0 if sm2 is sm1 else sm2.gpu_kernel_launches)

        self.originalStateY, self.originalStateContribution = [], []
        for i in range(len(sm1.eFieldList)):
            jmax1 = int(np.argmax(sm1.highlight[i]))       # first maximum, like the reference's scan
            jmax2 = int(np.argmax(sm2.highlight[i]))
            self.originalStateY.append(sm1.y[i][jmax1] + sm2.y[i][jmax2] - self.pairStateEnergy)
            self.originalStateContribution.append((sm1.highlight[i][jmax1] + sm2.highlight[i][jmax2]) / 2.0)

        dmlist1 = [1, 0]
        if self.state1[3] != 0.5:
            dmlist1.append(-1)
        dmlist2 = [1, 0]
        if self.state2[3] != 0.5:
            dmlist2.append(-1)
        n1, l1, j1, mj1 = self.state1[0], self.state1[1] + 1, self.state1[2] + 1, self.state1[3]
        n2, l2, j2, mj2 = self.state2[0], self.state2[1] + 1, self.state2[2] + 1, self.state2[3]
        self.r, self.y, self.composition = [], [], []
        # NOTE (reference behaviour, kept): for identical states sm2 IS sm1, so the inner
        # defineBasis below re-defines the map used for atom 1 as well
        # (calculations_atom_pairstate.py:4661-4676, 4725-4751)
        base_launches = sm1.gpu_kernel_launches + (0 if sm2 is sm1 else sm2.gpu_kernel_launches)
        for dm1 in dmlist1:
            sm1.defineBasis(n1, l1, j1, mj1 + dm1, nMin, nMax, maxL, Bz=self.Bz,
                            progressOutput=progressOutput, s=self.state1[4])
            sm1.diagonalise(eFieldList, progressOutput=progressOutput)
            for dm2 in dmlist2:
                sm2.defineBasis(n2, l2, j2, mj2 + dm2, nMin, nMax, maxL, Bz=self.Bz,
                                progressOutput=progressOutput, s=self.state2[4])
                sm2.diagonalise(eFieldList, progressOutput=progressOutput)
                oe, oj, ojj, fo = self._scan(sm1, sm2, eMin, eMax)
                for i in range(len(sm1.eFieldList)):
                    lo, hi = int(fo[i]), int(fo[i + 1])
                    yList = list(oe[lo:hi])
                    c1, c2 = sm1.composition[i], sm2.composition[i]
                    compositionList = [[sm1._stateComposition(c1[int(a)]), sm2._stateComposition(c2[int(b)])]
                                       for a, b in zip(oj[lo:hi], ojj[lo:hi])]
                    if len(self.y) <= i:
                        self.y.append(yList)
                        self.composition.append(compositionList)
                    else:
                        self.y[i].extend(yList)
                        self.composition[i].extend(compositionList)
        self.gpu_kernel_launches += (sm1.gpu_kernel_launches
                                     + (0 if sm2 is sm1 else sm2.gpu_kernel_launches) - base_launches)
        for i in range(len(self.eFieldList)):
            self.y[i] = np.array(self.y[i])
            self.composition[i] = np.array(self.composition[i])
