"""Host side of subsystem (1): batches radial states / pairs for the CUDA kernels.

`RadialBatch` packs a set of radial states into `arcb200_state_t` records, runs the
batched Numerov kernel (one warp per state) and keeps the wavefunctions RESIDENT in
HBM; radial dipole/quadrupole integrals for any list of state pairs are then one
more kernel launch (one warp per pair).  Replaces the per-element call chain
getRadialMatrixElement -> radialWavefunction x2 -> NumerovWavefunction -> trapezoid
(alkali_atom_functions.py:978-1105, 503-625) of the reference.
"""

import numpy as np

from . import _lib


def make_state(innerLimit, outerLimit, step, init1, init2, l, s, j, stateEnergy,
               alphaC, alpha, Z, a1, a2, a3, a4, rc, mu):
    """One arcb200_state_t record from the reference's 18 C-extension arguments
    (arc_c_extensions.c:113-116)."""
    rec = np.zeros((), dtype=_lib.STATE_DTYPE)
    rec["innerLimit"], rec["outerLimit"], rec["step"] = innerLimit, outerLimit, step
    rec["init1"], rec["init2"] = init1, init2
    rec["l"], rec["s"], rec["j"], rec["stateEnergy"] = int(l), s, j, stateEnergy
    rec["alphaC"], rec["alpha"], rec["Z"] = alphaC, alpha, int(Z)
    rec["a1"], rec["a2"], rec["a3"], rec["a4"], rec["rc"], rec["mu"] = a1, a2, a3, a4, rc, mu
    return rec


# Gram path: elements with |value| < GRAM_REFINE_THRESH * sqrt(<r^p>_a <r^p>_b) are recomputed by
# the exact per-pair kernel.  The tensor-core table differs from the per-pair formula by up to
# 3e-12 of that Cauchy-Schwarz scale (common r grid, summation order), i.e. 3e-9 relative at the
# threshold: inside the 1e-8 matrix-element tolerance for every element of the table.
GRAM_REFINE_THRESH = 1e-3


def sharded_table(compute, pairs, group="world"):
    """One matrix-element table for all ranks of the process group: rank g evaluates the
    contiguous shard pairs[g*P/G:(g+1)*P/G] with `compute(shard) -> tensor[len(shard)]` and
    the shards are all-gathered once (NCCL over NVLink on GPUs, gloo in the CPU tests), so the
    table is computed once per box instead of once per GPU.  Collective: every
    rank must call it with the same `pairs`."""
    from . import sweep
    rank, world = sweep.dist_info(group)
    if world == 1:
        return compute(pairs)
    lo, hi = sweep.shard_bounds(len(pairs), rank, world)
    local = compute(pairs[lo:hi])
    return sweep.gather_rows(local, len(pairs), group)


class RadialBatch:
    """Wavefunctions of `states` (structured array of arcb200_state_t) on the GPU."""

    def __init__(self, states, normalise=True, device=None):
        import torch
        lib = _lib.load()
        self.device = _lib.require_device(device)
        self.states = np.ascontiguousarray(states, dtype=_lib.STATE_DTYPE).reshape(-1)
        n = self.n = len(self.states)
        self.offsets = np.zeros(n + 1, dtype=np.int64)
        total = lib.arcb200_numerov_lengths(n, self.states.ctypes.data,
                                            self.offsets.ctypes.data)
        if total < 0:
            _lib.check(int(total), "arcb200_numerov_lengths")
        self.total = int(total)
        self.lengths = self.states["length"].copy()
        dev = torch.device("cuda", self.device)
        order = np.argsort(-self.lengths, kind="stable").astype(np.int32)
        self.d_states = torch.from_numpy(self.states.view(np.uint8).copy()).to(dev)
        self.d_offsets = torch.from_numpy(self.offsets).to(dev)
        self.d_order = torch.from_numpy(order).to(dev)
        self.d_psi = torch.empty(self.total + 64, dtype=torch.float64, device=dev)
        self.d_r = torch.empty(self.total + 64, dtype=torch.float64, device=dev)
        self.d_norm = torch.empty(n, dtype=torch.float64, device=dev)
        self.d_dp = torch.empty(n, dtype=torch.int32, device=dev)
        self.d_status = torch.zeros(n, dtype=torch.int32, device=dev)
        rc = lib.arcb200_numerov_batch(
            self.device, _lib.stream_ptr(self.device), n, _lib.ptr(self.d_states),
            _lib.ptr(self.d_offsets), _lib.ptr(self.d_order), _lib.ptr(self.d_psi),
            _lib.ptr(self.d_r), _lib.ptr(self.d_norm), _lib.ptr(self.d_dp),
            1 if normalise else 0, _lib.ptr(self.d_status))
        _lib.check(rc, "arcb200_numerov_batch")
        self.kernel_launches = 1
        self._status = None

    def status(self):
        """Per-state status of the integration (0 ok, 1 the reference's "Numerov error",
        2 no usable norm); one device -> host copy, cached."""
        if self._status is None:
            self._status = self.d_status.cpu().numpy()
        return self._status

    def raise_on_failure(self):
        """The reference's C extension returns NULL for a failed state (arc_c_extensions.c:200);
        the drop-in raises RuntimeError naming the states (SURVEY 8b seam B1)."""
        st = self.status()
        bad = np.nonzero(st)[0]
        if len(bad):
            s = self.states[bad[0]]
            raise RuntimeError("Numerov integration failed for %d state(s); first: index %d (l=%d, j=%.1f, "
                               "E=%.6g a.u.), status %d" % (len(bad), bad[0], s["l"], s["j"], s["stateEnergy"],
                                                            st[bad[0]]))

    def wavefunction(self, i):
        """(r, psi) of state i as numpy arrays (device -> host copy)."""
        o, L = int(self.offsets[i]), int(self.lengths[i])
        return (self.d_r[o:o + L].cpu().numpy(), self.d_psi[o:o + L].cpu().numpy())

    def divergence_points(self):
        return self.d_dp.cpu().numpy()

    def norms(self):
        return self.d_norm.cpu().numpy()

    def integrals_device(self, pairs):
        """pairs: int32[npairs,3] = (state a (lower energy), state b, power). Returns a
        CUDA float64 tensor of trapezoid(psi_a psi_b r_a^power, r_a)[0:min(La,Lb)]."""
        import torch
        lib = _lib.load()
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 3)
        dev = torch.device("cuda", self.device)
        d_pairs = torch.from_numpy(pairs).to(dev)
        out = torch.empty(len(pairs), dtype=torch.float64, device=dev)
        rc = lib.arcb200_radial_integrals(
            self.device, _lib.stream_ptr(self.device), len(pairs), _lib.ptr(d_pairs),
            _lib.ptr(self.d_states), _lib.ptr(self.d_offsets), _lib.ptr(self.d_psi),
            _lib.ptr(self.d_r), _lib.ptr(out))
        _lib.check(rc, "arcb200_radial_integrals")
        self.kernel_launches += 1
        return out

    def integrals(self, pairs):
        return self.integrals_device(pairs).cpu().numpy()

    def integrals_sharded(self, pairs, group="world"):
        """integrals_device with the pairs sharded over the ranks of `group` and the results
        all-gathered (every rank holds all wavefunctions: Numerov is ~10 % of a table)."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 3)
        return sharded_table(self.integrals_device, pairs, group)

    def gram_device(self, blocks):
        """blocks: list of (row_states, col_states, power).  Returns (out tensor, offsets):
        block g's dense |rows| x |cols| table starts at offsets[g] (row-major).  Tensor-core
        path (arcb200_radial_gram): every wavefunction is read once per block pair."""
        import torch
        lib = _lib.load()
        dev = torch.device("cuda", self.device)
        rows, cols, desc, off = [], [], np.zeros(len(blocks), dtype=_lib.GRAM_DTYPE), 0
        offs = []
        for g, (ra, cb, power) in enumerate(blocks):
            ra = np.asarray(ra, dtype=np.int32)
            cb = np.asarray(cb, dtype=np.int32)
            desc[g] = (len(rows), len(ra), len(cols), len(cb), int(power),
                       int(min(self.lengths[ra].max(), self.lengths[cb].max())), off)
            offs.append(off)
            rows.extend(ra.tolist())
            cols.extend(cb.tolist())
            off += len(ra) * len(cb)
        out = torch.zeros(max(off, 1), dtype=torch.float64, device=dev)
        if not blocks:
            return out, offs
        imax = int(np.argmax(self.lengths))
        Lg = int(self.lengths[imax])
        d_rgrid = self.d_r[int(self.offsets[imax]):]
        d_wgt = torch.empty(2 * Lg, dtype=torch.float64, device=dev)
        d_desc = torch.from_numpy(desc.view(np.uint8).copy()).to(dev)
        d_rows = torch.tensor(rows, dtype=torch.int32, device=dev)
        d_cols = torch.tensor(cols, dtype=torch.int32, device=dev)
        rc = lib.arcb200_radial_gram(
            self.device, _lib.stream_ptr(self.device), len(blocks), _lib.ptr(d_desc),
            _lib.ptr(d_rows), _lib.ptr(d_cols), _lib.ptr(self.d_states), _lib.ptr(self.d_offsets),
            _lib.ptr(self.d_psi), _lib.ptr(d_rgrid), Lg, _lib.ptr(d_wgt), _lib.ptr(out),
            int(desc["nrows"].max()), int(desc["ncols"].max()), int(desc["kmax"].max()))
        _lib.check(rc, "arcb200_radial_gram")
        self.kernel_launches += 3
        return out, offs


    def moments_device(self):
        """<r>, <r^2> of every state (trapezoid(psi^2 r^p, r)), CUDA tensor [n, 2]."""
        import torch
        if getattr(self, "_moments", None) is None:
            idx = np.arange(self.n, dtype=np.int32)
            pr = np.concatenate([np.stack([idx, idx, np.full_like(idx, 1)], 1),
                                 np.stack([idx, idx, np.full_like(idx, 2)], 1)])
            self._moments = self.integrals_device(pr).reshape(2, self.n).t().contiguous()
        return self._moments

    def gram_refined_device(self, blocks, thresh=GRAM_REFINE_THRESH):
        """Gram tables with per-element accuracy control.  The tensor-core sums carry an
        absolute rounding error of ~2e-14 * int|psi_a psi_b| r^p; for strongly cancelling
        elements (|value| < thresh * sqrt(<r^p>_a <r^p>_b), physically ~zero couplings) that
        is no longer 1e-8 relative, so exactly those elements are recomputed by the
        one-warp-per-pair kernel, whose summation order matches the reference's pairwise
        trapezoid sum.  Returns (out, offsets, number of refined elements)."""
        import torch
        out, offs = self.gram_device(blocks)
        if not blocks:
            return out, offs, 0
        dev = out.device
        mom = self.moments_device()
        rows_f, cols_f, pow_f = [], [], []
        for ra, cb, power in blocks:
            ra_t = torch.as_tensor(np.asarray(ra, dtype=np.int64), device=dev)
            cb_t = torch.as_tensor(np.asarray(cb, dtype=np.int64), device=dev)
            rows_f.append(ra_t.repeat_interleave(len(cb)))
            cols_f.append(cb_t.repeat(len(ra)))
            pow_f.append(torch.full((len(ra) * len(cb),), int(power), dtype=torch.int64, device=dev))
        rows_f, cols_f, pow_f = torch.cat(rows_f), torch.cat(cols_f), torch.cat(pow_f)
        n_el = rows_f.numel()
        scale = torch.sqrt(mom[rows_f, pow_f - 1].abs() * mom[cols_f, pow_f - 1].abs())
        flag = torch.nonzero(out[:n_el].abs() < thresh * scale).flatten()
        if flag.numel():
            # the per-pair kernel uses the r grid of its first state: put the lower-energy
            # (= smaller stateEnergy) state first like the reference does
            E = torch.as_tensor(self.states["stateEnergy"].copy(), device=dev)
            a, b = rows_f[flag], cols_f[flag]
            swap = E[a] > E[b]
            lo, hi = torch.where(swap, b, a), torch.where(swap, a, b)
            pairs = torch.stack([lo, hi, pow_f[flag]], 1).to(torch.int32).contiguous()
            lib = _lib.load()
            vals = torch.empty(flag.numel(), dtype=torch.float64, device=dev)
            rc = lib.arcb200_radial_integrals(
                self.device, _lib.stream_ptr(self.device), flag.numel(), _lib.ptr(pairs),
                _lib.ptr(self.d_states), _lib.ptr(self.d_offsets), _lib.ptr(self.d_psi),
                _lib.ptr(self.d_r), _lib.ptr(vals))
            _lib.check(rc, "arcb200_radial_integrals")
            self.kernel_launches += 1
            out[flag] = vals
        return out, offs, int(flag.numel())


_GL_CACHE = {}


def semiclassical_table(params, kind, device=None, nq=256):
    """params: float64[n,4] = (nu, nu1, l1, l2) -> semiclassical radial elements
    (alkali_atom_functions.py:2683-2802) evaluated on the GPU."""
    import torch
    lib = _lib.load()
    device = _lib.require_device(device)
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 4)
    if len(params) == 0:
        return np.zeros(0)
    dn = np.abs(params[:, 0] - params[:, 1]).max()
    while nq < 2.0 * dn + 64:
        nq *= 2
    if nq not in _GL_CACHE:
        x, w = np.polynomial.legendre.leggauss(nq)
        _GL_CACHE[nq] = np.concatenate([x, w])
    dev = torch.device("cuda", device)
    d_in = torch.from_numpy(params).to(dev)
    d_q = torch.from_numpy(_GL_CACHE[nq]).to(dev)
    out = torch.empty(len(params), dtype=torch.float64, device=dev)
    rc = lib.arcb200_semiclassical_table(device, _lib.stream_ptr(device), len(params),
                                         _lib.ptr(d_in), int(kind), _lib.ptr(d_q), nq,
                                         _lib.ptr(out))
    _lib.check(rc, "arcb200_semiclassical_table")
    return out.cpu().numpy()
