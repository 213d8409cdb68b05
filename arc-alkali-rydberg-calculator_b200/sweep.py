"""Sweep driver: Hamiltonian assembly + batched eigensolve over field / distance points.

Host side of subsystems (2)+(3).  Sweep points are independent eigenproblems that
share H0 and V (or matDiagonal and matR), so they shard with NO data-path
collective: rank g of G takes the contiguous slice points[g*P/G:(g+1)*P/G]
(SURVEY.md 8e), solves it on its own GPU, and the per-rank results are
all-gathered once at the end (NCCL on GPUs; gloo in the CPU tests).

Sharding is OPT-IN: every entry point takes `group=None` (= this process solves all
points, no collective is ever entered).  Pass `group=WORLD` (the default process group)
or a `torch.distributed` ProcessGroup to shard; the call is then collective and every
rank of the group has to make it.
"""
import numpy as np

from . import _lib


def shard_bounds(npoints, rank, world):
    """Contiguous slice [lo, hi) of `npoints` sweep points owned by `rank`."""
    lo = (npoints * rank) // world
    hi = (npoints * (rank + 1)) // world
    return lo, hi


WORLD = "world"      # group=WORLD: shard over the default torch.distributed process group


def _pg(group):
    return None if (isinstance(group, str) and group == WORLD) else group


def dist_info(group=None):
    """(rank, world) of `group`; (0, 1) for group=None (no sharding requested).  A group was
    asked for but torch.distributed is not initialised -> error (a silent single-rank run of
    a call the user meant to be collective would solve everything on every rank)."""
    if group is None:
        return 0, 1
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("group=%r requested but torch.distributed is not initialised" % (group,))
    return dist.get_rank(_pg(group)), dist.get_world_size(_pg(group))


def gather_rows(local, npoints, group=WORLD):
    """All-gather per-rank row blocks (torch tensors, first dim = local points) into the
    full [npoints, ...] tensor on every rank.  Ragged shards are padded to the largest."""
    import torch
    import torch.distributed as dist
    rank, world = dist_info(group)
    if world == 1:
        return local
    sizes = [shard_bounds(npoints, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=_pg(group))
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)


def _pick_batch(N, npoints, nvec, device, batch=None):
    import os
    import torch
    if batch is None and os.environ.get("ARCB200_BATCH"):      # tuning knob
        batch = int(os.environ["ARCB200_BATCH"])
    if batch is not None:
        return max(1, min(int(batch), npoints))
    free, _total = torch.cuda.mem_get_info(device)
    # blocks cached by torch from earlier sweeps are reusable: count them as free ...
    cached = 0
    try:
        cached = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    except Exception:
        pass
    free_now = free
    free += cached
    lib = _lib.load()
    per = max(1, lib.arcb200_sweep_workspace_bytes(N, 1, nvec))
    # large matrices: the kernels that give a matrix one CTA (bulge chasing, panel QR, Q2) cost
    # the same for any batch below the SM count, so take as many as the memory holds
    frac = 0.8 if N > 4096 else 0.6
    b = int(min(npoints, max(1, (frac * free) // per)))
    # one CTA (or a small cluster) per matrix: 148 matrices in flight fill the GPU
    # (two per SM while two CTAs' shared memory fits, N <= ~4000)
    # small matrices: fewer, larger batches -- the sweep is launch-bound on the host (measured: C1 600 fields
    # 12.7 k solves/s in two batches, 13.9 k in one; C4 singlet 4.4 k at 286 per batch, 4.7 k at 667)
    cap = 1184 if N <= 512 else (740 if N <= 1024 else (296 if N <= 4000 else 148))
    b = max(1, min(b, cap))
    # equal chunks: a nearly empty tail pass costs a full sweep latency
    nchunks = -(-npoints // b)
    b = -(-npoints // nchunks)
    # ... and hand them back to the driver first when the workspace does not fit beside them
    if cached and per * b > 0.9 * free_now:
        torch.cuda.empty_cache()
    return b


class SweepResult:
    __slots__ = ("y", "hl", "comp_c", "comp_i", "kernel_launches", "h2d_bytes", "d2h_bytes",
                 "overlap")


def stark_sweep(h0diag, V, fields, idx0, drive_w=None, topk=3, contrib_max=0.95,
                device=None, batch=None, group=None, to_host=True):
    """Solve H = diag(h0) + F*V for every F (calculations_atom_single.py:976-1021).
    group=None: all fields on this GPU; group=WORLD / a ProcessGroup: collective, sharded."""
    import torch
    lib = _lib.load()
    device = _lib.require_device(device)
    dev = torch.device("cuda", device)
    fields = np.ascontiguousarray(fields, dtype=np.float64)
    nF_all = len(fields)
    rank, world = dist_info(group)
    lo, hi = shard_bounds(nF_all, rank, world)
    nF = hi - lo
    N = int(len(h0diag))
    res = SweepResult()
    d_h0 = torch.as_tensor(np.ascontiguousarray(h0diag, dtype=np.float64)).to(dev, non_blocking=True)
    d_V = torch.as_tensor(np.ascontiguousarray(V, dtype=np.float64)).to(dev, non_blocking=True)
    d_F = torch.as_tensor(fields[lo:hi].copy()).to(dev, non_blocking=True)
    d_w = None
    if drive_w is not None:
        d_w = torch.as_tensor(np.ascontiguousarray(drive_w, dtype=np.float64)).to(dev)
    res.h2d_bytes = (d_h0.numel() + d_V.numel() + d_F.numel()) * 8
    y = torch.empty((nF, N), dtype=torch.float64, device=dev)
    hl = torch.empty((nF, N), dtype=torch.float64, device=dev)
    comp_c = torch.zeros((nF, N, max(topk, 1)), dtype=torch.float64, device=dev)
    comp_i = torch.full((nF, N, max(topk, 1)), -1, dtype=torch.int32, device=dev)
    nl = _lib.LaunchCount()
    if nF > 0:
        b = _pick_batch(N, nF, N, device, batch)
        wbytes = int(lib.arcb200_sweep_workspace_bytes(N, b, N))
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        rc = lib.arcb200_stark_sweep(device, _lib.stream_ptr(device), N, _lib.ptr(d_h0),
                                     _lib.ptr(d_V), nF, _lib.ptr(d_F), int(idx0),
                                     _lib.ptr(d_w), int(topk), float(contrib_max),
                                     _lib.ptr(y), _lib.ptr(hl), _lib.ptr(comp_c),
                                     _lib.ptr(comp_i), _lib.ptr(work), wbytes, b, nl.addr)
        _lib.check(rc, "arcb200_stark_sweep")
        res.kernel_launches = nl.value
    else:
        res.kernel_launches = 0
    if world > 1:
        y, hl = gather_rows(y, nF_all, group), gather_rows(hl, nF_all, group)
        comp_c, comp_i = gather_rows(comp_c, nF_all, group), gather_rows(comp_i, nF_all, group)
    if to_host:
        res.y, res.hl = y.cpu().numpy(), hl.cpu().numpy()
        res.comp_c, res.comp_i = comp_c.cpu().numpy(), comp_i.cpu().numpy()
        res.d2h_bytes = res.y.nbytes + res.hl.nbytes + res.comp_c.nbytes + res.comp_i.nbytes
    else:
        res.y, res.hl, res.comp_c, res.comp_i = y, hl, comp_c, comp_i
        res.d2h_bytes = 0
    return res


def floquet_sweep(h0diag, B, fields, idx0, dim0, device=None, batch=None, d_B=None):
    """Solve H = diag(h0) + F*B for every F and fold the Floquet blocks into transition
    probabilities (calculations_atom_single.py:3931-3957).  Returns (y, hl, tprob) as numpy
    arrays [nF, N], [nF, N], [nF, dim0].  `d_B` may carry B already resident on the device
    (it does not change between drive frequencies)."""
    import torch
    lib = _lib.load()
    device = _lib.require_device(device)
    dev = torch.device("cuda", device)
    fields = np.ascontiguousarray(fields, dtype=np.float64).reshape(-1)
    nF, N = len(fields), int(len(h0diag))
    d_h0 = torch.as_tensor(np.ascontiguousarray(h0diag, dtype=np.float64)).to(dev)
    if d_B is None:
        d_B = torch.as_tensor(np.ascontiguousarray(B, dtype=np.float64)).to(dev)
    d_F = torch.as_tensor(fields).to(dev)
    y = torch.empty((nF, N), dtype=torch.float64, device=dev)
    hl = torch.empty((nF, N), dtype=torch.float64, device=dev)
    tp = torch.empty((nF, dim0), dtype=torch.float64, device=dev)
    b = _pick_batch(N, nF, N, device, batch)
    wbytes = int(lib.arcb200_sweep_workspace_bytes(N, b, N))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    nl = _lib.LaunchCount()
    rc = lib.arcb200_floquet_sweep(device, _lib.stream_ptr(device), N, int(dim0), _lib.ptr(d_h0),
                                   _lib.ptr(d_B), nF, _lib.ptr(d_F), int(idx0), _lib.ptr(y),
                                   _lib.ptr(hl), _lib.ptr(tp), _lib.ptr(work), wbytes, b, nl.addr)
    _lib.check(rc, "arcb200_floquet_sweep")
    return y.cpu().numpy(), hl.cpu().numpy(), tp.cpu().numpy(), nl.value


class PairOperators:
    """matDiagonal + dense matR[o] staged once on the device (scatter of the COO
    entries; the reference keeps scipy CSR, calculations_atom_pairstate.py:3222-3239)."""

    def __init__(self, h0diag, matR_coo, N, device=None):
        import torch
        lib = _lib.load()
        self.device = _lib.require_device(device)
        dev = torch.device("cuda", self.device)
        self.N = int(N)
        self.d_h0 = torch.as_tensor(np.ascontiguousarray(h0diag, dtype=np.float64)).to(dev)
        self.d_M = []
        self.h2d_bytes = self.d_h0.numel() * 8
        for rows, cols, vals in matR_coo:
            rows = torch.as_tensor(np.ascontiguousarray(rows, dtype=np.int32)).to(dev)
            cols = torch.as_tensor(np.ascontiguousarray(cols, dtype=np.int32)).to(dev)
            vals = torch.as_tensor(np.ascontiguousarray(vals, dtype=np.float64)).to(dev)
            self.h2d_bytes += rows.numel() * 4 + cols.numel() * 4 + vals.numel() * 8
            dense = torch.zeros((self.N, self.N), dtype=torch.float64, device=dev)
            if vals.numel():
                rc = lib.arcb200_scatter_coo(self.device, _lib.stream_ptr(self.device),
                                             self.N, vals.numel(), _lib.ptr(rows),
                                             _lib.ptr(cols), _lib.ptr(vals), _lib.ptr(dense))
                _lib.check(rc, "arcb200_scatter_coo")
            self.d_M.append(dense)
        self.d_Mptrs = torch.tensor([int(m.data_ptr()) for m in self.d_M],
                                    dtype=torch.int64, device=dev)


def pair_sweep(ops, rangeR, k, sigma_ghz, opi, drive_w=None, topk=4, contrib_max=0.95, batch=None,
               group=None, to_host=True, want_overlap=False):
    """k eigenpairs nearest sigma of H(R) for every R (already sorted by the caller,
    calculations_atom_pairstate.py:3307, 3425-3520).
    group=None: all points on this GPU; group=WORLD / a ProcessGroup: collective, sharded."""
    import torch
    lib = _lib.load()
    device = ops.device
    dev = torch.device("cuda", device)
    rangeR = np.ascontiguousarray(rangeR, dtype=np.float64)
    nR_all = len(rangeR)
    rank, world = dist_info(group)
    lo, hi = shard_bounds(nR_all, rank, world)
    # overlaps chain consecutive points: a shard also solves the point before its first one
    # (one redundant solve per rank instead of a peer-to-peer exchange of N x k eigenvectors)
    lead = 1 if (want_overlap and lo > 0 and hi > lo) else 0
    lo -= lead
    nR = hi - lo
    N = ops.N
    res = SweepResult()
    res.overlap = None
    d_R = torch.as_tensor(rangeR[lo:hi].copy()).to(dev)
    d_w = None
    if drive_w is not None:
        d_w = torch.as_tensor(np.ascontiguousarray(drive_w, dtype=np.float64)).to(dev)
    res.h2d_bytes = ops.h2d_bytes + d_R.numel() * 8
    y = torch.empty((nR, k), dtype=torch.float64, device=dev)
    hl = torch.empty((nR, k), dtype=torch.float64, device=dev)
    comp_c = torch.zeros((nR, k, max(topk, 1)), dtype=torch.float64, device=dev)
    comp_i = torch.full((nR, k, max(topk, 1)), -1, dtype=torch.int32, device=dev)
    ov = torch.empty((nR, k, k), dtype=torch.float64, device=dev) if want_overlap else None
    nl = _lib.LaunchCount()
    if nR > 0:
        b = _pick_batch(N, nR, k, device, batch)
        wbytes = int(lib.arcb200_sweep_workspace_bytes(N, b, k)) + (N * k * 8 + 256 if want_overlap else 0)
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        rc = lib.arcb200_pair_sweep(device, _lib.stream_ptr(device), N, _lib.ptr(ops.d_h0),
                                    len(ops.d_M), _lib.ptr(ops.d_Mptrs), nR, _lib.ptr(d_R),
                                    int(k), float(sigma_ghz), int(opi), _lib.ptr(d_w),
                                    int(topk), float(contrib_max), _lib.ptr(y), _lib.ptr(hl),
                                    _lib.ptr(comp_c),
                                    _lib.ptr(comp_i), _lib.ptr(ov), _lib.ptr(work), wbytes, b, nl.addr)
        _lib.check(rc, "arcb200_pair_sweep")
        res.kernel_launches = nl.value
    else:
        res.kernel_launches = 0
    if lead:      # drop the redundant leading point (its overlap row belongs to the previous rank)
        y, hl, comp_c, comp_i, ov = y[1:], hl[1:], comp_c[1:], comp_i[1:], ov[1:]
    if world > 1:
        y, hl = gather_rows(y, nR_all, group), gather_rows(hl, nR_all, group)
        comp_c, comp_i = gather_rows(comp_c, nR_all, group), gather_rows(comp_i, nR_all, group)
        if ov is not None:
            ov = gather_rows(ov, nR_all, group)
    if ov is not None:
        res.overlap = ov.cpu().numpy() if to_host else ov
    if to_host:
        res.y, res.hl = y.cpu().numpy(), hl.cpu().numpy()
        res.comp_c, res.comp_i = comp_c.cpu().numpy(), comp_i.cpu().numpy()
        res.d2h_bytes = res.y.nbytes + res.hl.nbytes + res.comp_c.nbytes + res.comp_i.nbytes
    else:
        res.y, res.hl, res.comp_c, res.comp_i = y, hl, comp_c, comp_i
        res.d2h_bytes = 0
    return res


def syevd_batched(A, device=None):
    """numpy.linalg.eigh replacement for a stack of symmetric matrices [batch,N,N].
    Returns (w[batch,N] ascending, v[batch,N,N] with v[b][:, i] the i-th eigenvector)."""
    import torch
    lib = _lib.load()
    device = _lib.require_device(device)
    dev = torch.device("cuda", device)
    A = np.ascontiguousarray(A, dtype=np.float64)
    if A.ndim == 2:
        A = A[None]
    batch, N, _ = A.shape
    d_A = torch.as_tensor(A).to(dev)
    d_w = torch.empty((batch, N), dtype=torch.float64, device=dev)
    d_Z = torch.empty((batch, N, N), dtype=torch.float64, device=dev)
    wbytes = int(lib.arcb200_sweep_workspace_bytes(N, batch, N))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    rc = lib.arcb200_syevd_batched(device, _lib.stream_ptr(device), N, batch, _lib.ptr(d_A),
                                   _lib.ptr(d_w), _lib.ptr(d_Z), _lib.ptr(work), wbytes, 0)
    _lib.check(rc, "arcb200_syevd_batched")
    return d_w.cpu().numpy(), np.ascontiguousarray(d_Z.cpu().numpy().transpose(0, 2, 1))


def sytrd_batched(A, cluster=0, device=None):
    """Stage 1 only: (d, e, tau) of the Householder tridiagonalisation of each A[b]."""
    import torch
    lib = _lib.load()
    device = _lib.require_device(device)
    dev = torch.device("cuda", device)
    A = np.ascontiguousarray(A, dtype=np.float64)
    if A.ndim == 2:
        A = A[None]
    batch, N, _ = A.shape
    d_A = torch.as_tensor(A).to(dev)
    outs = [torch.zeros((batch, N), dtype=torch.float64, device=dev) for _ in range(3)]
    wbytes = int(lib.arcb200_sweep_workspace_bytes(N, batch, N))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    rc = lib.arcb200_sytrd_batched(device, _lib.stream_ptr(device), N, batch, _lib.ptr(d_A),
                                   _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]),
                                   _lib.ptr(work), wbytes, int(cluster), 0)
    _lib.check(rc, "arcb200_sytrd_batched")
    return [o.cpu().numpy() for o in outs]


def sy2sb_batched(A, device=None):
    """Stage 1 of the two-stage path only: the symmetric band (width 32) each A[b] is reduced
    to, returned as band[b, j, d] = B[j+d, j]."""
    import torch
    lib = _lib.load()
    device = _lib.require_device(device)
    dev = torch.device("cuda", device)
    A = np.ascontiguousarray(A, dtype=np.float64)
    if A.ndim == 2:
        A = A[None]
    batch, N, _ = A.shape
    d_A = torch.as_tensor(A).to(dev)
    band = torch.zeros((batch, N, 33), dtype=torch.float64, device=dev)
    wbytes = int(lib.arcb200_sweep_workspace_bytes(N, batch, N))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    rc = lib.arcb200_sy2sb_batched(device, _lib.stream_ptr(device), N, batch, _lib.ptr(d_A),
                                   _lib.ptr(band), _lib.ptr(work), wbytes, 0)
    _lib.check(rc, "arcb200_sy2sb_batched")
    return band.cpu().numpy()


def stedc_batched(d, e, device=None):
    """Stage 2 only: eigen-decomposition of symmetric tridiagonal matrices (d, e)."""
    import torch
    lib = _lib.load()
    device = _lib.require_device(device)
    dev = torch.device("cuda", device)
    d = np.ascontiguousarray(np.atleast_2d(d), dtype=np.float64)
    batch, N = d.shape
    ee = np.zeros((batch, N))
    ee[:, :N - 1] = np.atleast_2d(e)[:, :N - 1]
    d_d, d_e = torch.as_tensor(d).to(dev), torch.as_tensor(ee).to(dev)
    d_w = torch.empty((batch, N), dtype=torch.float64, device=dev)
    d_Q = torch.empty((batch, N, N), dtype=torch.float64, device=dev)
    wbytes = int(lib.arcb200_sweep_workspace_bytes(N, batch, N))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    rc = lib.arcb200_stedc_batched(device, _lib.stream_ptr(device), N, batch, _lib.ptr(d_d),
                                   _lib.ptr(d_e), _lib.ptr(d_w), _lib.ptr(d_Q), _lib.ptr(work),
                                   wbytes, 0)
    _lib.check(rc, "arcb200_stedc_batched")
    return d_w.cpu().numpy(), np.ascontiguousarray(d_Q.cpu().numpy().transpose(0, 2, 1))
