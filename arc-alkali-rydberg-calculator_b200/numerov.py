"""Drop-in for the reference's Numerov entry points (seam B1).

`NumerovWavefunction` keeps the 18-positional-argument signature of the C extension
(arc/arc_c_extensions.c:113-116, bound at alkali_atom_functions.py:240-243) and
`NumerovBack` the signature and behaviour of the pure-Python twin
(alkali_atom_functions.py:3939-4063, arbitrary callable kfun); both run CUDA kernels.
"""
import numpy as np

from . import radial as _radial


def NumerovWavefunction(innerLimit, outerLimit, step, init1, init2, l, s, j,
                        stateEnergy, alphaC, alpha, Z, a1, a2, a3, a4, rc, mu):
    """Returns a C-contiguous float64 array d of shape (2, L): d[0] = R(r)*r
    (un-normalised, zero inside the divergence point), d[1] = r, index 0 innermost --
    the two rows callers use from the reference's (mis-shaped) return value
    (alkali_atom_functions.py:603-604).  Raises RuntimeError where the reference
    returns NULL (arc_c_extensions.c:145,200,258)."""
    rec = _radial.make_state(innerLimit, outerLimit, step, init1, init2, l, s, j,
                             stateEnergy, alphaC, alpha, Z, a1, a2, a3, a4, rc, mu)
    rb = _radial.RadialBatch(np.array([rec]), normalise=False)
    if int(rb.lengths[0]) < 3 or int(rb.divergence_points()[0]) < 0:
        raise RuntimeError("Numerov integration failed (grid too short)")
    rb.raise_on_failure()
    r, psi = rb.wavefunction(0)
    return np.ascontiguousarray(np.stack([psi, r]))


def _eval_on_grid(kfun, x):
    """kfun on a numpy grid: one vectorised call when the callable accepts arrays (numpy
    expressions do), else the reference's own scalar calls (math.exp based potentials)."""
    try:
        v = np.asarray(kfun(x), dtype=np.float64)
        if v.shape == x.shape:
            return v
    except Exception:
        pass
    return np.array([float(kfun(float(t))) for t in x], dtype=np.float64)


def NumerovBack(innerLimit, outerLimit, kfun, step, init1, init2):
    """Drop-in for the pure-Python integrator alkali_atom_functions.py:3939-4063: same
    signature, same return value (r, rad(r)*r, unnormalised), same grid (`round` of the length,
    start at sqrt(inner) + step*(L-2), aaf:3986-3991), divergence control and tail rule.

    `kfun` is an arbitrary Python callable, which a CUDA kernel cannot call, so the host
    evaluates it on the grid (vectorised when the callable takes arrays) and forms the three
    Numerov coefficients per point with the reference's operation order; the recurrence with
    its IEEE division and the divergence logic run in `numerov_coeff_kernel` (one warp per
    state).  A callable carrying `arcb200_params` = (l, s, j, stateEnergy, alphaC, alpha, Z, a1,
    a2, a3, a4, rc, mu) is routed to the model-potential kernel of the C path instead
    (NumerovWavefunction), which also evaluates k(x) on the GPU."""
    p = getattr(kfun, "arcb200_params", None)
    if p is not None:
        d = NumerovWavefunction(innerLimit, outerLimit, step, init1, init2, *p)
        return d[1], d[0]
    import torch
    from . import _lib
    lib = _lib.load()
    device = _lib.require_device(None)
    dev = torch.device("cuda", device)
    L = int(round((np.sqrt(outerLimit) - np.sqrt(innerLimit)) / step))      # aaf:3984
    if L < 3:
        raise RuntimeError("NumerovBack: grid too short")
    # x_{L-1} = sqrt(inner) + step*(L-2) (aaf:3990-3991), then x -= step: sequential rounding
    seq = np.full(L, -float(step))
    seq[0] = np.sqrt(innerLimit) + step * (L - 2)
    x = np.add.accumulate(seq)[::-1].copy()                                 # x[i], i = 0 .. L-1
    k0, kp, km = _eval_on_grid(kfun, x), _eval_on_grid(kfun, x + step), _eval_on_grid(kfun, x - step)
    step2 = step ** 2
    A = 2.0 * (1.0 - 5.0 / 12.0 * step2 * k0)                               # aaf:3995, 4019
    B = 1.0 + 1.0 / 12.0 * step2 * kp
    D = 1 + 1 / 12.0 * step2 * km
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)   # noqa: E731
    d_off, d_len = t([0, L + 32], torch.int64), t([L], torch.int32)
    d_A, d_B, d_D = t(A), t(B), t(D)
    d_xtop, d_step, d_in = t([x[L - 1]]), t([float(step)]), t([float(init1), float(init2)])
    d_psi = torch.zeros(L + 64, dtype=torch.float64, device=dev)
    d_r = torch.zeros(L + 64, dtype=torch.float64, device=dev)
    d_dp = torch.zeros(1, dtype=torch.int32, device=dev)
    d_st = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.arcb200_numerov_coeff_batch(device, _lib.stream_ptr(device), 1, _lib.ptr(d_off), _lib.ptr(d_len),
                                         _lib.ptr(d_A), _lib.ptr(d_B), _lib.ptr(d_D), _lib.ptr(d_xtop),
                                         _lib.ptr(d_step), _lib.ptr(d_in), _lib.ptr(d_psi), _lib.ptr(d_r),
                                         _lib.ptr(d_dp), _lib.ptr(d_st))
    _lib.check(rc, "arcb200_numerov_coeff_batch")
    if int(d_st.cpu()[0]) != 0:
        raise RuntimeError("Numerov error")                                  # aaf:4045-4047 (exit())
    return d_r[:L].cpu().numpy(), d_psi[:L].cpu().numpy()
