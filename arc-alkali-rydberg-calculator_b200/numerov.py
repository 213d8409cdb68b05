"""Drop-in for the reference's Numerov entry points (seam B1).

`NumerovWavefunction` keeps the 18-positional-argument signature of the C extension
(arc/arc_c_extensions.c:113-116, bound at alkali_atom_functions.py:240-243) and
`NumerovBack` the signature of the pure-Python twin (alkali_atom_functions.py:3939-4063);
both run the batched CUDA kernel (a batch of one state).
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
    r, psi = rb.wavefunction(0)
    return np.ascontiguousarray(np.stack([psi, r]))


def NumerovBack(innerLimit, outerLimit, kfun, step, init1, init2):
    """Signature of alkali_atom_functions.py:3939.  The reference's Python twin takes an
    arbitrary callable kfun(x); a GPU kernel cannot call back into Python, and the
    Python twin is not the parity target (it differs from the C path by 6e-5),
    so this entry point accepts only callables that carry the 12 physical parameters of
    the C path as attribute `arcb200_params` = (l, s, j, stateEnergy, alphaC, alpha, Z,
    a1, a2, a3, a4, rc, mu) and routes them to the CUDA kernel.  Returns (r, psi)."""
    p = getattr(kfun, "arcb200_params", None)
    if p is None:
        raise NotImplementedError(
            "NumerovBack on the B200 path needs kfun.arcb200_params "
            "(l,s,j,E,alphaC,alpha,Z,a1,a2,a3,a4,rc,mu); arbitrary Python callables "
            "cannot run inside the CUDA kernel and there is no CPU fallback")
    d = NumerovWavefunction(innerLimit, outerLimit, step, init1, init2, *p)
    return d[1], d[0]
