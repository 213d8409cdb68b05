"""ctypes binding of libarcb200.so (the C-ABI declared in include/arcb200.h).

There is NO CPU fallback: if the CUDA library is missing or no B200 is visible the
calls raise.  PyTorch is used only for device buffers and streams.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libarcb200.so")
_lib = None

c_int, c_i32, c_i64 = ctypes.c_int, ctypes.c_int32, ctypes.c_int64
c_dbl, c_vp = ctypes.c_double, ctypes.c_void_p

# numpy mirror of arcb200_state_t (include/arcb200.h); 144 bytes
STATE_DTYPE = np.dtype([
    ("innerLimit", "f8"), ("outerLimit", "f8"), ("step", "f8"), ("init1", "f8"),
    ("init2", "f8"), ("s", "f8"), ("j", "f8"), ("stateEnergy", "f8"),
    ("alphaC", "f8"), ("alpha", "f8"), ("a1", "f8"), ("a2", "f8"), ("a3", "f8"),
    ("a4", "f8"), ("rc", "f8"), ("mu", "f8"), ("l", "i4"), ("Z", "i4"),
    ("length", "i4"), ("_pad", "i4")], align=True)
assert STATE_DTYPE.itemsize == 144
GRAM_DTYPE = np.dtype([("row0", "i4"), ("nrows", "i4"), ("col0", "i4"), ("ncols", "i4"),
                       ("power", "i4"), ("kmax", "i4"), ("out_off", "i8")], align=True)
assert GRAM_DTYPE.itemsize == 32

_SIGNATURES = {
    "arcb200_abi_version": (c_int, []),
    "arcb200_last_error": (ctypes.c_char_p, []),
    "arcb200_device_info": (c_int, [c_int, c_vp]),
    "arcb200_numerov_lengths": (c_i64, [c_i32, c_vp, c_vp]),
    "arcb200_numerov_batch": (c_int, [c_int, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                                      c_vp, c_vp, c_i32, c_vp]),
    "arcb200_numerov_coeff_batch": (c_int, [c_int, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                            c_vp, c_vp, c_vp, c_vp, c_vp]),
    "arcb200_radial_integrals": (c_int, [c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp,
                                         c_vp, c_vp]),
    "arcb200_radial_gram": (c_int, [c_int, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                    c_i32, c_vp, c_vp, c_i32, c_i32, c_i32]),
    "arcb200_semiclassical_table": (c_int, [c_int, c_vp, c_i64, c_vp, c_i32, c_vp, c_i32,
                                            c_vp]),
    "arcb200_sweep_workspace_bytes": (c_i64, [c_i32, c_i32, c_i32]),
    "arcb200_stark_sweep": (c_int, [c_int, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_i32,
                                    c_vp, c_i32, c_dbl, c_vp, c_vp, c_vp, c_vp, c_vp,
                                    c_i64, c_i32, c_vp]),
    "arcb200_pair_sweep": (c_int, [c_int, c_vp, c_i32, c_vp, c_i32, c_vp, c_i32, c_vp,
                                   c_i32, c_dbl, c_i32, c_vp, c_i32, c_dbl, c_vp, c_vp, c_vp,
                                   c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "arcb200_scatter_coo": (c_int, [c_int, c_vp, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "arcb200_syevd_batched": (c_int, [c_int, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp,
                                      c_i64, c_vp]),
    "arcb200_sytrd_batched": (c_int, [c_int, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                                      c_i64, c_i32, c_vp]),
    "arcb200_stedc_batched": (c_int, [c_int, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                                      c_i64, c_vp]),
    "arcb200_floquet_sweep": (c_int, [c_int, c_vp, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_i32,
                                      c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "arcb200_resonance_scan": (c_int, [c_int, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp,
                                       c_dbl, c_dbl, c_dbl, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "arcb200_dmma_peak": (c_int, [c_int, c_vp, c_vp, c_vp]),
    "arcb200_dfma_peak": (c_int, [c_int, c_vp, c_vp, c_vp]),
    "arcb200_tile_stream_probe": (c_int, [c_int, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_dbl, c_vp]),
    "arcb200_sy2sb_batched": (c_int, [c_int, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_i64, c_vp]),
}


class ArcB200Error(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Load libarcb200.so; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ArcB200Error(
                "libarcb200.so not found at %s -- build it with "
                "`python arc-alkali-rydberg-calculator_b200/build.py` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.arcb200_abi_version() != 2:
            raise ArcB200Error("libarcb200 ABI version mismatch")
        _lib = lib
    return _lib


def check(rc, what):
    """0 = success; anything else is an error (include/arcb200.h)."""
    if rc != 0:
        raise ArcB200Error("%s failed (%d): %s"
                           % (what, rc, load().arcb200_last_error().decode()))
    return rc


class LaunchCount:
    """Host int32 for the trailing `h_launches` out-parameter of the sweep entry points."""

    def __init__(self):
        self._v = ctypes.c_int32(0)

    @property
    def addr(self):
        return ctypes.addressof(self._v)

    @property
    def value(self):
        return int(self._v.value)


_dev_checked = set()


def require_device(device=None):
    """Return the CUDA ordinal to use; raise if no sm_100 GPU is visible."""
    import torch
    if not torch.cuda.is_available():
        raise ArcB200Error("no CUDA device visible: arc_b200 has no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    if device not in _dev_checked:
        props = (c_i64 * 8)()
        check(load().arcb200_device_info(int(device), ctypes.addressof(props)),
              "arcb200_device_info")
        _dev_checked.add(device)
    return int(device)


def stream_ptr(device):
    import torch
    return int(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """device pointer of a torch tensor (or 0 for None)."""
    return 0 if t is None else int(t.data_ptr())
