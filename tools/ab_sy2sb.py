"""A/B of the band-reduction kernels (stage 1 of the two-stage tridiagonalisation) through the
C-ABI: ARCB200_SY2SB_VARIANT=lockstep (round-1 kernel: block barrier per tile step, per-thread
cp.async) against the default (decoupled warpgroups + TMA-unit bulk copies).  Prints one JSON line.
    python tools/ab_sy2sb.py [N] [batch] [--small]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from arc_b200 import _lib

small = "--small" in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if args else (200 if small else 2363)
nb = int(args[1]) if len(args) > 1 else (3 if small else 296)
if small:
    os.environ["ARCB200_TWO_STAGE"] = "1"
    os.environ["ARCB200_SY2SB_MULTI"] = "0"
lib = _lib.load()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(1)
A = rng.standard_normal((nb, N, N))
A = A + A.transpose(0, 2, 1)
d_A = torch.as_tensor(A).to(dev)
wbytes = int(lib.arcb200_sweep_workspace_bytes(N, nb, N))
work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
out = {"N": N, "batch": nb}
bands = {}
for variant in ("lockstep", "wg", "fused", "ring", "per_phase"):
    os.environ["ARCB200_SY2SB_VARIANT"] = variant
    os.environ["ARCB200_SY2SB_MULTI"] = "1" if variant == "per_phase" else "0"      # per_phase = the default path
    band = torch.zeros((nb, N, 33), dtype=torch.float64, device=dev)

    def run():
        _lib.check(lib.arcb200_sy2sb_batched(0, _lib.stream_ptr(0), N, nb, _lib.ptr(d_A), _lib.ptr(band),
                                             _lib.ptr(work), wbytes, 0), "sy2sb")
    run()
    torch.cuda.synchronize()
    reps = 1 if small else 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bands[variant] = band.cpu().numpy()
    out[variant] = {"ms": ms, "ms_per_matrix": ms / nb, "tflops": (4.0 / 3.0) * N ** 3 * nb / (ms * 1e-3) / 1e12}
# parity: eigenvalues of the band of matrix 0 against LAPACK on the dense matrix, and the two variants
b = bands["fused"][0]
T = np.zeros((N, N))
for d in range(33):
    idx = np.arange(N - d)
    T[idx + d, idx] = b[:N - d, d]
    T[idx, idx + d] = b[:N - d, d]
w_band = np.linalg.eigvalsh(T)
w_ref = np.linalg.eigvalsh(A[0])
out["eig_err_fused_vs_lapack"] = float(np.abs(w_band - w_ref).max() / np.abs(w_ref).max())
# the band itself is not unique to rounding (reflector signs are, but the reductions of the two
# algorithms differ in summation order): compare the spectra of all matrices of the batch instead
def band_eigs(bb):
    T = np.zeros((N, N))
    for d in range(33):
        idx = np.arange(N - d)
        T[idx + d, idx] = bb[:N - d, d]
        T[idx, idx + d] = bb[:N - d, d]
    return np.linalg.eigvalsh(T)
worst = 0.0
for m in range(min(nb, 3)):
    wr = np.linalg.eigvalsh(A[m])
    for v in ("wg", "fused", "ring", "per_phase"):
        worst = max(worst, float(np.abs(band_eigs(bands[v][m]) - wr).max() / np.abs(wr).max()))
out["eig_err_worst_of_3_matrices_wg_fused"] = worst
out["band_maxdiff_wg_vs_lockstep"] = float(np.abs(bands["wg"] - bands["lockstep"]).max())
out["band_maxdiff_fused_vs_lockstep"] = float(np.abs(bands["fused"] - bands["lockstep"]).max())
out["speedup_wg"] = out["lockstep"]["ms"] / out["wg"]["ms"]
out["speedup_fused"] = out["lockstep"]["ms"] / out["fused"]["ms"]
out["speedup_ring"] = out["lockstep"]["ms"] / out["ring"]["ms"]
out["speedup_per_phase"] = out["lockstep"]["ms"] / out["per_phase"]["ms"]
out["band_maxdiff_per_phase_vs_lockstep"] = float(np.abs(bands["per_phase"] - bands["lockstep"]).max())
out["band_maxdiff_ring_vs_lockstep"] = float(np.abs(bands["ring"] - bands["lockstep"]).max())
print(json.dumps(out))
