"""A/B of the band-reduction kernel choice for batches that do not fill the GPU with one CTA per
matrix: ARCB200_SY2SB_MULTI=0 (one CTA per matrix) / 1 (per-phase kernels, many CTAs per matrix).
    python tools/ab_multi.py N batch"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arc_b200 import _lib
N, nb = int(sys.argv[1]), int(sys.argv[2])
lib = _lib.load()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(1)
A = rng.standard_normal((min(nb, 4), N, N)); A = A + A.transpose(0, 2, 1)
d_A = torch.as_tensor(A).to(dev).repeat((nb + len(A) - 1) // len(A), 1, 1)[:nb].contiguous()
wbytes = int(lib.arcb200_sweep_workspace_bytes(N, nb, N))
work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
band = torch.zeros((nb, N, 33), dtype=torch.float64, device=dev)
out = {"N": N, "batch": nb}
bands = {}
cases = (("one_cta_per_matrix", "0"), ("many_ctas_per_matrix", "1"))
if os.environ.get("ARCB200_SY2SB_VARIANT"):
    cases = cases[:1]
if os.environ.get("AB_ONLY") == "multi":
    cases = cases[1:]
for tag, env in cases:
    os.environ["ARCB200_SY2SB_MULTI"] = env
    def run():
        _lib.check(lib.arcb200_sy2sb_batched(0, _lib.stream_ptr(0), N, nb, _lib.ptr(d_A), _lib.ptr(band), _lib.ptr(work), wbytes, 0), "sy2sb")
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    out[tag] = {"ms": ms, "tflops": (4.0 / 3.0) * N ** 3 * nb / (ms * 1e-3) / 1e12}
    bands[tag] = band.clone()
if len(bands) == 2:
    b0, b1 = bands.values()
    out["band_maxdiff"] = float((b0 - b1).abs().max())
    # eigenvalues of the first matrix from the band of the many-CTA path against LAPACK on the dense matrix
    import scipy.linalg as sl
    bb = b1[0].cpu().numpy().T.copy()          # (33, N) lower band storage
    out["eig_err_multi_vs_lapack"] = float(np.abs(sl.eigvals_banded(bb, lower=True) - np.linalg.eigvalsh(A[0])).max() / np.abs(A[0]).max())
print(json.dumps(out))
