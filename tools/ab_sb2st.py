"""A/B of the bulge-chasing kernels (stage 2 of the two-stage tridiagonalisation):
ARCB200_SB2ST_SPLIT=0 (one warp per sweep) / 1 (E-chain warp + D-block warp per sweep).
Times arcb200_sytrd_batched (both stages) minus arcb200_sy2sb_batched (stage 1) on the same batch and
checks the eigenvalues of the tridiagonal matrix against LAPACK on the dense matrix.
    python tools/ab_sb2st.py N batch"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.linalg import eigvalsh_tridiagonal
from arc_b200 import _lib
N, nb = int(sys.argv[1]), int(sys.argv[2])
lib = _lib.load()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(2)
A = rng.standard_normal((min(nb, 3), N, N)); A = A + A.transpose(0, 2, 1)
d_A = torch.as_tensor(A).to(dev).repeat((nb + len(A) - 1) // len(A), 1, 1)[:nb].contiguous()
wbytes = int(lib.arcb200_sweep_workspace_bytes(N, nb, N))
work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
band = torch.zeros((nb, N, 33), dtype=torch.float64, device=dev)
outs = [torch.zeros((nb, N), dtype=torch.float64, device=dev) for _ in range(3)]
def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 2
def stage1():
    _lib.check(lib.arcb200_sy2sb_batched(0, _lib.stream_ptr(0), N, nb, _lib.ptr(d_A), _lib.ptr(band), _lib.ptr(work), wbytes, 0), "sy2sb")
def both():
    _lib.check(lib.arcb200_sytrd_batched(0, _lib.stream_ptr(0), N, nb, _lib.ptr(d_A), _lib.ptr(outs[0]), _lib.ptr(outs[1]),
                                         _lib.ptr(outs[2]), _lib.ptr(work), wbytes, 0, 0), "sytrd")
out = {"N": N, "batch": nb, "sy2sb_ms": timed(stage1)}
wref = [np.linalg.eigvalsh(a) for a in A]
res = {}
for tag, env in (("one_warp_per_sweep", "0"), ("split", "1")):
    os.environ["ARCB200_SB2ST_SPLIT"] = env
    ms = timed(both)
    d, e = outs[0].cpu().numpy(), outs[1].cpu().numpy()
    err = max(float(np.abs(eigvalsh_tridiagonal(d[m], e[m][:N - 1]) - wref[m % len(A)]).max() / np.abs(wref[m % len(A)]).max())
              for m in (0, 1, 2, nb - 1) if m < nb)
    res[tag] = (d.copy(), e.copy())
    out[tag] = {"both_stages_ms": ms, "sb2st_ms": ms - out["sy2sb_ms"], "eig_err_vs_lapack": err}
out["d_maxdiff"] = float(np.abs(res["split"][0] - res["one_warp_per_sweep"][0]).max())
out["e_maxdiff"] = float(np.abs(np.abs(res["split"][1]) - np.abs(res["one_warp_per_sweep"][1])).max())
print(json.dumps(out))
