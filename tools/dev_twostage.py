"""Stage-by-stage check of the two-stage eigensolver path on the GPU (run with
ARCB200_TWO_STAGE=1 for small N).  Usage: python tools/dev_twostage.py N batch [time_batch]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.linalg import eig_banded, eigvalsh_tridiagonal
from arc_b200 import sweep, _lib
N = int(sys.argv[1]); nb = int(sys.argv[2]); tb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rng = np.random.default_rng(0)
A = rng.standard_normal((nb, N, N)); A = A + A.transpose(0, 2, 1)
w0 = np.linalg.eigvalsh(A)
scale = np.abs(w0).max()
def stage(name, fn):
    try:
        t0 = time.time(); fn(); torch.cuda.synchronize()
        print("  [%s ok, %.2fs]" % (name, time.time() - t0), flush=True)
    except Exception as e:
        print("  [%s FAILED: %s]" % (name, e), flush=True)
def s1():
    band = sweep.sy2sb_batched(A)            # [b, j, d]
    err = 0
    for b in range(nb):
        ab = band[b].T.copy()                # lower form: ab[d, j]
        wb = eig_banded(ab, lower=True, eigvals_only=True)
        err = max(err, np.abs(wb - w0[b]).max())
    print("stage 1 band eigenvalues: max err %.3e (scale %.2e)" % (err, scale))
def s2():
    d, e, tau = sweep.sytrd_batched(A)
    err = max(np.abs(eigvalsh_tridiagonal(d[b], e[b][:N - 1]) - w0[b]).max() for b in range(nb))
    print("stage 1+2 tridiagonal eigenvalues: max err %.3e" % err)
def s3():
    w, Z = sweep.syevd_batched(A)
    print("syevd eigenvalues: max err %.3e" % np.abs(w - w0).max())
    res = max(np.abs(A[b] @ Z[b] - Z[b] * w[b]).max() for b in range(nb))
    orth = max(np.abs(Z[b].T @ Z[b] - np.eye(N)).max() for b in range(nb))
    print("syevd residual %.3e orthogonality %.3e" % (res, orth))
stage("sy2sb", s1); stage("sy2sb+sb2st", s2); stage("syevd", s3)
if tb:
    lib = _lib.load(); dev = torch.device("cuda", 0)
    A2 = rng.standard_normal((tb, N, N)); A2 = A2 + A2.transpose(0, 2, 1)
    d_A = torch.as_tensor(A2).to(dev)
    outs = [torch.zeros((tb, N), dtype=torch.float64, device=dev) for _ in range(3)]
    band = torch.zeros((tb, N, 33), dtype=torch.float64, device=dev)
    wbytes = int(lib.arcb200_sweep_workspace_bytes(N, tb, N)); work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    def f1(): _lib.check(lib.arcb200_sy2sb_batched(0, _lib.stream_ptr(0), N, tb, _lib.ptr(d_A), _lib.ptr(band), _lib.ptr(work), wbytes, 0), "sy2sb")
    def f2(): _lib.check(lib.arcb200_sytrd_batched(0, _lib.stream_ptr(0), N, tb, _lib.ptr(d_A), _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]), _lib.ptr(work), wbytes, 0, 0), "sytrd")
    for name, f in (("sy2sb", f1), ("sy2sb+sb2st", f2)):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("%s N=%d nb=%d: %.1f ms (%.3f ms/matrix)" % (name, N, tb, ms, ms / tb))
