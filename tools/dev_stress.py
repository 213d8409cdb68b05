"""Randomised parity stress of the two-stage eigensolver path on the GPU: many sizes, both band
reduction kernels, full spectrum and windows, structured matrices."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ARCB200_TWO_STAGE"] = "1"
import numpy as np
from arc_b200 import sweep
rng = np.random.default_rng(123)
worst = 0.0
for N in (96, 97, 127, 129, 160, 191, 225, 257, 320, 449, 515, 777, 1031):
    for multi in ("0", "1"):
        os.environ["ARCB200_SY2SB_MULTI"] = multi
        kind = rng.integers(0, 3)
        A = rng.standard_normal((2, N, N))
        if kind == 1:       # banded-ish + clusters
            A *= (np.abs(np.subtract.outer(np.arange(N), np.arange(N))) < 40)
        if kind == 2:       # graded diagonal, weak coupling
            A = A * 1e-3 + np.diag(np.logspace(-6, 3, N))[None]
        A = A + A.transpose(0, 2, 1)
        w, Z = sweep.syevd_batched(A)
        w0 = np.linalg.eigvalsh(A)
        scale = np.abs(w0).max()
        e1 = np.abs(w - w0).max() / scale
        e2 = max(np.abs(A[b] @ Z[b] - Z[b] * w[b]).max() for b in range(2)) / scale
        e3 = max(np.abs(Z[b].T @ Z[b] - np.eye(N)).max() for b in range(2))
        worst = max(worst, e1, e2, e3)
        flag = "" if max(e1, e2) < 1e-11 and e3 < 1e-11 else "  <-- CHECK"
        print("N=%4d multi=%s kind=%d: eig %.1e res %.1e orth %.1e%s" % (N, multi, kind, e1, e2, e3, flag), flush=True)
print("worst", worst)
