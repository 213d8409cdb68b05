"""Mono-kernel band reduction (one CTA per matrix, >= 74 matrices per launch) on mid-sized
matrices: eigenvalues of every matrix of the batch against numpy, residuals for a few."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arc_b200 import sweep
rng = np.random.default_rng(77)
for N, nb in ((1500, 80), (1029, 150)):
    A = rng.standard_normal((nb, N, N)) * rng.random((nb, 1, 1)) * 10
    A = A + A.transpose(0, 2, 1)
    w, Z = sweep.syevd_batched(A)
    worst = 0.0
    for b in range(nb):
        w0 = np.linalg.eigvalsh(A[b])
        worst = max(worst, np.abs(w[b] - w0).max() / np.abs(w0).max())
    res = max(np.abs(A[b] @ Z[b] - Z[b] * w[b]).max() / np.abs(w[b]).max() for b in (0, nb // 2, nb - 1))
    orth = max(np.abs(Z[b].T @ Z[b] - np.eye(N)).max() for b in (0, nb - 1))
    print("N=%d batch=%d: worst eigenvalue error %.2e, residual %.2e, orthogonality %.2e" % (N, nb, worst, res, orth), flush=True)
