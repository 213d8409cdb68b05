import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.linalg import eigvalsh_tridiagonal
from arc_b200 import sweep
rng = np.random.default_rng(0)
for N in [int(a) for a in sys.argv[1:]]:
    M = rng.standard_normal((2, N, N)); M = M + M.transpose(0, 2, 1)
    wr = np.linalg.eigvalsh(M[0])
    for cl in (1, 2, 4, 8):
        d, e, tau = sweep.sytrd_batched(M, cluster=cl)
        wt = eigvalsh_tridiagonal(d[0], e[0][:N - 1])
        print("N=%d cluster=%d sytrd: eig(T) vs eig(A) rel err %.2e" % (N, cl, np.abs(wt - wr).max() / np.abs(wr).max()), flush=True)
    # reference tridiagonal from scipy hessenberg-like: use numpy on T built from d,e of cluster=1
    d, e, tau = sweep.sytrd_batched(M, cluster=1)
    w, Q = sweep.stedc_batched(d, e)
    T = np.diag(d[0]) + np.diag(e[0][:N - 1], 1) + np.diag(e[0][:N - 1], -1)
    wt = np.linalg.eigvalsh(T)
    print("N=%d stedc: eigval rel err %.2e resid %.2e orth %.2e" % (
        N, np.abs(w[0] - wt).max() / np.abs(wt).max(),
        np.abs(T @ Q[0] - Q[0] * w[0][None, :]).max() / np.abs(wt).max(),
        np.abs(Q[0].T @ Q[0] - np.eye(N)).max()), flush=True)
