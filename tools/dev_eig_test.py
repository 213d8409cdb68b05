import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import arc_b200
from arc_b200 import sweep


def check(name, A):
    t0 = time.time()
    w, v = sweep.syevd_batched(A)
    torch.cuda.synchronize()
    dt = time.time() - t0
    A = A if A.ndim == 3 else A[None]
    worst = [0, 0, 0]
    for b in range(A.shape[0]):
        wr = np.linalg.eigvalsh(A[b])
        scale = max(np.abs(wr).max(), 1e-300)
        e_w = np.abs(w[b] - wr).max() / scale
        res = np.abs(A[b] @ v[b] - v[b] * w[b][None, :]).max() / scale
        orth = np.abs(v[b].T @ v[b] - np.eye(A.shape[1])).max()
        worst = [max(worst[0], e_w), max(worst[1], res), max(worst[2], orth)]
    ok = worst[0] < 1e-12 and worst[1] < 1e-11 and worst[2] < 1e-11
    print("%-28s N=%4d batch=%3d  eigval %.1e  resid %.1e  orth %.1e  %s  (%.3fs)"
          % (name, A.shape[1], A.shape[0], worst[0], worst[1], worst[2], "OK" if ok else "FAIL", dt), flush=True)
    return ok


rng = np.random.default_rng(0)
allok = True
sizes = [int(a) for a in sys.argv[1:]] or [3, 8, 31, 32, 33, 64, 65, 100, 129, 257, 410]
for N in sizes:
    B = 3 if N < 300 else 2
    M = rng.standard_normal((B, N, N)); M = M + M.transpose(0, 2, 1)
    allok &= check("random", M)
    D = np.zeros((1, N, N)); D[0][np.arange(N), np.arange(N)] = rng.standard_normal(N)
    allok &= check("diagonal", D)
    if N >= 8:
        # clustered spectrum: few distinct eigenvalues
        Q, _ = np.linalg.qr(rng.standard_normal((N, N)))
        lam = np.repeat(rng.standard_normal(4), (N + 3) // 4)[:N]
        C = (Q * lam) @ Q.T; C = (C + C.T) / 2
        allok &= check("clustered", C[None])
        # weakly coupled diagonal (like a Stark map at small field)
        W = np.diag(np.sort(rng.standard_normal(N)) * 100) + 1e-3 * M[0]
        allok &= check("weak coupling", W[None])
print("ALL OK" if allok else "SOME FAILED")
