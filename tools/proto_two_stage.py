"""numpy model of the two-stage tridiagonalisation used by csrc/eig_sy2sb.cu / eig_sb2st.cu /
eig_q2.cu (same blocking, storage and application orders), checked against numpy.linalg.eigh.
Design aid and documentation of the index conventions; not imported by the product.

  stage 1  dense -> band (bandwidth b): per block column p, Householder QR of the panel below the
           band, compact WY (V, T), two-sided update  A22 -= V W^T + W V^T,
           W = X T - 1/2 V (T^T (V^T X) T),  X = A22 V.
           Reflector q = b p + c is stored in A[q+b+1:, q] with an implicit 1 at row q + b.
  stage 2  band -> tridiagonal by bulge chasing; sweep s, step t uses a reflector on rows
           [s+1+t b, s+1+(t+1) b); all reflectors of sweep s are stored contiguously in
           V2[s+1:, s] (explicit leading 1), tau2[s, t].
  back     Z = Q1 (Q2 Z_T).  Q2 = prod_s prod_t H(s,t) (generation order).  Q2 Z is applied group
           by group (g consecutive sweeps, last group first); inside a group the blocks
           t = 0, 1, ... in ascending order, inside a block the sweeps in descending order.
"""
import sys
import numpy as np


def house(x):
    """LAPACK dlarfg: returns (v with v[0]=1, tau, beta)."""
    alpha = x[0]
    xn = np.linalg.norm(x[1:]) if len(x) > 1 else 0.0
    if xn == 0.0:
        v = np.zeros_like(x); v[0] = 1.0
        return v, 0.0, alpha
    beta = -np.copysign(np.hypot(alpha, xn), alpha)
    tau = (beta - alpha) / beta
    v = x / (alpha - beta); v[0] = 1.0
    return v, tau, beta


def larft(V, tau):
    k = V.shape[1]
    G = V.T @ V
    T = np.zeros((k, k))
    for i in range(k):
        T[:i, i] = -tau[i] * (T[:i, :i] @ G[:i, i])
        T[i, i] = tau[i]
    return T


def sy2sb(A, b):
    """A: (Npad, Npad) symmetric, Npad multiple of b.  Returns (A with band + V1, tau1)."""
    A = A.copy()
    n = A.shape[0]
    NB = n // b
    tau1 = np.zeros(n)
    for p in range(NB - 1):
        m0 = b * (p + 1)
        P = A[m0:, b * p:b * p + b]          # view (m x b)
        m = n - m0
        V = np.zeros((m, b))
        for c in range(b):
            if c >= m - 1:
                # no sub-diagonal part: H = I
                if c < m:
                    V[c, c] = 1.0
                continue
            v, tau, beta = house(P[c:, c].copy())
            w = v @ P[c:, c + 1:]
            P[c:, c + 1:] -= tau * np.outer(v, w)
            P[c, c] = beta
            P[c + 1:, c] = v[1:]
            V[c:, c] = v
            tau1[b * p + c] = tau
        T = larft(V, tau1[b * p:b * p + b])
        A22 = A[m0:, m0:]
        A22s = np.tril(A22) + np.tril(A22, -1).T        # only the lower triangle is valid
        X = A22s @ V
        S = V.T @ X
        W = X @ T - 0.5 * V @ (T.T @ S @ T)
        A22 -= V @ W.T + W @ V.T
    return A, tau1


def band_of(A, N, b):
    """N x N symmetric band matrix from the lower triangle of the stage-1 output."""
    B = np.zeros((N, N))
    for j in range(N):
        hi = min(N, j + b + 1)
        B[j:hi, j] = A[j:hi, j]
    return np.tril(B) + np.tril(B, -1).T


def sb2st(B, b):
    """Bulge chasing on the full symmetric copy B (N x N).  Returns d, e, V2, tau2."""
    B = B.copy()
    N = B.shape[0]
    TM = (N + b - 1) // b + 1
    V2 = np.zeros((N, N))
    tau2 = np.zeros((N, TM))
    for s in range(N - 2):
        r = s + 1
        L = min(b, N - r)
        v, tau, beta = house(B[r:r + L, s].copy())
        B[r:r + L, s] = 0.0; B[r, s] = beta
        B[s, r:r + L] = B[r:r + L, s]
        t = 0
        while True:
            V2[r:r + L, s] = v
            tau2[s, t] = tau
            # (a) D <- H D H
            D = B[r:r + L, r:r + L]
            pvec = tau * (D @ v)
            alpha = 0.5 * tau * (pvec @ v)
            w = pvec - alpha * v
            D -= np.outer(v, w) + np.outer(w, v)
            r2 = r + L
            L2 = min(b, N - r2)
            if L2 <= 0:
                break
            # (b) E <- E H ; new reflector from E[:, 0]; E <- H' E
            E = B[r2:r2 + L2, r:r + L]
            y = tau * (E @ v)
            E -= np.outer(y, v)
            v2, tau_n, beta = house(E[:, 0].copy())
            E[:, 0] = 0.0; E[0, 0] = beta
            sj = v2 @ E[:, 1:]
            E[:, 1:] -= tau_n * np.outer(v2, sj)
            B[r:r + L, r2:r2 + L2] = E.T
            r, L, v, tau = r2, L2, v2, tau_n
            t += 1
    d = np.diag(B).copy()
    e = np.diag(B, -1).copy()
    return d, e, V2, tau2


def apply_q2(Z, V2, tau2, b, g):
    """Z <- Q2 Z in the kernel's order (groups of g sweeps, last group first)."""
    N = Z.shape[0]
    nsweeps = N - 2
    ngroups = (nsweeps + g - 1) // g
    for G in range(ngroups - 1, -1, -1):
        s0 = G * g
        tmax = (N - 1 - (s0 + 1) + b) // b + 1
        for t in range(tmax):
            for s in range(min(s0 + g, nsweeps) - 1, s0 - 1, -1):
                r = s + 1 + t * b
                if r >= N:
                    continue
                L = min(b, N - r)
                tau = tau2[s, t]
                if tau == 0.0:
                    continue
                v = V2[r:r + L, s]
                Z[r:r + L] -= tau * np.outer(v, v @ Z[r:r + L])
    return Z


def apply_q1(Z, A, tau1, b):
    """Z <- Q1 Z, block reflectors last to first; reflector q has its unit at row q + b."""
    n = A.shape[0]
    NB = n // b
    for p in range(NB - 2, -1, -1):
        m0 = b * (p + 1)
        m = n - m0
        V = np.zeros((m, b))
        for c in range(b):
            if c < m:
                V[c, c] = 1.0
                V[c + 1:, c] = A[m0 + c + 1:, b * p + c]
        T = larft(V, tau1[b * p:b * p + b])
        Z[m0:] -= V @ (T @ (V.T @ Z[m0:]))
    return Z


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    g = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    rng = np.random.default_rng(1)
    A0 = rng.standard_normal((N, N)); A0 = A0 + A0.T
    Npad = (N + b - 1) // b * b
    Ap = np.zeros((Npad, Npad)); Ap[:N, :N] = A0
    A1, tau1 = sy2sb(Ap, b)
    B = band_of(A1, N, b)
    w0 = np.linalg.eigvalsh(A0)
    print("stage 1: band eigenvalues vs dense   %.2e" % np.abs(np.linalg.eigvalsh(B) - w0).max())
    d, e, V2, tau2 = sb2st(B, b)
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    wt, Zt = np.linalg.eigh(Tm)
    print("stage 2: tridiagonal eigenvalues     %.2e" % np.abs(wt - w0).max())
    Z = apply_q2(Zt.copy(), V2, tau2, b, g)
    print("Q2: || B Z - Z w ||                  %.2e" % np.abs(B @ Z - Z * wt).max())
    Zp = np.zeros((Npad, N)); Zp[:N] = Z
    Zp = apply_q1(Zp, A1, tau1, b)
    Zf = Zp[:N]
    print("Q1: || A Z - Z w ||                  %.2e" % np.abs(A0 @ Zf - Zf * wt).max())
    print("orthogonality                        %.2e" % np.abs(Zf.T @ Zf - np.eye(N)).max())
    print("padding rows of Z                    %.2e" % np.abs(Zp[N:]).max() if Npad > N else "no padding")


if __name__ == "__main__":
    main()
