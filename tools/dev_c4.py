"""BASELINE configs[3]-class run: Sr88 divalent Stark maps, singlet (N=651) and triplet (N=1932)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import arc_b200
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 148
for (j, s) in ((0, 0), (1, 1)):
    t0 = time.time()
    calc = arc_b200.StarkMap(arc_b200.Strontium88())
    calc.defineBasis(60, 0, j, 0, 50, 70, 30, s=s)
    N = len(calc.basisStates)
    t1 = time.time()
    F = np.linspace(0, 200, nF)
    calc.diagonalise(F); torch.cuda.synchronize()
    t2 = time.time()
    calc.diagonalise(F); torch.cuda.synchronize()
    t3 = time.time()
    i = nF // 2
    w = np.linalg.eigvalsh(calc.mat1 + calc.mat2 * F[i])
    print("C4 s=%d: N=%d defineBasis %.1f s, diagonalise %d fields %.2f s (%.1f solves/s), max|dy|/scale %.1e" % (
        s, N, t1 - t0, nF, t3 - t2, nF / (t3 - t2), np.abs(calc.y[i] - w).max() / np.abs(w).max()), flush=True)
