"""C6 extraction on the C3 level diagram (BASELINE configs[2]) and a check of the tracked level
against numpy on the same matrices at a few distances."""
import sys, os, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import arc_b200
calc = arc_b200.PairStateInteractions(arc_b200.Rubidium(), 60, 0, .5, 60, 0, .5, .5, .5)
calc.defineBasis(0., 0., 5, 5, 25e9)
R = np.linspace(1.0, 10.0, 296)
calc.diagonalise(R, 250)
for rng in ((4.0, 9.0), (3.0, 9.0), (2.0, 6.0), (4.0, 9.5)):
    with contextlib.redirect_stdout(io.StringIO()):
        c6 = calc.getC6fromLevelDiagram(*rng)
    print("C6 fit over", rng, "=", c6)
for br in (10, 100, 200, 290):
    hl = np.asarray(calc.highlight[br]); i = int(np.argmax(hl))
    r = calc.r[br]
    H = calc.matDiagonal.toarray() + calc.matR[0].toarray() / (r * 1e-6) ** 3
    w, v = np.linalg.eigh(H)
    j = int(np.argmax(v[calc.originalPairStateIndex] ** 2))
    print("R=%.3f tracked E=%.6e hl=%.4f | numpy E=%.6e hl=%.4f | C6/R^6 with 137.5: %.6e" % (
        r, calc.y[br][i], hl[i], w[j], v[calc.originalPairStateIndex, j] ** 2, -137.5 / r ** 6))
