import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, json
import arc_b200
from arc_b200 import radial
from oracle import oracle
atom = arc_b200.Rubidium()
g = json.load(open("tests/golden/radial_golden.json"))["atoms"]["Rubidium"]
worst = []
rows = [(r, 1) for r in g["dipole"]] + [(r, 2) for r in g["quadrupole"]]
atom.precompute_radial(dipole_pairs=[tuple(r[:6]) for r in g["dipole"]], quadrupole_pairs=[tuple(r[:6]) for r in g["quadrupole"]])
for r, p in rows:
    v = atom.getRadialMatrixElement(*r[:6]) if p == 1 else atom.getQuadrupoleMatrixElement(*r[:6])
    worst.append((abs(v - r[6]) / abs(r[6]), r, p, v))
worst.sort(key=lambda t: -t[0])
for w in worst[:6]:
    print("rel %.2e" % w[0], w[1], w[2], w[3])
print("median rel %.2e" % np.median([w[0] for w in worst]))
# detailed look at the worst
rel, r, p, v = worst[0]
key = atom._ordered_key(*r[:6])
sa, sb = (key[0], key[1], key[2] / 2), (key[3], key[4], key[5] / 2)
recs = np.array([atom._state_record(*sa), atom._state_record(*sb)])
rb = radial.RadialBatch(recs)
def args(rec):
    return [float(rec[k]) for k in ("innerLimit", "outerLimit", "step", "init1", "init2")] + \
           [int(rec["l"]), float(rec["s"]), float(rec["j"]), float(rec["stateEnergy"]),
            float(rec["alphaC"]), float(rec["alpha"]), int(rec["Z"])] + \
           [float(rec[k]) for k in ("a1", "a2", "a3", "a4", "rc", "mu")]
W = []
for i in range(2):
    po, ro, dpo = oracle.numerov(*args(recs[i]))
    po = po / np.sqrt(oracle.trapezoid(po ** 2, ro))
    rg, pg = rb.wavefunction(i)
    print("state", i, (sa, sb)[i], "L", len(ro), "dp oracle", dpo, "gpu", int(rb.divergence_points()[i]),
          "max|dpsi|/max|psi| %.2e" % (np.abs(pg - po).max() / np.abs(po).max()),
          "r equal", np.array_equal(rg, ro))
    W.append((ro, po, rg, pg))
up = min(len(W[0][0]), len(W[1][0]))
wgt = (lambda x: x) if p == 1 else (lambda x: x * x)
f_o = W[0][1][:up] * W[1][1][:up] * wgt(W[0][0][:up])
f_g = W[0][3][:up] * W[1][3][:up] * wgt(W[0][2][:up])
print("oracle integral", oracle.trapezoid(f_o, W[0][0][:up]), "host-trapz of gpu psi", oracle.trapezoid(f_g, W[0][2][:up]),
      "gpu kernel", float(rb.integrals(np.array([[0, 1, p]], dtype=np.int32))[0]), "golden", r[6])
print("cancellation scale int|f|", oracle.trapezoid(np.abs(f_o), W[0][0][:up]))
