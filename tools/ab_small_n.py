"""A/B of the tridiagonalisation path on small bases (C1 N=410, C4 singlet N=651): one-stage cluster
kernel (default below N=1024) against the two-stage path (ARCB200_TWO_STAGE=1).  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import arc_b200
from arc_b200 import sweep
out = {}
cases = []
c1 = arc_b200.StarkMap(arc_b200.Rubidium()); c1.defineBasis(28, 0, 0.5, 0.5, 23, 32, 20)
cases.append(("c1_N410", c1, np.linspace(0, 600, 600)))
c4 = arc_b200.StarkMap(arc_b200.Strontium88()); c4.defineBasis(60, 0, 0, 0, 50, 70, 30, s=0)
cases.append(("c4_singlet_N651", c4, np.linspace(0, 200, 2000)))
for name, calc, F in cases:
    h0 = np.diag(calc.mat1).copy()
    res = {}
    ys = {}
    for tag, env in (("one_stage", "0"), ("two_stage", "1")):
        os.environ["ARCB200_TWO_STAGE"] = env
        for _ in range(2):
            r = sweep.stark_sweep(h0, calc.mat2, F, calc.indexOfCoupledState, to_host=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = sweep.stark_sweep(h0, calc.mat2, F, calc.indexOfCoupledState, to_host=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[tag] = {"ms_per_sweep": ms, "solves_per_s": len(F) / (ms * 1e-3)}
        ys[tag] = r.y.cpu().numpy()
    res["y_maxdiff_rel"] = float(np.abs(ys["one_stage"] - ys["two_stage"]).max() / np.abs(ys["one_stage"]).max())
    out[name] = res
os.environ.pop("ARCB200_TWO_STAGE", None)
print(json.dumps(out))
