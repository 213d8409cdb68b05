"""One Stark sweep of a small basis for an ncu launch list: python tools/prof_small.py c1|c4 (ARCB200_TWO_STAGE=0/1 from the env)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import arc_b200
from arc_b200 import sweep
if sys.argv[1] == "c1":
    calc = arc_b200.StarkMap(arc_b200.Rubidium()); calc.defineBasis(28, 0, 0.5, 0.5, 23, 32, 20); F = np.linspace(0, 600, 600)
else:
    calc = arc_b200.StarkMap(arc_b200.Strontium88()); calc.defineBasis(60, 0, 0, 0, 50, 70, 30, s=0); F = np.linspace(0, 200, 2000)
h0 = np.diag(calc.mat1).copy()
r = sweep.stark_sweep(h0, calc.mat2, F, calc.indexOfCoupledState, to_host=False)
torch.cuda.synchronize()
