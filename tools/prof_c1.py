"""Short driver for ncu: one C1 Stark sweep (N=410, 600 fields) through the C-ABI."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import arc_b200
from arc_b200 import sweep
calc = arc_b200.StarkMap(arc_b200.Rubidium())
calc.defineBasis(28, 0, 0.5, 0.5, 23, 32, 20)
F = np.linspace(0, 600, 600)
h0 = np.diag(calc.mat1).copy()
for it in range(2):
    r = sweep.stark_sweep(h0, calc.mat2, F, calc.indexOfCoupledState, to_host=False)
torch.cuda.synchronize()
print("launches per sweep", r.kernel_launches)
