"""Short driver for ncu: one C3 pair sweep batch (N=2363, k=250) through the C-ABI."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from arc_b200 import sweep
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 36
calc = bench.build_c3()
ops = calc.device_operators()
R = np.linspace(10.0, 1.0, npts)
for it in range(2):
    r = sweep.pair_sweep(ops, R, 250, 0.0, calc.originalPairStateIndex, topk=4, to_host=False)
torch.cuda.synchronize()
print("launches per sweep", r.kernel_launches)
