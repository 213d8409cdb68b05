"""Short driver for ncu: a slice of the C2 radial table (Rb n=20..70, l<=20)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import arc_b200
from arc_b200 import radial
from dev_radial_bench import c2_workload
atom = arc_b200.Rubidium()
nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 70
states, pairs = c2_workload(atom, nmax=nmax)
recs = np.array([atom._state_record(*s) for s in states])
for it in range(2):
    rb = radial.RadialBatch(recs)
    out = rb.integrals_device(pairs)
torch.cuda.synchronize()
lens = rb.lengths.astype(np.int64)
print("states", len(states), "pairs", len(pairs), "points", int(lens.sum()),
      "point-pairs", int(np.minimum(lens[pairs[:, 0]], lens[pairs[:, 1]]).sum()))
