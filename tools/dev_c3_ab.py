"""A/B timing of the C3 pair sweep under the tuning knobs (ARCB200_LANES, ARCB200_BATCH,
ARCB200_SYTRD_LEAN).  Usage: python tools/dev_c3_ab.py [points] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import arc_b200
from arc_b200 import sweep
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 296
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
calc = arc_b200.PairStateInteractions(arc_b200.Rubidium(), 60, 0, .5, 60, 0, .5, .5, .5)
calc.defineBasis(0., 0., 5, 5, 25e9)
ops = calc.device_operators()
R = np.sort(np.resize(np.linspace(1, 10, 1000), npts))[::-1].copy()
def step():
    return sweep.pair_sweep(ops, R, 250, 0.0, calc.originalPairStateIndex, topk=4, to_host=False)
r0 = step(); torch.cuda.synchronize()
ref = os.environ.get("AB_REF")
y = r0.y.cpu().numpy()
if ref:
    if os.path.exists(ref):
        y0 = np.load(ref); print("max |dy| vs ref: %.3e (scale %.3e)" % (np.abs(y - y0).max(), np.abs(y0).max()))
    else:
        np.save(ref, y)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print("knobs lanes=%s batch=%s lean=%s: %d points %.1f ms/step -> %.1f solves/s" % (
    os.environ.get("ARCB200_LANES"), os.environ.get("ARCB200_BATCH"), os.environ.get("ARCB200_SYTRD_LEAN"), npts, ms, npts / ms * 1e3))
