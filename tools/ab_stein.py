"""A/B of stage 2 for pair sweeps on the C3 workload: ARCB200_STEIN=0 (divide & conquer, round 1)
against the default bisection + inverse iteration.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from arc_b200 import sweep
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 296
calc = bench.build_c3()
ops = calc.device_operators()
R = np.linspace(10.0, 1.0, npts)
out = {"points": npts}
ys = {}
for name, val in (("dc", "0"), ("stein", "")):
    if val:
        os.environ["ARCB200_STEIN"] = val
    else:
        os.environ.pop("ARCB200_STEIN", None)
    for _ in range(2):
        r = sweep.pair_sweep(ops, R, 250, 0.0, calc.originalPairStateIndex, topk=4, to_host=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = sweep.pair_sweep(ops, R, 250, 0.0, calc.originalPairStateIndex, topk=4, to_host=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ys[name] = (r.y.cpu().numpy(), r.hl.cpu().numpy())
    out[name] = {"ms_per_sweep": ms, "solves_per_s": npts / (ms * 1e-3), "launches": r.kernel_launches}
out["y_maxdiff_rel"] = float(np.abs(ys["dc"][0] - ys["stein"][0]).max() / np.abs(ys["dc"][0]).max())
out["hl_maxdiff_tracked"] = float(np.abs(ys["dc"][1].max(axis=1) - ys["stein"][1].max(axis=1)).max())
out["speedup"] = out["dc"]["ms_per_sweep"] / out["stein"]["ms_per_sweep"]
print(json.dumps(out))
