"""A/B of the half-batch pipeline of arcb200_pair_sweep (ARCB200_PIPELINE=0/1) on C3 shards that do not fill
the GPU (one sweep of npts points, batch = npts).   python tools/ab_pipeline.py npts [npts ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from arc_b200 import sweep
calc = bench.build_c3()
ops = calc.device_operators()
for npts in [int(a) for a in sys.argv[1:]]:
    R = np.linspace(10.0, 1.0, npts)
    out = {"N": 2363, "k": 250, "points": npts}
    ys = {}
    for tag, env in (("serial", "0"), ("pipelined", "1")):
        os.environ["ARCB200_PIPELINE"] = env
        for _ in range(2):
            r = sweep.pair_sweep(ops, R, 250, 0.0, calc.originalPairStateIndex, topk=4, to_host=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = sweep.pair_sweep(ops, R, 250, 0.0, calc.originalPairStateIndex, topk=4, to_host=False)
        e1.record(); torch.cuda.synchronize()
        out[tag] = {"ms": e0.elapsed_time(e1) / 3, "solves_per_s": npts / (e0.elapsed_time(e1) / 3e3)}
        ys[tag] = r.y.cpu().numpy()
    out["y_maxdiff"] = float(np.abs(ys["serial"] - ys["pipelined"]).max())
    print(json.dumps(out), flush=True)
os.environ.pop("ARCB200_PIPELINE", None)
