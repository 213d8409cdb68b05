#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python - <<PY
import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import arc_b200
from arc_b200 import sweep
calc = arc_b200.StarkMap(arc_b200.Strontium88()); calc.defineBasis(60, 0, 0, 0, 50, 70, 30, s=0)
F = np.linspace(0, 200, 2000)
for b in (286, 400, 500, 667):
    os.environ["ARCB200_BATCH"] = str(b)
    calc.diagonalise(F[:2*b])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    calc.diagonalise(F)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("c4 singlet batch", b, "wall %.3f s  %.0f solves/s" % (t1 - t0, 2000 / (t1 - t0)), "launches", calc.last_sweep.kernel_launches, flush=True)
os.environ.pop("ARCB200_BATCH")
c1 = arc_b200.StarkMap(arc_b200.Rubidium()); c1.defineBasis(28, 0, 0.5, 0.5, 23, 32, 20)
F1 = np.linspace(0, 600, 600)
for b in (300, 600):
    os.environ["ARCB200_BATCH"] = str(b)
    c1.diagonalise(F1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c1.diagonalise(F1)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("c1 batch", b, "wall %.4f s  %.0f solves/s" % (t1 - t0, 600 / (t1 - t0)), "launches", c1.last_sweep.kernel_launches, flush=True)
PY
echo "sb2st split on C5"; ARCB200_SB2ST_SPLIT=1 timeout -k 10 400 python tools/dev_c5.py 74 2>&1 | grep "diagonalise\|max" | tail -n 2
