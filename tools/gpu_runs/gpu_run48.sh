#!/bin/bash
# Round 2, last minute of GPU budget: C1 sharded over 2 GPUs, per-call timings with the gather separated
# (the short 2-GPU bench showed diagonalise 0.115 s where the earlier run had 0.031 s).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
cat > /tmp/c1_2gpu.py <<PY
import os, sys, time, datetime
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=60))
import arc_b200
from arc_b200 import sweep
tg = [0.0]
orig = sweep.gather_rows
def timed_gather(local_t, n, group=sweep.WORLD):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = orig(local_t, n, group); torch.cuda.synchronize(); tg[0] += time.perf_counter() - t0; return r
sweep.gather_rows = timed_gather
atom = arc_b200.Rubidium(); atom.radial_group = sweep.WORLD
calc = arc_b200.StarkMap(atom); calc.defineBasis(28, 0, 0.5, 0.5, 23, 32, 20)
F = np.linspace(0, 600, 600)
for it in range(7):
    tg[0] = 0.0
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    calc.diagonalise(F, group=sweep.WORLD)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("rank %d call %d: diagonalise %.4f s (gathers %.4f s)" % (rank, it, dt, tg[0]), flush=True)
    if it == 3: time.sleep(0.5)
dist.destroy_process_group()
PY
timeout -k 5 40 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 /tmp/c1_2gpu.py > gpurun_out/r02_c1_2gpu_calls.log 2>&1
echo "rc=$?"; grep "call" gpurun_out/r02_c1_2gpu_calls.log | sort | head -20; tail -n 3 gpurun_out/r02_c1_2gpu_calls.log
