#!/bin/bash
# the driver's multi-GPU command for N = $1
N=$1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "bench N=$N rc=$?" >> gpurun_out/r02_bench_${N}gpu.err
tail -n 3 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().split('\n')[-1])
    print(json.dumps(d['summary']))
except Exception as e:
    print('no json', e)
PY
