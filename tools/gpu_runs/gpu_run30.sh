#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest30.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest30.log
tail -n 6 gpurun_out/r02_pytest30.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench30.json 2> gpurun_out/r02_bench30.err
echo "bench rc=$?" >> gpurun_out/r02_bench30.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench30.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","summary") if k in d}); print(d["roofline"]["frac"], d["roofline"]["launch_ms"])
PY
tail -n 3 gpurun_out/r02_bench30.err
timeout -k 10 600 python tools/dev_c5.py > gpurun_out/r02_c5_30.log 2>&1; tail -n 5 gpurun_out/r02_c5_30.log
