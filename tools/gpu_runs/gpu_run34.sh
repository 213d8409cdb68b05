#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest34.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest34.log
tail -n 4 gpurun_out/r02_pytest34.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --skip-sub > gpurun_out/r02_bench34.json 2> gpurun_out/r02_bench34.err
echo "bench rc=$?" >> gpurun_out/r02_bench34.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench34.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches") if k in d}); print(d["roofline"]["frac"], d["roofline"]["launch_ms"])
PY
ARCB200_SY2SB_STREAMS=1 timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"q2_kernel" -c 2 --csv --log-file gpurun_out/r02_q2b_launches.csv python tools/prof_c3.py 296 > gpurun_out/r02_ncu34.log 2>&1
grep "q2_kernel" gpurun_out/r02_q2b_launches.csv | cut -c1-300
