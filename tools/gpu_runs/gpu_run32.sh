#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest32.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest32.log
tail -n 4 gpurun_out/r02_pytest32.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench32.json 2> gpurun_out/r02_bench32.err
echo "bench rc=$?" >> gpurun_out/r02_bench32.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench32.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","summary") if k in d}); print(d["roofline"]["frac"], d["roofline"]["launch_ms"])
PY
tail -n 3 gpurun_out/r02_bench32.err
ARCB200_SY2SB_STREAMS=1 timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_c3_launches_final.csv python tools/prof_c3.py 296 > gpurun_out/r02_ncu32.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_c3_launches_final.csv > gpurun_out/r02_c3_launches_final_summary.txt; head -n 24 gpurun_out/r02_c3_launches_final_summary.txt
