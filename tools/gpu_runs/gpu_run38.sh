#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest38.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest38.log
tail -n 4 gpurun_out/r02_pytest38.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench38.json 2> gpurun_out/r02_bench38.err
echo "bench rc=$?" >> gpurun_out/r02_bench38.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench38.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","summary") if k in d}); print(d["roofline"]["frac"], d["roofline"]["launch_ms"])
PY
tail -n 3 gpurun_out/r02_bench38.err
ARCB200_SY2SB_STREAMS=1 timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_c3_launches_final2.csv python tools/prof_c3.py 296 > gpurun_out/r02_ncu38.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_c3_launches_final2.csv > gpurun_out/r02_c3_launches_final2_summary.txt; head -n 24 gpurun_out/r02_c3_launches_final2_summary.txt
ARCB200_SY2SB_STREAMS=1 timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_c1_launches_twostage.csv python tools/prof_small.py c1 > gpurun_out/r02_ncu38b.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_c1_launches_twostage.csv > gpurun_out/r02_c1_launches_twostage_summary.txt; head -n 16 gpurun_out/r02_c1_launches_twostage_summary.txt
