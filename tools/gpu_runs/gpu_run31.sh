#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest31.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest31.log
tail -n 6 gpurun_out/r02_pytest31.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --skip-sub > gpurun_out/r02_bench31.json 2> gpurun_out/r02_bench31.err
echo "bench rc=$?" >> gpurun_out/r02_bench31.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench31.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches") if k in d}); print(d["roofline"]["frac"], d["roofline"]["launch_ms"])
PY
tail -n 3 gpurun_out/r02_bench31.err
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_c3_launches_perphase.csv python bench.py --steps 1 --warmup 1 --skip-sub > gpurun_out/r02_ncu31.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_c3_launches_perphase.csv > gpurun_out/r02_c3_launches_perphase_summary.txt; head -n 24 gpurun_out/r02_c3_launches_perphase_summary.txt
