#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest27.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest27.log
tail -n 6 gpurun_out/r02_pytest27.log
timeout -k 10 300 python tools/ab_sy2sb.py 2363 296 > gpurun_out/r02_ab27_c3.json 2>> gpurun_out/r02_ab27.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_ab27_c3.json")); print({k:(round(v["ms"],1) if isinstance(v,dict) else v) for k,v in d.items()})
PY
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench27.json 2> gpurun_out/r02_bench27.err
echo "bench rc=$?" >> gpurun_out/r02_bench27.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench27.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","summary") if k in d}); print(d["roofline"]["frac"], d["roofline"]["launch_ms"])
PY
tail -n 3 gpurun_out/r02_bench27.err
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_sy2sb_perphase_c3_launches.csv python tools/prof_sy2sb_once.py 2363 148 > gpurun_out/r02_ncu27.log 2>&1
tail -n 2 gpurun_out/r02_ncu27.log
