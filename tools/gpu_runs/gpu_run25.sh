#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
: > gpurun_out/r02_ab25.json
for cfg in "2363 250" "651 286" "410 300" "1027 3" "1100 20"; do
  timeout -k 10 200 python tools/ab_multi.py $cfg >> gpurun_out/r02_ab25.json 2>> gpurun_out/r02_ab25.err
done
timeout -k 10 300 python tools/ab_sy2sb.py 2363 296 > gpurun_out/r02_ab25_c3.json 2>> gpurun_out/r02_ab25.err
cat gpurun_out/r02_ab25.json
python - <<PY
import json
d=json.load(open("gpurun_out/r02_ab25_c3.json")); print({k:(round(v["ms"],1) if isinstance(v,dict) else v) for k,v in d.items()})
PY
timeout -k 10 600 python tools/ab_crossover.py 296 > gpurun_out/r02_crossover.json 2>> gpurun_out/r02_ab25.err
cat gpurun_out/r02_crossover.json
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest25.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest25.log
tail -n 4 gpurun_out/r02_pytest25.log; tail -n 5 gpurun_out/r02_ab25.err
