#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
: > gpurun_out/r02_ab20_multi.json
for cfg in "2363 250" "1100 60"; do
  timeout -k 10 300 python tools/ab_multi.py $cfg >> gpurun_out/r02_ab20_multi.json 2>> gpurun_out/r02_ab20.err
  echo "multi $cfg rc=$?" >> gpurun_out/r02_ab20.err
done
timeout -k 10 300 python tools/ab_sy2sb.py 2363 296 > gpurun_out/r02_ab20_c3.json 2>> gpurun_out/r02_ab20.err
echo "ab rc=$?" >> gpurun_out/r02_ab20.err
timeout -k 10 300 python tools/ab_sy2sb.py 1932 148 > gpurun_out/r02_ab20_c4.json 2>> gpurun_out/r02_ab20.err
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest20.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest20.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench20.json 2> gpurun_out/r02_bench20.err
echo "bench rc=$?" >> gpurun_out/r02_bench20.err
cat gpurun_out/r02_ab20_multi.json
python - <<PY
import json
for f in ("gpurun_out/r02_ab20_c3.json","gpurun_out/r02_ab20_c4.json"):
    d=json.load(open(f)); print({k:(round(v["ms"],1) if isinstance(v,dict) else v) for k,v in d.items()})
d=json.loads(open("gpurun_out/r02_bench20.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","roofline","summary") if k in d})
PY
tail -n 5 gpurun_out/r02_ab20.err; tail -n 4 gpurun_out/r02_pytest20.log; tail -n 3 gpurun_out/r02_bench20.err
