#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
: > gpurun_out/r02_ab29.json
for ns in 1 2 4 8; do
  for cfg in "2363 296" "2363 125" "651 286" "410 300"; do
    echo "streams=$ns" >> gpurun_out/r02_ab29.json
    ARCB200_TWO_STAGE=1 ARCB200_SY2SB_STREAMS=$ns AB_ONLY=multi timeout -k 10 200 python tools/ab_multi.py $cfg >> gpurun_out/r02_ab29.json 2>> gpurun_out/r02_ab29.err
  done
done
python - <<PY
import json
ns=None
for l in open("gpurun_out/r02_ab29.json"):
    l=l.strip()
    if l.startswith("streams"): ns=l; continue
    d=json.loads(l); print(ns, d["N"], d["batch"], round(d["many_ctas_per_matrix"]["ms"],2))
PY
tail -n 3 gpurun_out/r02_ab29.err
