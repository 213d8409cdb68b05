#!/bin/bash
# round-2 GPU pass 1: tests, smoke, bench, launch list, one full ncu capture of the band reduction
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r02_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke1.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02_smoke1.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
echo "bench rc=$?" >> gpurun_out/r02_bench1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c3_launches_base.csv python tools/prof_c3.py 148 > gpurun_out/r02_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sy2sb_kernel -c 1 -f -o gpurun_out/r02_sy2sb_base python tools/prof_c3.py 148 > gpurun_out/r02_ncu_full.log 2>&1
tail -n 5 gpurun_out/r02_pytest1.log gpurun_out/r02_smoke1.log gpurun_out/r02_bench1.err
