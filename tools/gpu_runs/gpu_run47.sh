#!/bin/bash
# Round 2, remaining 3.8 GPU-minutes: the driver's torchrun command on 2 GPUs with the final code (vector defineBasis,
# sharded radial tables through _compute_elements), C5 skipped and small CPU samples to fit the budget.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 5 95 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 2 --warmup 3 --c5-points 0 --cpu-pair-points 1 --cpu-stark-fields 10 --cpu-radial-pairs 10 \
    > gpurun_out/r02_final_bench_2gpu_short.json 2> gpurun_out/r02_final_bench_2gpu_short.err
echo "bench 2gpu rc=$?"; tail -c 1200 gpurun_out/r02_final_bench_2gpu_short.json; tail -n 5 gpurun_out/r02_final_bench_2gpu_short.err
