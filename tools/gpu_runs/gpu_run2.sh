#!/bin/bash
# round-2 GPU pass 2 (2 GPUs): NCCL test, C4 triplet parity, the driver's multi-GPU bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -k "two_gpus or c4_sr88 or angle_pairs" > gpurun_out/r02_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
echo "bench2 rc=$?" >> gpurun_out/r02_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
echo "ref rc=$?" >> gpurun_out/r02_bench_ref.err
tail -n 4 gpurun_out/r02_pytest2.log gpurun_out/r02_bench_2gpu.err gpurun_out/r02_bench_ref.err
