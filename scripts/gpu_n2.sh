#!/bin/bash
# 2-GPU data-parallel bench exactly as the driver launches it (torchrun, one rank per GPU, NCCL).
mkdir -p gpurun_out
t0=$(date +%s)
nvidia-smi -L > gpurun_out/gpus_n2.txt 2>&1
timeout -k 5 170 env DD_BENCH_VERBOSE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench_n2 rc=$? t=$(( $(date +%s) - t0 ))s"
cat gpurun_out/bench_n2.json | cut -c1-700; tail -n 15 gpurun_out/bench_n2.err | cut -c1-300
