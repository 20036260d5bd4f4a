#!/bin/bash
# Runs on the GPU box under gpurun: parity tests, smoke, bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 200 python scripts/tc_gemm_bench.py > gpurun_out/tc_bench.log 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu > gpurun_out/kernels.log 2>&1
echo "kernels rc=$?" | tee -a gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_train_gpu.py -q -m gpu > gpurun_out/train.log 2>&1
echo "train rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_agent_gpu.py -q -m gpu > gpurun_out/agent.log 2>&1
echo "agent rc=$?" | tee -a gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/tc_bench.log
for f in kernels train agent smoke; do echo "== $f"; tail -n 6 gpurun_out/$f.log; done
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
