#!/bin/bash
# Round-end style check on the GPU box under gpurun: all -m gpu tests, smoke, bench (both arms).
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$(date +%s)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/tests.log 2>&1
echo "tests rc=$? t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$? t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$? t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt
if [ "$1" == "ref" ]; then
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench_ref rc=$? t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt
fi
tail -n 25 gpurun_out/tests.log; tail -n 3 gpurun_out/smoke.log
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
