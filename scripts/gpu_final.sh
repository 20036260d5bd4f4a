#!/bin/bash
# Round-end style validation + ncu evidence in one box lease.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; uptime >> gpurun_out/gpu.txt
timeout -k 5 300 python -m pytest tests -q -m gpu --timeout 120 --durations=8 > gpurun_out/tests.log 2>&1; el tests $?
tail -n 14 gpurun_out/tests.log
timeout -k 5 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; el smoke $?
tail -n 1 gpurun_out/smoke.log
timeout -k 5 150 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; el bench $?
cat gpurun_out/bench.json
timeout -k 5 240 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; el bench_ref $?
cat gpurun_out/bench_ref.json
bash scripts/gpu_prof.sh
cat gpurun_out/summary.txt
