#!/bin/bash
# Full validation on the GPU box: every -m gpu test, smoke, bench (both arms), GEMM dissection, phase times.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; uptime >> gpurun_out/gpu.txt
timeout -k 5 200 env DD_BENCH_VERBOSE=1 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; el bench $?
timeout -k 5 300 env DD_BENCH_VERBOSE=1 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; el bench_ref $?
timeout -k 5 60 python scripts/tc_dissect.py > gpurun_out/tc_dissect.txt 2>&1; el dissect $?
timeout -k 5 60 python scripts/tc_dissect2.py >> gpurun_out/tc_dissect.txt 2>&1; el dissect2 $?
timeout -k 5 90 python scripts/section_times.py > gpurun_out/section_times.txt 2>&1; el sections $?
timeout -k 5 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; el smoke $?
timeout -k 5 900 python -m pytest tests -q -m gpu --timeout 400 --durations=12 > gpurun_out/tests.log 2>&1; el tests $?
tail -n 30 gpurun_out/tests.log; tail -n 2 gpurun_out/smoke.log
cat gpurun_out/bench.json; tail -n 12 gpurun_out/bench.err; cat gpurun_out/bench_ref.json; tail -n 5 gpurun_out/bench_ref.err
cat gpurun_out/tc_dissect.txt gpurun_out/section_times.txt
