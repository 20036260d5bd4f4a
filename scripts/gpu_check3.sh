#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
timeout -k 5 80 python bench.py --no-cpu > gpurun_out/bench_thin1.json 2> gpurun_out/bench.err; el bench_thin1 $?
cut -c1-260 gpurun_out/bench_thin1.json
timeout -k 5 80 env DD_GEMM_THIN=0 python bench.py --no-cpu > gpurun_out/bench_thin0.json 2> gpurun_out/bench_old.err; el bench_thin0 $?
cut -c1-260 gpurun_out/bench_thin0.json
timeout -k 5 60 python scripts/section_times.py > gpurun_out/section_times.txt 2>&1; el sections $?
cat gpurun_out/section_times.txt
