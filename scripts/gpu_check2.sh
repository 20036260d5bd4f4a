#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
timeout -k 5 240 python -m pytest tests -q -m gpu --timeout 100 -x > gpurun_out/tests.log 2>&1; rc=$?; el tests $rc
tail -n 12 gpurun_out/tests.log | cut -c1-300
if [ $rc -eq 124 ]; then exit 1; fi
timeout -k 5 60 python scripts/section_times.py > gpurun_out/section_times.txt 2>&1; el sections $?
cat gpurun_out/section_times.txt
timeout -k 5 100 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; el bench $?
cut -c1-330 gpurun_out/bench.json
timeout -k 5 100 env DD_GEMM_THIN=0 DD_LN_WARP_MAX=128 python bench.py --no-cpu > gpurun_out/bench_old.json 2> gpurun_out/bench_old.err; el bench_old $?
cut -c1-330 gpurun_out/bench_old.json
