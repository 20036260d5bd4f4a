#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
timeout -k 5 90 python -m pytest tests/test_train_gpu.py -q -x --timeout 60 -k "bptt or fused_observe" > gpurun_out/scan_tests.log 2>&1; rc=$?; el scan_tests $rc
tail -n 4 gpurun_out/scan_tests.log | cut -c1-250
if [ $rc -eq 124 ]; then exit 1; fi
timeout -k 5 60 python scripts/scan_timing.py > gpurun_out/scan_timing.txt 2>&1; el scan_timing $?
cat gpurun_out/scan_timing.txt
timeout -k 5 60 python scripts/section_times.py > gpurun_out/section_times.txt 2>&1; el sections $?
cat gpurun_out/section_times.txt
timeout -k 5 100 env DD_BENCH_VERBOSE=1 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; el bench $?
cut -c1-330 gpurun_out/bench.json
