#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
timeout -k 5 200 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 100 -k "gemm" > gpurun_out/gemm.log 2>&1; el gemm_tests $?
timeout -k 5 300 python -m pytest tests/test_train_gpu.py -q -x --timeout 200 -k "bptt" > gpurun_out/bptt.log 2>&1; el bptt $?
timeout -k 5 60 python scripts/tc_dissect.py > gpurun_out/tc_dissect.txt 2>&1; el dissect $?
timeout -k 5 60 python scripts/tc_gemm_bench.py > gpurun_out/tc_bench.txt 2>&1; el tc_bench $?
timeout -k 5 90 python scripts/scan_timing.py > gpurun_out/scan_timing.txt 2>&1; el scan_timing $?
timeout -k 5 90 python scripts/section_times.py > gpurun_out/section_times.txt 2>&1; el sections $?
timeout -k 5 120 env DD_BENCH_VERBOSE=1 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; el bench $?
tail -n 6 gpurun_out/gemm.log; tail -n 15 gpurun_out/bptt.log | cut -c1-300; cat gpurun_out/tc_dissect.txt gpurun_out/tc_bench.txt gpurun_out/scan_timing.txt gpurun_out/section_times.txt; cut -c1-330 gpurun_out/bench.json
