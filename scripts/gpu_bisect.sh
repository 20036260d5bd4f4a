#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
timeout -k 5 60 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 30 -k "gemm" > gpurun_out/gemm.log 2>&1; rc=$?; el gemm_tests $rc
if [ $rc -ne 0 ]; then tail -5 gpurun_out/gemm.log; exit 1; fi
for f in 0 1 2 3; do
  timeout -k 5 60 env DD_BPTT_FLAGS=$f python -m pytest tests/test_train_gpu.py -q -x --timeout 40 -k "bptt" > gpurun_out/bptt_$f.log 2>&1; rc=$?; el bptt_flags$f $rc
  tail -n 4 gpurun_out/bptt_$f.log | cut -c1-200
  if [ $rc -eq 124 ]; then exit 1; fi
done
for f in 0 3; do
  timeout -k 5 60 env DD_BPTT_FLAGS=$f python scripts/scan_timing.py > gpurun_out/scan_timing_$f.txt 2>&1; el scan_timing$f $?
  grep bwd gpurun_out/scan_timing_$f.txt
  timeout -k 5 60 env DD_BPTT_FLAGS=$f python scripts/section_times.py > gpurun_out/section_times_$f.txt 2>&1; el sections$f $?
  grep -E "wm_backward|sum" gpurun_out/section_times_$f.txt
done
