#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:observe_fwd_kernel -s 1 -c 1 \
  -o gpurun_out/scan_prof -f python scripts/scan_timing.py > gpurun_out/ncu_scan.log 2>&1
echo "rc=$?"; tail -n 4 gpurun_out/ncu_scan.log
