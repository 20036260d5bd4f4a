#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_machine_gpu.py -x -q --timeout 60 -s > gpurun_out/machine_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|FAILED|machine debug|Error" gpurun_out/machine_tests.log | cut -c1-250 | tail -6
for cfg in "DD_X=1" "DD_MACHINE_BWD=1"; do
echo "== $cfg"
timeout -k 5 120 env $cfg python scripts/section_times.py 2>&1 | tail -n 12
done
timeout -k 5 60 python scripts/machine_prof.py > gpurun_out/machine_prof.txt 2>&1; grep -B6 "^LN chain" gpurun_out/machine_prof.txt | cut -c1-200
