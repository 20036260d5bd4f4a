#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_machine_gpu.py -q --timeout 60 > gpurun_out/machine_tests.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/machine_tests.log | cut -c1-200
timeout -k 5 60 python scripts/machine_prof.py > gpurun_out/machine_prof.txt 2>&1; grep -B7 "^LN chain" gpurun_out/machine_prof.txt | cut -c1-200
timeout -k 5 120 python scripts/machine_bench.py > gpurun_out/machine_bench.txt 2>&1; cat gpurun_out/machine_bench.txt | cut -c1-230
timeout -k 5 120 python scripts/section_times.py 2>&1 | tail -n 12
