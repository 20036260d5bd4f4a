#!/bin/bash
# layer-machine bring-up: unit tests under a short timeout (a protocol bug traps after 2 s, a
# real hang is cut at 150 s), then the per-layer timings
mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests/test_machine_gpu.py -x -q --timeout 60 -s > gpurun_out/machine_tests.log 2>&1
echo "tests rc=$?"; tail -n 25 gpurun_out/machine_tests.log | cut -c1-300
timeout -k 5 120 python scripts/machine_bench.py > gpurun_out/machine_bench.txt 2>&1
echo "bench rc=$?"; cat gpurun_out/machine_bench.txt | cut -c1-250
