#!/bin/bash
# layer-machine bring-up: unit tests under a short timeout (a protocol bug traps after 2 s, a
# real hang is cut at 200 s), then timings / wait profile
mkdir -p gpurun_out
if [ "$1" != "prof" ]; then
timeout -k 5 200 python -m pytest tests/test_machine_gpu.py -q --timeout 60 -s > gpurun_out/machine_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|FAILED|3xTF32|machine debug" gpurun_out/machine_tests.log | cut -c1-300 | tail -40
fi
timeout -k 5 120 python scripts/machine_prof.py > gpurun_out/machine_prof.txt 2>&1
echo "prof rc=$?"; cat gpurun_out/machine_prof.txt | cut -c1-330
if [ "$1" = "bench" ]; then
timeout -k 5 120 python scripts/machine_bench.py > gpurun_out/machine_bench.txt 2>&1
echo "bench rc=$?"; cat gpurun_out/machine_bench.txt | cut -c1-250
fi
