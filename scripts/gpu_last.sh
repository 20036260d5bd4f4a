#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout -k 5 100 python -m pytest tests -q -m gpu --timeout 60 -x > gpurun_out/tests_last.log 2>&1; echo "tests rc=$? t=$(( $(date +%s) - t0 ))s"
tail -n 6 gpurun_out/tests_last.log | cut -c1-300
timeout -k 5 40 python bench.py --no-cpu > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))s"
cut -c1-300 gpurun_out/bench_last.json
