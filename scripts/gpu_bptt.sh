#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
timeout -k 5 300 python -m pytest tests/test_train_gpu.py -q -x --timeout 200 -k "bptt" > gpurun_out/bptt.log 2>&1; el bptt $?
timeout -k 5 400 python -m pytest tests/test_train_gpu.py -q -x --timeout 300 -k "not bptt and not fused_observe" > gpurun_out/train.log 2>&1; el train $?
timeout -k 5 90 python scripts/section_times.py > gpurun_out/section_times.txt 2>&1; el sections $?
timeout -k 5 120 env DD_BENCH_VERBOSE=1 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; el bench $?
tail -n 30 gpurun_out/bptt.log; tail -n 12 gpurun_out/train.log; cat gpurun_out/section_times.txt; cat gpurun_out/bench.json | cut -c1-400
