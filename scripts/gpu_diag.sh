#!/bin/bash
# Diagnostic pass on the GPU box: kernel unit tests, bench with/without the persistent GEMM and
# PDL (no CPU leg), PDL micro-benchmark, then the train/agent parity tests that are new.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; uptime >> gpurun_out/gpu.txt
timeout -k 5 240 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 60 > gpurun_out/kernels.log 2>&1; el kernels $?
B="python bench.py --steps 10 --warmup 3 --no-cpu"
timeout -k 5 150 env DD_BENCH_VERBOSE=1 DD_GEMM_TC_PERSISTENT=0 $B > gpurun_out/bench_np.json 2> gpurun_out/bench_np.err; rc=$?; el bench_nopersist $rc
if [ $rc -eq 0 ]; then
  timeout -k 5 150 env DD_BENCH_VERBOSE=1 $B > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; el bench_persist $?
  (cd scripts/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o pdl_chain pdl_chain.cu) && ./scripts/micro/pdl_chain > gpurun_out/pdl_chain.txt 2>&1; el pdl_chain $?
  timeout -k 5 150 env DD_BENCH_VERBOSE=1 DD_GEMM_TC_PERSISTENT=0 DD_PDL=1 $B > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; el bench_pdl $?
fi
timeout -k 5 500 python -m pytest tests/test_train_gpu.py tests/test_agent_gpu.py -q -x --timeout 150 -k "not a1" --durations=10 > gpurun_out/new_tests.log 2>&1; el new_tests $?
tail -n 15 gpurun_out/kernels.log; tail -n 30 gpurun_out/new_tests.log
for f in np p pdl; do echo "== bench_$f"; cat gpurun_out/bench_$f.json; tail -n 25 gpurun_out/bench_$f.err; done
cat gpurun_out/pdl_chain.txt
