#!/bin/bash
# Round-2 first call: correctness of the experimental kernels written blind at the end of r01
# (each under a short timeout -- a hang must not eat the GPU budget), then A/B of each switch on the
# a1 bench line.  Stops at the first hang.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
export DD_TEST_EXPERIMENTAL=1
for k in ln_vec_experimental "gemm_thin and 2" staged; do
  timeout -k 5 90 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 30 -k "$k" > "gpurun_out/exp_${k// /_}.log" 2>&1; rc=$?; el "test[$k]" $rc
  tail -n 3 "gpurun_out/exp_${k// /_}.log" | cut -c1-200
  if [ $rc -eq 124 ]; then echo "hang in $k: stopping"; exit 1; fi
done
unset DD_TEST_EXPERIMENTAL
B="python bench.py --no-cpu --steps 20 --warmup 3"
run() { timeout -k 5 80 env "$@" $B > gpurun_out/ab.json 2> gpurun_out/ab.err; rc=$?; echo "$* rc=$rc $(python -c "import json;d=json.load(open('gpurun_out/ab.json'));print(round(d['ms_per_step'],3),'ms',round(d['last_metrics']['model_loss'],4))" 2>/dev/null)" | tee -a gpurun_out/summary.txt; }
run DD_NONE=1
run DD_LN_VEC=1
run DD_GEMM_THIN=2
run DD_GEMM_TC_STAGED=1
run DD_LN_VEC=1 DD_GEMM_THIN=2 DD_GEMM_TC_STAGED=1
