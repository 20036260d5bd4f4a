#!/bin/bash
mkdir -p gpurun_out
# (1) full-detail capture of the tensor-core GEMM (3 launches of the big NN shape)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 12 -c 3 \
  -o gpurun_out/tc_prof -f python scripts/tc_gemm_bench.py > gpurun_out/ncu_tc.log 2>&1
echo "ncu tc rc=$?"
# (2) every launch of one eager train step with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 2500 --csv \
  --log-file gpurun_out/launches.csv python scripts/profile_step.py 2 > gpurun_out/ncu_step.log 2>&1
echo "ncu step rc=$?"
tail -n 3 gpurun_out/ncu_tc.log gpurun_out/ncu_step.log
ls -la gpurun_out | head -20
