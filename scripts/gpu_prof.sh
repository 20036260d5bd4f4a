#!/bin/bash
# ncu evidence for profiles/: (1) every launch of one eager a1 train step with its device time,
# (2) full-detail captures of the tensor-core GEMM and of the fused BPTT scan kernel.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
timeout -k 5 60 python scripts/profile_step.py 2 > gpurun_out/step_counts.txt 2>&1; el counts $?
cat gpurun_out/step_counts.txt
S0=$(grep "step 0" gpurun_out/step_counts.txt | awk '{print $4}')
S1=$(grep "step 1" gpurun_out/step_counts.txt | awk '{print $4}')
timeout -k 5 240 ncu --metrics gpu__time_duration.sum --clock-control none -s $S0 -c $S1 --csv \
  --log-file gpurun_out/launches.csv python scripts/profile_step.py 2 > gpurun_out/ncu_step.log 2>&1; el ncu_step $?
timeout -k 5 150 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 200 -c 3 \
  -o gpurun_out/tc_prof -f python scripts/profile_step.py 1 > gpurun_out/ncu_tc.log 2>&1; el ncu_tc $?
timeout -k 5 150 ncu --set full --clock-control none --import-source on -k regex:observe_bwd_kernel -c 1 \
  -o gpurun_out/bptt_prof -f python scripts/profile_step.py 1 > gpurun_out/ncu_bptt.log 2>&1; el ncu_bptt $?
tail -n 2 gpurun_out/ncu_step.log gpurun_out/ncu_tc.log gpurun_out/ncu_bptt.log; ls -la gpurun_out | head -30
