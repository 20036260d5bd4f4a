#!/bin/bash
# Round-2 starting point: full ncu captures of the kernels r01 flagged (scripts/ncu_extract.py summarises them):
#   block-per-row LayerNorm at M ~ 15K (26-29 % of HBM copy peak), the SIMT tile kernel on the N = 12
#   action-head shapes, the experimental thin / staged GEMM variants, the forward scan.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/summary.txt; }
P="python scripts/profile_step.py 1"
timeout -k 5 120 ncu --set full --clock-control none --import-source on -k regex:ln_fwd_kernel -s 60 -c 2 -o gpurun_out/ln_fwd_prof -f $P > gpurun_out/ncu_ln_fwd.log 2>&1; el ncu_ln_fwd $?
timeout -k 5 120 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_kernel -s 20 -c 2 -o gpurun_out/ln_bwd_prof -f $P > gpurun_out/ncu_ln_bwd.log 2>&1; el ncu_ln_bwd $?
timeout -k 5 120 ncu --set full --clock-control none --import-source on -k regex:^gemm_kernel -s 10 -c 3 -o gpurun_out/simt_prof -f $P > gpurun_out/ncu_simt.log 2>&1; el ncu_simt $?
timeout -k 5 120 env DD_LN_VEC=1 ncu --set full --clock-control none --import-source on -k regex:ln_.*_vec_kernel -s 60 -c 3 -o gpurun_out/ln_vec_prof -f $P > gpurun_out/ncu_ln_vec.log 2>&1; el ncu_ln_vec $?
timeout -k 5 120 env DD_GEMM_THIN=2 ncu --set full --clock-control none --import-source on -k regex:thin_ -s 4 -c 3 -o gpurun_out/thin_prof -f $P > gpurun_out/ncu_thin.log 2>&1; el ncu_thin $?
timeout -k 5 120 env DD_GEMM_TC_STAGED=1 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 200 -c 3 -o gpurun_out/staged_prof -f $P > gpurun_out/ncu_staged.log 2>&1; el ncu_staged $?
timeout -k 5 120 ncu --set full --clock-control none --import-source on -k regex:observe_fwd_kernel -c 1 -o gpurun_out/fwd_scan_prof -f $P > gpurun_out/ncu_fwd_scan.log 2>&1; el ncu_fwd_scan $?
ls -la gpurun_out/*.ncu-rep
