#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lm_kernel \
  -o gpurun_out/r02_lm_rollout_prof -f python scripts/profile_rollout.py > gpurun_out/ncu_lm.log 2>&1; echo "ncu_rollout rc=$?"
tail -n 3 gpurun_out/ncu_lm.log
