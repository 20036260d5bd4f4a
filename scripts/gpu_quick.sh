#!/bin/bash
mkdir -p gpurun_out
for ks in 2 1; do
echo "== DD_MACHINE_KS=$ks"
timeout -k 5 120 env DD_MACHINE_KS=$ks python scripts/rollout_prof.py 2>&1 | tail -n 6
done
