#!/bin/bash
# MMA k-loop microbenchmark + machine trace with the issue / commit split
mkdir -p gpurun_out
timeout -k 5 120 scripts/micro/mma_loop > gpurun_out/mma_loop.txt 2>&1; echo "mma_loop rc=$?"
cat gpurun_out/mma_loop.txt
timeout -k 5 200 python scripts/machine_prof.py > gpurun_out/machine_prof.txt 2>&1; echo "prof rc=$?"
cut -c1-200 gpurun_out/machine_prof.txt | head -40
