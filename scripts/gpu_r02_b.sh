#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 120 scripts/micro/mma_loop > gpurun_out/mma_loop.txt 2>&1; echo "mma_loop rc=$?"
cat gpurun_out/mma_loop.txt
timeout -k 5 200 python scripts/machine_prof.py > gpurun_out/machine_prof.txt 2>&1; echo "prof rc=$?"
cut -c1-200 gpurun_out/machine_prof.txt | sed -n 14,32p
timeout -k 5 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n1.json'));print('bench', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'],'roofline',round(d['roofline']['frac'],4), round(d['roofline']['achieved'],1),'cpu',d['cpu_baseline'])"
timeout -k 5 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "bench_ref rc=$?"
cut -c1-200 gpurun_out/r02_bench_reference_arm.json
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 5 -c 1 \
  -o gpurun_out/r02_lm_rollout_prof -f python scripts/profile_step.py 1 > gpurun_out/ncu_lm.log 2>&1; echo "ncu rc=$?"
tail -n 2 gpurun_out/ncu_lm.log
timeout -k 5 200 python -m pytest tests/test_train_gpu.py -q -k three_steps -s 2>&1 | grep -E "passed|failed|worst|Error" | cut -c1-300
