#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 env DD_PROFILE_LAST=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_one_step.csv python scripts/profile_step.py 2 > gpurun_out/ncu_step.log 2>&1; echo "ncu_step rc=$?"
tail -n 2 gpurun_out/ncu_step.log
timeout -k 5 120 python scripts/rollout_prof.py > gpurun_out/r02_rollout_timeline.txt 2>&1; echo "rollout_prof rc=$?"; tail -n 5 gpurun_out/r02_rollout_timeline.txt | cut -c1-200
timeout -k 5 200 python bench.py --config a1_b256 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_a1_b256_n1.json 2> gpurun_out/r02_bench_a1_b256.err; echo "b256 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_a1_b256_n1.json'));print('b256', round(d['ms_per_step'],3), d['launches_per_step'], round(d['value']))"; tail -n 3 gpurun_out/r02_bench_a1_b256.err | cut -c1-300
