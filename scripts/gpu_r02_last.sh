#!/bin/bash
# last pass of the round: launch list of exactly one step (profiler-bracketed), smoke(), every GPU test
mkdir -p gpurun_out
timeout -k 5 300 env DD_PROFILE_LAST=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_one_step.csv python scripts/profile_step.py 2 > gpurun_out/ncu_step.log 2>&1; echo "ncu_step rc=$?"
tail -n 2 gpurun_out/ncu_step.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02_smoke.log
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02_gpu_tests.log 2>&1; echo "pytest_gpu rc=$?"
tail -n 3 gpurun_out/r02_gpu_tests.log
grep -E "worst gradient" gpurun_out/r02_gpu_tests.log | cut -c1-260
timeout -k 5 100 python bench.py --no-cpu --steps 20 --warmup 5 > gpurun_out/step_last.json 2> gpurun_out/step_last.err; echo "bench rc=$? $(python -c "import json;d=json.load(open('gpurun_out/step_last.json'));print(round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'],d['last_metrics'],'traffic',d['roofline']['traffic'])")"
