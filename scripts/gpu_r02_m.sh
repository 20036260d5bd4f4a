#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_machine_gpu.py -x -q --timeout 60 > gpurun_out/machine_tests.log 2>&1
echo "machine tests rc=$?"; grep -E "passed|failed|FAILED|machine debug|Error|assert" gpurun_out/machine_tests.log | cut -c1-250 | tail -6
timeout -k 5 120 python scripts/rollout_prof.py > gpurun_out/rollout_prof.txt 2>&1; echo "rollout_prof rc=$?"
cut -c1-200 gpurun_out/rollout_prof.txt | sed -n 12,26p; tail -n 5 gpurun_out/rollout_prof.txt | cut -c1-250
for cfg in "DD_X=1" "DD_MACHINE_DEBUG=32"; do
  echo "== $cfg"
  timeout -k 5 120 env $cfg python scripts/section_times.py 2>&1 | tail -n 12 | grep -E "imagine|actor_update|wm_|sum"
done
for cfg in "DD_X=1" "DD_MACHINE_DEBUG=32" "DD_X=2"; do
  timeout -k 5 100 env $cfg python bench.py --no-cpu --steps 20 --warmup 3 > gpurun_out/step_$cfg.json 2> gpurun_out/step_$cfg.err; echo "$cfg rc=$? $(python -c "import json;d=json.load(open('gpurun_out/step_$cfg.json'));print(round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'],d['last_metrics'], 'roofline', round(d['roofline']['frac'],4), round(d['roofline']['achieved'],1))" 2>/dev/null)"
done
timeout -k 5 600 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -n 4
