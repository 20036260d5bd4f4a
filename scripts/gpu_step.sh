#!/bin/bash
# engine on the layer machine: the GPU test suite, then phase times and the bench line with the
# machine / the fused rollout switched on and off
mkdir -p gpurun_out
timeout -k 5 420 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/gpu_tests.log 2>&1
echo "tests rc=$?"; tail -n 12 gpurun_out/gpu_tests.log | cut -c1-250
for cfg in "DD_ROLLOUT=0" "DD_X=1" "DD_MACHINE_MAXROWS=100000"; do
  echo "== $cfg"
  timeout -k 5 120 env $cfg python scripts/section_times.py 2>&1 | tail -n 12
done
B="python bench.py --no-cpu --steps 20 --warmup 3"
for cfg in "DD_X=1"; do
  timeout -k 5 100 env $cfg $B > gpurun_out/step_$cfg.json 2> gpurun_out/step_$cfg.err; echo "$cfg rc=$? $(python -c "import json;d=json.load(open('gpurun_out/step_$cfg.json'));print(round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'],round(d['last_metrics']['model_loss'],4))" 2>/dev/null)"
done
timeout -k 5 60 python scripts/machine_prof.py > gpurun_out/machine_prof.txt 2>&1; grep -A8 "LN chain" gpurun_out/machine_prof.txt | head -24
