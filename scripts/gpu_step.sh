#!/bin/bash
# engine on the layer machine: machine unit tests, the GPU test suite, phase times with the
# reverse program on / off, and the bench line
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_machine_gpu.py -x -q --timeout 60 -s > gpurun_out/machine_tests.log 2>&1
echo "machine tests rc=$?"; grep -E "passed|failed|FAILED|machine debug|Error" gpurun_out/machine_tests.log | cut -c1-250 | tail -6
timeout -k 5 600 python -m pytest tests -m gpu -x -q --timeout 200 -s > gpurun_out/gpu_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|FAILED|worst gradient|Error" gpurun_out/gpu_tests.log | cut -c1-300 | tail -12
for cfg in "DD_X=1" "DD_REVERSE=0"; do
  echo "== $cfg"
  timeout -k 5 120 env $cfg python scripts/section_times.py 2>&1 | tail -n 12
done
B="python bench.py --no-cpu --steps 20 --warmup 3"
for cfg in "DD_X=1"; do
  timeout -k 5 100 env $cfg $B > gpurun_out/step_$cfg.json 2> gpurun_out/step_$cfg.err; echo "$cfg rc=$? $(python -c "import json;d=json.load(open('gpurun_out/step_$cfg.json'));print(round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'],round(d['last_metrics']['model_loss'],4), 'roofline', round(d['roofline']['frac'],4), round(d['roofline']['achieved'],1))" 2>/dev/null)"
done
