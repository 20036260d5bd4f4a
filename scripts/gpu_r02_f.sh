#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -x -q --timeout 60 2>&1 | tail -n 3
for cfg in "DD_X=1" "DD_REVERSE=1"; do
  echo "== $cfg"
  timeout -k 5 120 env $cfg python scripts/section_times.py 2>&1 | tail -n 12
done
for cfg in "DD_X=1" "DD_REVERSE=1"; do
  timeout -k 5 100 env $cfg python bench.py --no-cpu --steps 20 --warmup 3 > gpurun_out/step_$cfg.json 2> gpurun_out/step_$cfg.err; echo "$cfg rc=$? $(python -c "import json;d=json.load(open('gpurun_out/step_$cfg.json'));print(round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'],round(d['last_metrics']['model_loss'],4), 'roofline', round(d['roofline']['frac'],4), round(d['roofline']['achieved'],1))" 2>/dev/null)"
done
timeout -k 5 600 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -n 4
