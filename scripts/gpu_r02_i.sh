#!/bin/bash
# reproducibility of the training trajectory: the same 23 steps, several times per configuration
mkdir -p gpurun_out
for cfg in "DD_REVERSE=0" "DD_REVERSE=0" "DD_REVERSE=0 DD_OVERLAP=0" "DD_REVERSE=0 DD_OVERLAP=0" "DD_X=1" "DD_X=1" "DD_OVERLAP=0" "DD_MACHINE=0" "DD_MACHINE=0"; do
  timeout -k 5 100 env $cfg python bench.py --no-cpu --steps 20 --warmup 3 > gpurun_out/rep.json 2> gpurun_out/rep.err; echo "$cfg rc=$? $(python -c "import json;d=json.load(open('gpurun_out/rep.json'));print(round(d['ms_per_step'],3), d['last_metrics'])" 2>/dev/null)"
done
