#!/bin/bash
# HEAD sanity: the shipped .so on a fresh box
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_machine_gpu.py tests/test_train_gpu.py -x -q --timeout 120 2>&1 | tail -n 3
timeout -k 5 100 python bench.py --no-cpu --steps 20 --warmup 5 > gpurun_out/step_head.json 2> gpurun_out/step_head.err; echo "bench rc=$? $(python -c "import json;d=json.load(open('gpurun_out/step_head.json'));print(round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'],d['last_metrics'])")"
