#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_machine_gpu.py -x -q --timeout 90 > gpurun_out/machine_tests.log 2>&1
echo "machine tests rc=$?"; grep -E "passed|failed|FAILED|machine debug|Error|assert|stuck" gpurun_out/machine_tests.log | cut -c1-250 | tail -8
timeout -k 5 120 python scripts/machine_bench.py > gpurun_out/machine_bench.txt 2>&1; echo "bench rc=$?"; tail -n 8 gpurun_out/machine_bench.txt | cut -c1-220
timeout -k 5 120 env DD_MACHINE_DEBUG=64 python scripts/machine_bench.py > gpurun_out/machine_bench_nodual.txt 2>&1; echo "bench nodual rc=$?"; tail -n 4 gpurun_out/machine_bench_nodual.txt | cut -c1-220
for cfg in "DD_X=1" "DD_MACHINE_MAXROWS=16384"; do
  echo "== $cfg"
  timeout -k 5 120 env $cfg python scripts/section_times.py 2>&1 | tail -n 12 | grep -E "wm_|imagine|traj|critic|actor|sum"
done
for cfg in "DD_X=1" "DD_MACHINE_MAXROWS=16384"; do
  timeout -k 5 100 env $cfg python bench.py --no-cpu --steps 20 --warmup 3 > gpurun_out/step_$cfg.json 2> gpurun_out/step_$cfg.err; echo "$cfg rc=$? $(python -c "import json;d=json.load(open('gpurun_out/step_$cfg.json'));print(round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'],d['last_metrics'])" 2>/dev/null)"
  tail -n 2 gpurun_out/step_$cfg.err | cut -c1-200
done
