#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02_gpu_tests.log 2>&1; echo "pytest_gpu rc=$?"
tail -n 3 gpurun_out/r02_gpu_tests.log; grep -E "worst gradient|FAILED|Error" gpurun_out/r02_gpu_tests.log | cut -c1-260 | head
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02_smoke.log
timeout -k 5 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n1.json'));print('bench', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'],'roofline',round(d['roofline']['frac'],4), round(d['roofline']['achieved'],1), d['last_metrics'])"
timeout -k 5 120 python scripts/section_times.py > gpurun_out/r02_phase_times.txt 2>&1; tail -n 12 gpurun_out/r02_phase_times.txt
timeout -k 5 120 python scripts/machine_bench.py > gpurun_out/r02_machine_bench.txt 2>&1
timeout -k 5 200 python bench.py --config ur5 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_ur5.json 2> gpurun_out/r02_bench_ur5.err; echo "ur5 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r02_bench_ur5.json'));print('ur5', round(d['ms_per_step'],3), d['launches_per_step'])"
