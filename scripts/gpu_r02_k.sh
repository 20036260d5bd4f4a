#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "n2 rc=$? after $(( $(date +%s) - t0 ))s"
tail -n 3 gpurun_out/r02_bench_n2.err | cut -c1-300
grep '^{' gpurun_out/r02_bench_n2.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('n2', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'], d['last_metrics'])"
timeout -k 5 120 python scripts/rollout_prof.py > gpurun_out/rollout_prof.txt 2>&1; echo "rollout_prof rc=$?"
cut -c1-260 gpurun_out/rollout_prof.txt | sed -n 1,40p; tail -n 5 gpurun_out/rollout_prof.txt | cut -c1-250
