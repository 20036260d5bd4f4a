#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.raw 2> gpurun_out/r02_bench_n8.err; echo "n8 rc=$? after $(( $(date +%s) - t0 ))s"
tail -n 3 gpurun_out/r02_bench_n8.err | cut -c1-300
grep '^{' gpurun_out/r02_bench_n8.raw > gpurun_out/r02_bench_n8.json
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n8.json'));print('n8', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'value',round(d['value']),'launches',d['launches_per_step'], d['last_metrics'], d['clocks'])"
t1=$(date +%s)
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_n4.raw 2> gpurun_out/r02_bench_n4.err; echo "n4 rc=$? after $(( $(date +%s) - t1 ))s"
grep '^{' gpurun_out/r02_bench_n4.raw > gpurun_out/r02_bench_n4.json
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n4.json'));print('n4', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'value',round(d['value']))"
