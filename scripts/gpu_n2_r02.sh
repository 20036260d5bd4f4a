#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_n2.raw 2> gpurun_out/r02_bench_n2.err; echo "n2 rc=$? after $(( $(date +%s) - t0 ))s"
grep '^{' gpurun_out/r02_bench_n2.raw > gpurun_out/r02_bench_n2.json
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n2.json'));print('n2', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'value',round(d['value']),'launches',d['launches_per_step'], d['last_metrics'])"
t1=$(date +%s)
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --config a1_b256 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_a1_b256_n2.raw 2> gpurun_out/r02_bench_a1_b256_n2.err; echo "b256 n2 rc=$? after $(( $(date +%s) - t1 ))s"
grep '^{' gpurun_out/r02_bench_a1_b256_n2.raw > gpurun_out/r02_bench_a1_b256_n2.json
python -c "import json;d=json.load(open('gpurun_out/r02_bench_a1_b256_n2.json'));print('b256 n2', round(d['ms_per_step'],3),'ms value',round(d['value']),'launches',d['launches_per_step'])"
tail -n 3 gpurun_out/r02_bench_a1_b256_n2.err | cut -c1-300
