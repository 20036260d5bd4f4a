#!/bin/bash
# new defaults (reverse program + backward trunk programs): every GPU test, then the 2-GPU line
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -n 5
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "n2 rc=$?"
tail -n 5 gpurun_out/r02_bench_n2.err | cut -c1-300
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n2.json'));print('n2', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'], d['last_metrics'])"
timeout -k 5 300 env DD_GRAPH_NCCL=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_n2_seg.json 2> gpurun_out/r02_bench_n2_seg.err; echo "n2 segments rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n2_seg.json'));print('n2 segmented', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'], d['last_metrics'])"
