#!/bin/bash
# Round-2 evidence pass: every -m gpu test, smoke(), the bench line of both arms, phase times, and
# the ncu launch list + a full capture of the layer machine (rollout program) and of one trunk GEMM.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/r02_summary.txt; }
rm -f gpurun_out/r02_summary.txt
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 300 -s > gpurun_out/r02_gpu_tests.log 2>&1; el pytest_gpu $?
grep -E "passed|failed|FAILED|worst gradient|3xTF32" gpurun_out/r02_gpu_tests.log | cut -c1-300 | tail -12
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; el smoke $?; tail -n 2 gpurun_out/r02_smoke.log
timeout -k 5 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; el bench $?
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n1.json'));print('bench', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'],'roofline',round(d['roofline']['frac'],4), round(d['roofline']['achieved'],1),'cpu',d['cpu_baseline'] and round(d['cpu_baseline']['value']))"
timeout -k 5 400 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; el bench_ref $?
cut -c1-300 gpurun_out/r02_bench_reference_arm.json
timeout -k 5 120 python scripts/section_times.py > gpurun_out/r02_phase_times.txt 2>&1; el phases $?; tail -n 12 gpurun_out/r02_phase_times.txt
timeout -k 5 60 python scripts/profile_step.py 2 > gpurun_out/step_counts.txt 2>&1; el counts $?
cat gpurun_out/step_counts.txt
S0=$(grep "step 0" gpurun_out/step_counts.txt | awk '{print $4}')
S1=$(grep "step 1" gpurun_out/step_counts.txt | awk '{print $4}')
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -s $S0 -c $S1 --csv \
  --log-file gpurun_out/r02_launches_one_step.csv python scripts/profile_step.py 2 > gpurun_out/ncu_step.log 2>&1; el ncu_step $?
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 3 -c 2 \
  -o gpurun_out/r02_lm_prof -f python scripts/profile_step.py 1 > gpurun_out/ncu_lm.log 2>&1; el ncu_lm $?
tail -n 2 gpurun_out/ncu_step.log gpurun_out/ncu_lm.log; ls -la gpurun_out | grep r02
