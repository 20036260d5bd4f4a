#!/bin/bash
# Round-2 evidence pass (1 GPU): bench line of both arms, phase times, machine traces, the ncu launch
# list of one step, full captures of the two rollout programs, the other BASELINE configs, memcheck.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/r02_summary.txt; }
rm -f gpurun_out/r02_summary.txt
timeout -k 5 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; el bench $?
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n1.json'));print('bench', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['ms_per_step'],3),'launches',d['launches_per_step'],'roofline',round(d['roofline']['frac'],4), round(d['roofline']['achieved'],1),'cpu',d['cpu_baseline'] and round(d['cpu_baseline']['value']), d['clocks'])"
timeout -k 5 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; el bench_ref $?
cut -c1-200 gpurun_out/r02_bench_reference_arm.json
timeout -k 5 120 python scripts/section_times.py > gpurun_out/r02_phase_times.txt 2>&1; el phases $?; tail -n 12 gpurun_out/r02_phase_times.txt
timeout -k 5 200 python scripts/machine_prof.py > gpurun_out/r02_machine_trace.txt 2>&1; el machine_prof $?
timeout -k 5 120 python scripts/rollout_prof.py > gpurun_out/r02_rollout_timeline.txt 2>&1; el rollout_prof $?
timeout -k 5 120 python scripts/machine_bench.py > gpurun_out/r02_machine_bench.txt 2>&1; el machine_bench $?; tail -n 12 gpurun_out/r02_machine_bench.txt | cut -c1-200
timeout -k 5 60 python scripts/profile_step.py 2 > gpurun_out/step_counts.txt 2>&1; el counts $?
cat gpurun_out/step_counts.txt
S0=$(grep "step 0" gpurun_out/step_counts.txt | awk '{print $4}')
S1=$(grep "step 1" gpurun_out/step_counts.txt | awk '{print $4}')
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -s $S0 -c $S1 --csv \
  --log-file gpurun_out/r02_launches_one_step.csv python scripts/profile_step.py 2 > gpurun_out/ncu_step.log 2>&1; el ncu_step $?
timeout -k 5 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lm_kernel \
  -o gpurun_out/r02_lm_rollout_prof -f python scripts/profile_rollout.py > gpurun_out/ncu_lm.log 2>&1; el ncu_rollout $?
tail -n 3 gpurun_out/ncu_lm.log
timeout -k 5 300 python bench.py --config ur5 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_ur5.json 2> gpurun_out/r02_bench_ur5.err; el bench_ur5 $?
cut -c1-260 gpurun_out/r02_bench_ur5.json; tail -n 3 gpurun_out/r02_bench_ur5.err | cut -c1-300
timeout -k 5 400 python bench.py --config a1_b256 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_a1_b256_n1.json 2> gpurun_out/r02_bench_a1_b256.err; el bench_b256 $?
cut -c1-260 gpurun_out/r02_bench_a1_b256_n1.json; tail -n 3 gpurun_out/r02_bench_a1_b256.err | cut -c1-300
timeout -k 5 240 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_machine_gpu.py -x -q -k "trunk_chain or gru_layer" > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; el memcheck $?
tail -n 6 gpurun_out/r02_sanitizer_memcheck.txt | cut -c1-200
ls -la gpurun_out | grep r02
