python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_gputest_final3.log; cat gpurun_out/r02_gputest_final3.log
python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/bench_err.log > gpurun_out/r02_bench_c3_v15.json; tail -2 gpurun_out/bench_err.log
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c3_v15.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['n_states'], 'launches', d['gpu_launches'], d['clocks'])
print(d['roofline']['us_per_launch'], d['roofline']['frac']); print(d['cpu_baseline']['value'])
for k,v in d['other_workloads'].items(): print(k, v['value'], v.get('ms_per_step'), v.get('concurrency_gain'))
"
CMD="python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
ncu --set full --import-source on --clock-control none -k regex:k_pair_tiles --launch-skip 8 -c 1 -f -o gpurun_out/prof_r02_v15_pair $CMD > gpurun_out/r02_v15_ncu_p.log 2>&1
ncu -i gpurun_out/prof_r02_v15_pair.ncu-rep --page source --csv > gpurun_out/r02_v15_pair_src.csv 2>/dev/null
ncu -i gpurun_out/prof_r02_v15_pair.ncu-rep --page raw --csv > gpurun_out/r02_v15_pair_raw.csv 2>/dev/null
ls -la gpurun_out/prof_r02_v15_pair.ncu-rep
