python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_gputest_final2.log; cat gpurun_out/r02_gputest_final2.log
python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/bench_err.log > gpurun_out/r02_bench_c3_v14.json; tail -2 gpurun_out/bench_err.log
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c3_v14.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['n_states'], 'launches', d['gpu_launches'], d['clocks'])
print(d['roofline']['us_per_launch'], d['roofline']['frac']); print(d['cpu_baseline']['value'])
for k,v in d['other_workloads'].items(): print(k, v['value'], v.get('ms_per_step'), v.get('concurrency_gain'))
"
