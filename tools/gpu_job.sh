python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_gputest_final.log; cat gpurun_out/r02_gputest_final.log
python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/bench_err.log > gpurun_out/r02_bench_c3_final.json; tail -2 gpurun_out/bench_err.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>gpurun_out/bench_ref_err.log > gpurun_out/r02_bench_reference.json; tail -2 gpurun_out/bench_ref_err.log; cat gpurun_out/r02_bench_reference.json | cut -c1-600
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c3_final.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['n_states'], 'launches', d['gpu_launches'], d['clocks'])
print(d['roofline']); print(d['cpu_baseline'])
for k,v in d['other_workloads'].items(): print(k, v['value'], v.get('ms_per_step'), v.get('concurrency_gain'))
"
