python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r02_gputest_2gpu_v4.log; cat gpurun_out/r02_gputest_2gpu_v4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench2_err.log > gpurun_out/r02_bench_c3_n2_v4.json; tail -3 gpurun_out/bench2_err.log
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c3_n2_v4.json').read().strip().splitlines()[-1]); print('c3 n2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['n_states'], d['gpu_launches'])
for k,v in d['other_workloads'].items(): print(k, v['value'], v.get('ms_per_step'), v.get('concurrency_gain'))
"
