python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/bench8_err.log > gpurun_out/r02_bench_c3_n8.json; tail -2 gpurun_out/bench8_err.log
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c3_n8.json').read().strip().splitlines()[-1]); print('c3 n8', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['n_states'])
for k,v in d['other_workloads'].items(): print(k, v['value'], v.get('ms_per_step'), v.get('concurrency_gain'))
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 2>gpurun_out/bench4_err.log > gpurun_out/r02_bench_c3_n4.json; tail -2 gpurun_out/bench4_err.log
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c3_n4.json').read().strip().splitlines()[-1]); print('c3 n4', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['n_states'])
for k,v in d['other_workloads'].items(): print(k, v['value'], v.get('ms_per_step'), v.get('concurrency_gain'))
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench2_err.log > gpurun_out/r02_bench_c3_n2.json; tail -2 gpurun_out/bench2_err.log
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c3_n2.json').read().strip().splitlines()[-1]); print('c3 n2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['n_states'])
for k,v in d['other_workloads'].items(): print(k, v['value'], v.get('ms_per_step'), v.get('concurrency_gain'))
"
