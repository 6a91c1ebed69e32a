python -m pytest tests/test_gpu_full_size.py tests/test_gpu_evb_cases.py -m gpu -q -x -k "real_peers or peer_memory" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra 2>gpurun_out/bench2_err.log > gpurun_out/r02_bench_c3_n2_v15.json; tail -2 gpurun_out/bench2_err.log
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c3_n2_v15.json').read().strip().splitlines()[-1]); print('c3 n2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['n_states'], d['gpu_launches'])
"
