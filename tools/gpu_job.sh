python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_tiles_v2.log; cat gpurun_out/r02_tiles_v2.log
for v in 1 2 3 4 5; do RPB_PAIR_VARIANT=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload c3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant',$v, d['value'], d['kernels']['pair_real_space']['ms_per_step'], d['kernels']['verlet'])"; done 2>&1 | tee gpurun_out/r02_pair_sweep2.log
CMD="python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_tiles -s 4 -c 1 -f -o gpurun_out/prof_r02_pair_v2 $CMD > gpurun_out/ncu_pair.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_verlet_rebuild -c 1 -f -o gpurun_out/prof_r02_rebuild_v2 $CMD > gpurun_out/ncu_rebuild.log 2>&1
ls -la gpurun_out/*.ncu-rep
