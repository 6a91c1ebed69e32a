#!/bin/bash
# pair-kernel variants (RPB_PAIR_VARIANT, kernels_pair.cu) on one workload: serial-stream kernel time from bench.py's table
WL=${1:-c4}
for v in ${VARIANTS:-0 1 2 3 4 5 6 7 8 9 10 11}; do
  RPB_PAIR_VARIANT=$v python bench.py --workload $WL --steps 20 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernels']['pair_real_space']
print('variant $v  pair %.1f us  step %.4f ms' % (1e3 * k['ms_per_step'], d['ms_per_step']))"
done
