python -m pytest tests/test_gpu_nonreactive.py -m gpu -q -x 2>&1 | tail -2
for v in 1 2; do RPB_PAIR_VARIANT=$v python tools/time_kernels.py c2 10 2>&1 | grep -E "variant|pair_real"; done
for v in 1 2; do RPB_PAIR_VARIANT=$v python tools/time_kernels.py c4 5 2>&1 | grep -E "variant|pair_real"; done
