python tools/diag_c3_traj.py 20 2>&1 | tail -3
for v in 0 1; do RPB_PAIR_VARIANT=$v python tools/time_kernels.py c3 10 2>&1 | grep -E "variant|pair_real" ; done
python tools/time_kernels.py c2 10 2>&1 | grep -E "variant|pair_real"
python tools/time_kernels.py c4 10 2>&1 | grep -E "variant|pair_real"
