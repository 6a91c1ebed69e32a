for w in c3 c2 c4; do python tools/time_kernels.py $w 10 2>&1 | head -24; done > gpurun_out/r02_kernel_times_v12.log 2>&1; cat gpurun_out/r02_kernel_times_v12.log
RPB_PAIR_VARIANT=1 python tools/time_kernels.py c3 10 2>&1 | grep -E "variant|pair_real"
