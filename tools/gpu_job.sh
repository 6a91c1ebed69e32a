cat > /tmp/few.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
from reactive_pb_nn_md_b200 import engine
from reactive_pb_nn_md_b200._binding import load_cuda
s = bench.build_system("c3")
sim = engine.Simulation(s, bench.params_for("c3"), library=load_cuda())
sim.ms_evb_calculate_total_force_energy()
sim.md_integrate_atomic(6, ms_evb=True)
PY
ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,launch__grid_size,launch__block_size,launch__registers_per_thread --clock-control none --launch-skip 200 -c 60 --csv --log-file gpurun_out/r02_smtime.csv python /tmp/few.py > gpurun_out/r02_smtime.log 2>&1; tail -2 gpurun_out/r02_smtime.log
