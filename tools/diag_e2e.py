import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from reactive_pb_nn_md_b200 import engine
from reactive_pb_nn_md_b200._binding import load_cuda
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
s = bench.build_system(wl)
sim = engine.Simulation(s, bench.params_for(wl), library=load_cuda())
sim.ms_evb_calculate_total_force_energy(); sim.md_integrate_atomic(20, ms_evb=True)
st = sim.download_state()
T = dict(upload=0.0, step=0.0, download=0.0, energies=0.0)
n = 200
for k in range(n + 10):
    if k == 10: T = dict.fromkeys(T, 0.0); torch.cuda.synchronize()
    t0 = time.perf_counter(); sim.upload_state(st["xyz"], st["velocity"], st)
    t1 = time.perf_counter(); sim.md_integrate_atomic(1, ms_evb=True)
    t2 = time.perf_counter(); st = sim.download_state(out=st)
    t3 = time.perf_counter(); sim.energies()
    t4 = time.perf_counter()
    T["upload"] += t1 - t0; T["step"] += t2 - t1; T["download"] += t3 - t2; T["energies"] += t4 - t3
print({k: round(1e6 * v / n, 1) for k, v in T.items()}, "us per step; total", round(1e6 * sum(T.values()) / n, 1))
