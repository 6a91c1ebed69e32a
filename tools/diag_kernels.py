"""Per-kernel CUDA-event times (library timers) averaged over n steps of workload c3, no overlap distortion removed."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from reactive_pb_nn_md_b200 import system, engine
from reactive_pb_nn_md_b200._binding import load_cuda
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
names = sys.argv[2].split(",") if len(sys.argv) > 2 else None
s = system.config_c3()
sim = engine.Simulation(s, engine.SimulationParameters(pme_grid=48), library=load_cuda())
sim.ms_evb_calculate_total_force_energy()
sim.md_integrate_atomic(5, ms_evb=True)
sim.timers_enable(True); sim.timers(reset=True)
sim.md_integrate_atomic(n, ms_evb=True)
t = sim.timers()
print(" ".join("%s=%.1f" % (k, v[0] / n * 1e3) for k, v in t.items() if v[1] and (names is None or k in names)))
