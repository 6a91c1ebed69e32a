import numpy as np, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from reactive_pb_nn_md_b200 import system, engine
from reactive_pb_nn_md_b200._binding import load_cuda
s = system.config_c3()
sim = engine.Simulation(s, engine.SimulationParameters(pme_grid=48), library=load_cuda())
sim.ms_evb_calculate_total_force_energy()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
out = []
for k in range(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sim.md_integrate_atomic(1, ms_evb=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    out.append((k + 1, dt, sim.evb()["n_states"], sim.download_state()["hydronium_mol"]))
print(" ".join("%d:%.2fms/S%d/h%d" % o for o in out))
