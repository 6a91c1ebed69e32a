import os, sys
os.environ["RPB_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.getcwd())
import torch
from reactive_pb_nn_md_b200 import system, engine
from reactive_pb_nn_md_b200._binding import load_cuda
s = system.config_c3()
sim = engine.Simulation(s, engine.SimulationParameters(pme_grid=48), library=load_cuda())
sim.ms_evb_calculate_total_force_energy()
sim.md_integrate_atomic(int(sys.argv[1]) if len(sys.argv) > 1 else 8, ms_evb=True)
sim.timers_enable(True)
for k in range(3):
    sim.timers(reset=True)
    sim.md_integrate_atomic(1, ms_evb=True)
    sim.timers()
