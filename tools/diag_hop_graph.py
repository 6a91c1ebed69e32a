"""per-step device and host times around a committed hop, with the step graph on (default) or off (RPB_GRAPH=0)"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from reactive_pb_nn_md_b200 import system, engine
from reactive_pb_nn_md_b200._binding import load_cuda
from tests.util import small_params
s = system.config_c3()
sim = engine.Simulation(s, small_params(pme_grid=48), library=load_cuda())
sim.ms_evb_calculate_total_force_energy()
sim.md_integrate_atomic(8, ms_evb=True)
stream = torch.cuda.ExternalStream(sim.dll.rpb_get_stream(sim.ctx))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for k in range(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); a.record(stream)
    sim.md_integrate_atomic(1, ms_evb=True)
    b.record(stream); torch.cuda.synchronize(); t1 = time.perf_counter()
    if n > 20 and a.elapsed_time(b) < 0.35: continue
    print("step %2d  device %.3f ms  host %.3f ms  S %d hyd %d" % (9 + k, a.elapsed_time(b), 1e3 * (t1 - t0), sim.evb()["n_states"], sim.evb()["new_hydronium_mol"]), flush=True)
