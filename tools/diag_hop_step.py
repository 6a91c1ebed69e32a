"""timeline of the MS-EVB steps around a committed proton hop (C3: the hop of step ~15)"""
import os, sys
os.environ["RPB_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.getcwd())
from reactive_pb_nn_md_b200 import system, engine
from reactive_pb_nn_md_b200._binding import load_cuda
from tests.util import small_params
s = system.config_c3()
sim = engine.Simulation(s, small_params(pme_grid=48), library=load_cuda())
sim.ms_evb_calculate_total_force_energy()
sim.md_integrate_atomic(8, ms_evb=True)
sim.timers_enable(True)
for k in range(10):
    sim.timers(reset=True)
    sim.md_integrate_atomic(1, ms_evb=True)
    tm = sim.timers()
    print("step", 9 + k, "total %.1f us" % (1e3 * tm["step_total"][0]), "hyd", sim.evb()["new_hydronium_mol"], flush=True)
