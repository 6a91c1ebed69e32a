#!/usr/bin/env python
"""C3 trajectory, CUDA library against the CPU oracle, printing the deviations after every chunk of steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from reactive_pb_nn_md_b200 import engine, system
from reactive_pb_nn_md_b200._binding import load_cuda, Library
load_oracle = lambda: Library(os.path.join(ROOT, 'oracle', 'librpbmd_oracle.so'))
from tests.util import small_params, rel_rms
s = system.config_c3()
so = engine.Simulation(s, small_params(pme_grid=48, n_threads=16), library=load_oracle())
sg = engine.Simulation(s, small_params(pme_grid=48), library=load_cuda())
so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
print("eval0 force rel rms", rel_rms(sg.forces(), so.forces()))
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for it in range(40 // chunk):
    so.md_integrate_atomic(chunk, ms_evb=True); sg.md_integrate_atomic(chunk, ms_evb=True)
    a, b = sg.download_state(), so.download_state()
    dx = np.abs(a["xyz"] - b["xyz"]); df = np.abs(a["force"] - b["force"])
    i = int(np.argmax(dx.max(axis=1)))
    print("step %3d  max|dx| %.3e (atom %d)  max|dF| %.3e  relrms F %.3e  hyd %d/%d  S %d/%d  npairs %d/%d" % (
        (it + 1) * chunk, dx.max(), i, df.max(), rel_rms(a["force"], b["force"]), a["hydronium_mol"], b["hydronium_mol"],
        sg.evb()["n_states"], so.evb()["n_states"], len(sg.tile_pairs()[0]), len(so.neighbor_list()[1])))
