import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from reactive_pb_nn_md_b200 import system, engine
from reactive_pb_nn_md_b200._binding import Library, load_cuda
from tests.util import small_params, rel_rms
s = system.build_acid_box(10, ion_pair=True)
so = engine.Simulation(s, small_params(), library=Library("oracle/librpbmd_oracle.so"))
sg = engine.Simulation(s, small_params(), library=load_cuda())
so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
eo = so.evb()
print("S", eo["n_states"], "principal", eo["principal_diabat"])
for k in range(eo["n_states"]):
    c = np.zeros(eo["n_states"]); c[k] = 1.0
    fg, fo = sg.debug_mix_forces(c), so.debug_mix_forces(c)
    d = np.abs(fg - fo).max(axis=1)
    w = np.argsort(-d)[:6]
    print("state", k + 1, "log", eo["proton_log"][k, :, :].tolist(), "rel_rms %.2e" % rel_rms(fg, fo), "worst atoms", w.tolist(), np.round(d[w], 4).tolist())
c = eo["eigenvector"]
fg, fo = sg.debug_mix_forces(c), so.debug_mix_forces(c)
d = np.abs(fg - fo).max(axis=1); w = np.argsort(-d)[:8]
print("ground state mix rel_rms %.2e" % rel_rms(fg, fo), w.tolist(), np.round(d[w], 4).tolist())
a, b = sg.download_state(), so.download_state()
d = np.abs(a["force"] - b["force"]).max(axis=1); w = np.argsort(-d)[:8]
print("committed force diff", w.tolist(), np.round(d[w], 4).tolist())
print("mol_first of hydronium etc", b["mol_first_atom"][:3], b["mol_first_atom"][645:650])
