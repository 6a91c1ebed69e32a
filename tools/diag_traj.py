import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from reactive_pb_nn_md_b200 import system, engine
from reactive_pb_nn_md_b200._binding import Library, load_cuda
olib = Library(os.path.join("oracle","librpbmd_oracle.so")); glib = load_cuda()
s = system.config_c3()
so = engine.Simulation(s, engine.SimulationParameters(pme_grid=48, n_threads=16), library=olib)
sg = engine.Simulation(s, engine.SimulationParameters(pme_grid=48), library=glib)
so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    try:
        so.md_integrate_atomic(1, ms_evb=True); sg.md_integrate_atomic(1, ms_evb=True)
    except Exception as e:
        print("ERR step", k+1, e); break
    a, b = sg.download_state(), so.download_state()
    eg, eo = sg.evb(), so.evb()
    fg, fo = a["force"], b["force"]
    vg, ng, flg = sg.neighbor_list(); vo, no, flo = so.neighbor_list()
    print(k+1, "dx %.2e dv %.2e dF %.2e S %d/%d hyd %d/%d pd %d/%d E %.6f/%.6f nl_equal %s flag %d/%d types_eq %s" % (
        np.abs(a["xyz"]-b["xyz"]).max(), np.abs(a["velocity"]-b["velocity"]).max(), np.abs(fg-fo).max(), eg["n_states"], eo["n_states"],
        a["hydronium_mol"], b["hydronium_mol"], eg["principal_diabat"], eo["principal_diabat"], eg["adiabatic_potential"], eo["adiabatic_potential"],
        np.array_equal(ng,no) and np.array_equal(vg,vo), flg, flo, np.array_equal(a["atom_type"], b["atom_type"])), flush=True)
