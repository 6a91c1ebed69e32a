"""Long NVE run of a bench workload on the CUDA library: conservation, hop count, diabat count range, no status errors."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from reactive_pb_nn_md_b200 import engine
from reactive_pb_nn_md_b200._binding import load_cuda
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
chunk = 1000
s = bench.build_system(wl)
evb = bench.WORKLOADS[wl]["ms_evb"]
sim = engine.Simulation(s, bench.params_for(wl), library=load_cuda())
(sim.ms_evb_calculate_total_force_energy if evb else sim.calculate_total_force_energy)()
e = sim.energies(); e0 = e["potential_energy"] + e["kinetic_energy"]; ke0 = e["kinetic_energy"]
hyd, hops, smin, smax = sim.download_state()["hydronium_mol"], 0, 10 ** 9, 0
t0 = time.time()
for k in range(n_total // chunk):
    sim.md_integrate_atomic(chunk, ms_evb=evb)
    e = sim.energies(); st = sim.download_state()
    S = sim.evb()["n_states"] if evb else 1
    smin, smax = min(smin, S), max(smax, S)
    if st["hydronium_mol"] != hyd: hops += 1; hyd = st["hydronium_mol"]
    et = e["potential_energy"] + e["kinetic_energy"]
    print("step %6d  Etot-E0 %+10.3f kJ/mol (%.2e of KE)  KE %.1f  S %d  hydronium %d  max|F| %.0f" % ((k + 1) * chunk, et - e0, abs(et - e0) / ke0, e["kinetic_energy"], S, hyd, np.abs(st["force"]).max()), flush=True)
print("done: %d steps in %.1f s (%.0f steps/s incl. read-backs), hydronium changed at %d of %d checkpoints, S in [%d, %d]" % (n_total, time.time() - t0, n_total / (time.time() - t0), hops, n_total // chunk, smin, smax))
