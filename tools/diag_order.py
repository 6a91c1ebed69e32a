"""Pair-kernel time vs molecule order: the synthetic boxes list molecules in lattice order (spatially coherent); a
long-equilibrated liquid has no such order.  Same box, molecules shuffled."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from reactive_pb_nn_md_b200 import engine, system
from reactive_pb_nn_md_b200._binding import load_cuda
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
s0 = bench.build_system(wl)
def shuffled(s, seed):
    rng = np.random.default_rng(seed)
    M = s.n_mole
    keep_first = 1 if s.hydronium_mol else 0
    perm = np.concatenate((np.arange(keep_first), keep_first + rng.permutation(M - keep_first)))
    names = [s.ff.molecule_types[s.mol_type[m] - 1].name for m in perm]
    idx = np.concatenate([np.arange(s.mol_first_atom[m] - 1, s.mol_first_atom[m] - 1 + s.mol_n_atom[m]) for m in perm])
    return system.System(s.ff, s.box_length, names, s.xyz[idx], s.velocity[idx])
os.environ["RPB_SERIAL_STREAMS"] = "1"
evb = bench.WORKLOADS[wl]["ms_evb"]
for label, s in (("lattice order", s0), ("shuffled", shuffled(s0, 1))):
    sim = engine.Simulation(s, bench.params_for(wl), library=load_cuda())
    (sim.ms_evb_calculate_total_force_energy if evb else sim.calculate_total_force_energy)()
    sim.md_integrate_atomic(10, ms_evb=evb)
    sim.timers_enable(True); sim.timers(reset=True)
    sim.md_integrate_atomic(20, ms_evb=evb)
    tm = sim.timers()
    print("%-14s pair %.1f us  step %.1f us  verlet %.1f us" % (label, 1e3 * tm["pair_real_space"][0] / 20, 1e3 * tm["step_total"][0] / 20, 1e3 * tm["verlet"][0] / 20))
