"""Driver that runs an unmodified reference input deck on the CUDA library, in the order of main_ms_evb
(src/main_ms_evb.f90:20-118): read simulation parameters -> read .gro/.pmt/.top -> tables -> initial force ->
velocities -> print step 0 -> n_step x md_integrate_atomic with a frame + log block every n_output steps.

    python -m reactive_pb_nn_md_b200.run conf.gro ff.pmt topology.top simulation.pmt traj.gro md.log [--ms-evb yes|no] [--seed N]

Only the NVE ensemble of the force path is built (Langevin / MC barostat are out of scope).
The reference seeds its velocity sampler from the clock (general_routines.f90:726-737); here the seed is explicit.
"""
import argparse
import sys

import numpy as np

from . import engine, inputs, tables
from .forcefield import load_forcefield


def sample_atomic_velocities(system, temperature, rng):
    """Maxwell-Boltzmann at `temperature`, then the centre-of-mass momentum removed (general_routines.f90:700-790)."""
    sigma = np.sqrt(tables.BOLTZMANN * temperature / system.mass * tables.CONV_KJMOL)
    v = rng.normal(size=(system.n_atoms, 3)) * sigma[:, None]
    v -= (system.mass[:, None] * v).sum(axis=0) / system.mass.sum()
    return np.ascontiguousarray(v)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("gro"); ap.add_argument("pmt"); ap.add_argument("top"); ap.add_argument("simpmt")
    ap.add_argument("traj"); ap.add_argument("log")
    ap.add_argument("--ms-evb", default="yes", choices=["yes", "no"], help="glob_v.f90:45 ms_evb_simulation")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--library", default=None, help="path of a library exporting the rpbmd C-ABI (default: the CUDA library)")
    a = ap.parse_args(argv)
    sp = inputs.read_simulation_parameters(open(a.simpmt).read())
    if sp["ensemble"] != "NVE":
        raise SystemExit("only the NVE ensemble is on the built path (ensemble = %s)" % sp["ensemble"])
    ff = load_forcefield(open(a.pmt).read(), open(a.top).read(), sp.get("lj_comb_rule", "standard"), sp.get("n_exclusions", 3))
    system = inputs.system_from_gro(ff, open(a.gro).read())
    rng = np.random.default_rng(a.seed)
    system.velocity = sample_atomic_velocities(system, sp.get("initial_temp", sp.get("temperature", 300.0)), rng)
    evb = a.ms_evb == "yes" and ff.has_evb and system.hydronium_mol > 0
    lib = None
    if a.library:
        from ._binding import Library
        lib = Library(a.library)
    sim = engine.Simulation(system, inputs.force_path_parameters(sp), library=lib)
    force = sim.ms_evb_calculate_total_force_energy if evb else sim.calculate_total_force_energy
    force()                                                       # initialize_energy_force
    with open(a.traj, "w") as traj, open(a.log, "w") as log:
        def print_step(i_step):
            st, en = sim.download_state(), sim.energies()
            inputs.write_gro_frame(traj, i_step, i_step * sp["delta_t"], ff, st, system.box_length)
            inputs.write_log_step(log, i_step, i_step * sp["delta_t"], en, evb)
        print_step(0)
        done = 0
        while done < sp["n_step"]:
            n = min(sp["n_output"], sp["n_step"] - done)
            sim.md_integrate_atomic(n, ms_evb=evb)                # n_output steps on the device between two frames
            done += n
            if done % sp["n_output"] == 0:
                print_step(done)
    return 0


if __name__ == "__main__":
    sys.exit(main())
