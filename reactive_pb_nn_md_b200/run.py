"""Driver that runs an unmodified reference input deck on the CUDA library, in the order of main_ms_evb
(src/main_ms_evb.f90:20-118): read simulation parameters -> read .gro/.pmt/.top -> tables -> initial force ->
velocities -> print step 0 -> n_step x md_integrate_atomic with a frame + log block every n_output steps.

    python -m reactive_pb_nn_md_b200.run conf.gro ff.pmt topology.top simulation.pmt traj.gro md.log [--ms-evb yes|no] [--seed N]

Checkpoint / restart as in the reference (general_routines.f90:37-178, 997-1026; main_ms_evb.f90:40-118): with
`checkpoint_velocity n` in the simulation parameters the velocities go to the file `velocity_checkpoint` (next to the
trajectory) every n steps; when trajectory, log and velocity files exist and end at the same step, the run continues from
there -- positions from the last trajectory frame, velocities from the checkpoint, output appended, step numbers carried on.

Only the NVE ensemble of the force path is built (Langevin / MC barostat are out of scope).
The reference seeds its velocity sampler from the clock (general_routines.f90:726-737); here the seed is explicit.
"""
import argparse
import os
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
    ap.add_argument("--velocity-file", default=None, help="glob_v.f90:401 ifile_velocity (default: velocity_checkpoint next to the trajectory)")
    a = ap.parse_args(argv)
    vel_path = a.velocity_file or os.path.join(os.path.dirname(os.path.abspath(a.traj)), "velocity_checkpoint")
    n_old = inputs.check_restart_trajectory(a.traj, a.log, vel_path)             # main_ms_evb.f90:40
    sp = inputs.read_simulation_parameters(open(a.simpmt).read())
    if sp["ensemble"] != "NVE":
        raise SystemExit("only the NVE ensemble is on the built path (ensemble = %s)" % sp["ensemble"])
    n_step_velocity = sp.get("checkpoint_velocity", 0)
    if n_old and not n_step_velocity:                                            # read_simulation_parameters.f90:226-236
        raise SystemExit("if continuing a trajectory, must have the number of steps for velocity checkpointing set in the simulation parameters file")
    ff = load_forcefield(open(a.pmt).read(), open(a.top).read(), sp.get("lj_comb_rule", "standard"), sp.get("n_exclusions", 3))
    if n_old:      # initialize_routines.f90:47-72: last frame of the old trajectory, velocities of the same step
        system = inputs.system_from_gro(ff, inputs.last_gro_frame(open(a.traj).read(), n_old))
        system.velocity = inputs.read_velocity_restart_checkpoint(open(vel_path).read(), n_old, system.n_atoms)
    else:
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
    mode = "a" if n_old else "w"                                  # a continuation appends (general_routines.f90:99-111)
    vel = open(vel_path, mode) if n_step_velocity else None
    with open(a.traj, mode) as traj, open(a.log, mode) as log:
        def print_step(i_step):
            st, en = sim.download_state(), sim.energies()
            inputs.write_gro_frame(traj, i_step, i_step * sp["delta_t"], ff, st, system.box_length)
            inputs.write_log_step(log, i_step, i_step * sp["delta_t"], en, evb)
        if not n_old:
            print_step(0)
        # main_ms_evb.f90:100-118: i_step counts from the restart, trajectory_step from the beginning; frames every
        # n_output and checkpoints every n_step_velocity steps OF THIS RUN.  The steps in between stay on the device.
        n_run = sp["n_step"] - n_old
        done = 0
        while done < n_run:
            nxt = min(n_run, (done // sp["n_output"] + 1) * sp["n_output"])
            if n_step_velocity:
                nxt = min(nxt, (done // n_step_velocity + 1) * n_step_velocity)
            sim.md_integrate_atomic(nxt - done, ms_evb=evb)
            done = nxt
            if done % sp["n_output"] == 0:
                print_step(n_old + done)
            if n_step_velocity and done % n_step_velocity == 0:
                inputs.write_velocity_checkpoint(vel, n_old + done, ff, sim.download_state())
    if vel:
        vel.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
