"""Host-side readers / writers for the reference's input deck and output files, so that an unmodified deck
(conf.gro, ff.pmt, topology.top, simulation parameters) runs through this package.
Formats follow the reference's own read / write statements:

  read_gro                      src/general_routines.f90:215-313   '(I5,2A5,I5,3F8.3)', nm -> Angstrom, 3- or 9-number box line
  read_simulation_parameters    src/read_simulation_parameters.f90:46-146 (list-directed "name value" pairs in two blocks)
  write_gro_frame / write_log   src/general_routines.f90:870-945    (print_step, print_gro_file)
  write_velocity_checkpoint     src/general_routines.f90:997-1026   (print_velocities_checkpoint, '(I5,2A5,I5,3F14.6)')
  check_restart_trajectory      src/general_routines.f90:37-115     (restart iff trajectory, log and velocity files exist
                                                                     and end at the same step > 0)
  last_gro_frame / read_velocity_restart_checkpoint   src/general_routines.f90:120-178

Input parsing only -- no force-path arithmetic lives here.
"""
import numpy as np

from .engine import SimulationParameters
from .system import System


def read_gro(text):
    """-> (molecule names [M], atom names [N], xyz [N,3] in Angstrom, box [3,3] in Angstrom, mol_n_atom [M]).
    A new molecule starts whenever the residue number changes (general_routines.f90:264-282)."""
    lines = text.splitlines()
    n = int(lines[1].split()[0])
    mol_names, n_atom, anames = [], [], []
    xyz = np.zeros((n, 3))
    prev = None
    for i in range(n):
        ln = lines[2 + i]
        i_mole, mname, aname = int(ln[0:5]), ln[5:10].strip(), ln[10:15].strip()
        xyz[i] = [float(ln[20:28]) * 10.0, float(ln[28:36]) * 10.0, float(ln[36:44]) * 10.0]
        if i_mole != prev:
            mol_names.append(mname); n_atom.append(0); prev = i_mole
        n_atom[-1] += 1
        anames.append(aname)
    args = lines[2 + n].split()
    box = np.zeros((3, 3), order="F")
    if len(args) == 3:
        for k in range(3):
            box[k, k] = float(args[k])
    elif len(args) == 9:
        v = [float(a) for a in args]
        box[0, 0], box[1, 1], box[2, 2], box[0, 1], box[0, 2], box[1, 0], box[1, 2], box[2, 0], box[2, 1] = v
    else:
        raise ValueError("error reading box in read_trajectory_snapshot subroutine")
    box *= 10.0
    return mol_names, anames, xyz, box, np.array(n_atom, np.int32)


def system_from_gro(ff, text, velocity=None):
    """initialize_simulation's reading part: molecules by name from the topology's [ moleculetype ] blocks; the atom
    count per molecule must match the molecule type (the reference stops in gen_param / read_topology otherwise)."""
    mol_names, _, xyz, box, n_atom = read_gro(text)
    off = [box[i, j] for i in range(3) for j in range(3) if i != j]
    if max(abs(v) for v in off) > 10e-6:
        raise ValueError("code has been modified to assume orthorhombic box")        # main_ms_evb.f90:62-68
    if abs(box[0, 0] - box[1, 1]) > 1e-9 or abs(box[0, 0] - box[2, 2]) > 1e-9:
        raise ValueError("cubic box required: the reference's CB_array uses box(1,1)**3 as the volume (pme.f90:646)")
    for name, n in zip(mol_names, n_atom):
        if ff.molecule_types[ff.mtype(name) - 1].n_atom != n:
            raise ValueError("molecule %s: %d atoms in the .gro file, %d in the topology" % (name, n, ff.molecule_types[ff.mtype(name) - 1].n_atom))
    return System(ff, box[0, 0], mol_names, xyz, velocity)


_METHOD_KEYS = ("ensemble", "lj_comb_rule", "grid_Tang_Toennies")
_NUMBER_KEYS = {"n_step": int, "n_output": int, "n_exclusions": int, "checkpoint_velocity": int, "temperature": float,
                "initial_temp": float, "friction_coeff": float, "pressure": float, "barofreq": int, "baroscale": float,
                "delta_t": float, "real_space_cutoff": float, "na_nslist": int, "nb_nslist": int, "nc_nslist": int,
                "verlet_cutoff": float, "alpha_sqrt": float, "pme_grid": int, "spline_order": int, "n_threads": int, "debug": int}


def read_simulation_parameters(text):
    """-> dict.  Two blocks: after the line containing 'Simulation Methodology' come "name string" pairs until the line
    whose two tokens contain 'Simulation' and 'Param'; then "name number" pairs (integers through NINT)."""
    out = {}
    lines = iter(text.splitlines())
    for ln in lines:
        if "Simulation Methodology" in ln:
            break
    for ln in lines:
        tok = ln.split()
        if len(tok) < 2:
            continue
        if "Simulation" in tok[0] and "Param" in tok[1]:
            break
        if tok[0] in _METHOD_KEYS:
            out[tok[0]] = tok[1].strip("'\"")
    for ln in lines:
        tok = ln.split()
        if len(tok) < 2 or tok[0] not in _NUMBER_KEYS:
            continue
        val = float(tok[1].replace("d", "e").replace("D", "e"))
        out[tok[0]] = int(round(val)) if _NUMBER_KEYS[tok[0]] is int else val
    missing = [k for k in ("ensemble", "n_step", "n_output", "delta_t", "real_space_cutoff", "verlet_cutoff", "alpha_sqrt", "pme_grid",
                           "spline_order") if k not in out]
    if missing:
        raise ValueError("simulation parameter file: missing " + ", ".join(missing))
    return out


def force_path_parameters(sp):
    """the subset the force path consumes -> engine.SimulationParameters"""
    return SimulationParameters(delta_t=sp["delta_t"], real_space_cutoff=sp["real_space_cutoff"], verlet_cutoff=sp["verlet_cutoff"],
                                na_nslist=sp.get("na_nslist", 10), nb_nslist=sp.get("nb_nslist", 10), nc_nslist=sp.get("nc_nslist", 10),
                                alpha_sqrt=sp["alpha_sqrt"], pme_grid=sp["pme_grid"], spline_order=sp["spline_order"],
                                n_threads=sp.get("n_threads", 1))


def write_gro_frame(fh, i_step, time_ps, ff, state, box_length, atom_names=None):
    """print_gro_file (general_routines.f90:906-945): molecule names follow the CURRENT molecule types (a hop renames
    donor and acceptor), coordinates in nm, '(I5,2A5,I5,3F8.3)' records, '(9F7.4)' box line."""
    n = len(state["xyz"])
    fh.write(" step  %11d time(ps) %24.16f\n" % (i_step, time_ps))
    fh.write(" %11d\n" % n)
    count = 1
    for m, (f, na, t) in enumerate(zip(state["mol_first_atom"], state["mol_n_atom"], state["mol_type"])):
        mt = ff.molecule_types[t - 1]
        for a in range(na):
            g = f - 1 + a
            aname = atom_names[g] if atom_names is not None else ff.atype_name[state["atom_type"][g] - 1]
            x = state["xyz"][g] / 10.0
            fh.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % ((m + 1) % 100000, mt.name[:5], aname[:5], count % 100000, x[0], x[1], x[2]))
            count += 1
    L = box_length / 10.0
    fh.write("%7.4f%7.4f%7.4f%7.4f%7.4f%7.4f%7.4f%7.4f%7.4f\n" % (L, L, L, 0, 0, 0, 0, 0, 0))


def write_log_step(fh, i_step, time_ps, energies, ms_evb):
    """print_step's log block (general_routines.f90:886-899)"""
    fh.write(" i_step , time(ps), potential energy (kJ/mol), kinetic energy (kJ/mol)\n")
    fh.write("%9d%10.3f%16.6E%16.6E\n" % (i_step, time_ps, energies["potential_energy"], energies["kinetic_energy"]))
    if not ms_evb:
        fh.write(" Electrostatic ,   VDWs ,   Bond   ,   Angle  ,  Dihedral\n")
        fh.write("%16.6E%16.6E%16.6E%16.6E%16.6E\n" % tuple(energies[k] for k in ("E_elec", "E_vdw", "E_bond", "E_angle", "E_dihedral")))
    fh.write(" ------------------------------\n")


def write_velocity_checkpoint(fh, i_step, ff, state, atom_names=None):
    """print_velocities_checkpoint (general_routines.f90:997-1026): a ' step <i>' heading, then one
    '(I5,2A5,I5,3F14.6)' record per atom -- molecule index, molecule name, atom name, atom index WITHIN the molecule
    (the trajectory file counts atoms globally, this file does not), velocity in Angstrom/ps."""
    fh.write(" step  %11d\n" % i_step)
    for m, (f, na, t) in enumerate(zip(state["mol_first_atom"], state["mol_n_atom"], state["mol_type"])):
        mt = ff.molecule_types[t - 1]
        for a in range(na):
            g = f - 1 + a
            aname = atom_names[g] if atom_names is not None else ff.atype_name[state["atom_type"][g] - 1]
            v = state["velocity"][g]
            fh.write("%5d%-5s%5s%5d%14.6f%14.6f%14.6f\n" % ((m + 1) % 100000, mt.name[:5], aname[:5], (a + 1) % 100000, v[0], v[1], v[2]))


def _step_headings(text):
    """(line number, step) of every heading line whose first token is 'step' (read_file_find_heading + parse)"""
    out = []
    for k, ln in enumerate(text.splitlines()):
        tok = ln.split()
        if tok and tok[0] == "step":
            out.append((k, int(tok[1])))
    return out


def check_restart_trajectory(traj_path, log_path, velocity_path):
    """-> n_old_trajectory (0: a new run).  The reference restarts iff all three files exist (general_routines.f90:48-52);
    the last step of the trajectory and of the velocity checkpoint must then be the same positive number, anything else is
    its 'error restarting trajectory' stop (:86-97)."""
    import os
    if not (os.path.exists(traj_path) and os.path.exists(log_path) and os.path.exists(velocity_path)):
        return 0
    traj = _step_headings(open(traj_path).read())
    vel = _step_headings(open(velocity_path).read())
    i_traj = traj[-1][1] if traj else 0
    i_vel = vel[-1][1] if vel else 0
    if i_traj != i_vel or i_vel <= 0:
        raise ValueError("error restarting trajectory.  last step is not the same in the trajectory output and velocity checkpointing files")
    return i_traj


def last_gro_frame(traj_text, n_old_trajectory):
    """scan_grofile_restart (general_routines.f90:120-143): the text of the frame of step n_old_trajectory, headed by its
    'step' line like a .gro title line (read_gro then reads it like the input configuration -- positions as printed,
    F8.3 nm: a restart is lossy to 1e-3 nm, in the reference too)."""
    lines = traj_text.splitlines()
    for k, step in _step_headings(traj_text):
        if step == n_old_trajectory:
            n = int(lines[k + 1].split()[0])
            return "\n".join(lines[k:k + n + 3]) + "\n"
    raise ValueError("step %d not found in the trajectory file" % n_old_trajectory)


def read_velocity_restart_checkpoint(velocity_text, n_old_trajectory, n_atoms):
    """read_velocity_restart_checkpoint (general_routines.f90:148-178): the block of step n_old_trajectory -> [N,3]"""
    lines = velocity_text.splitlines()
    for k, step in _step_headings(velocity_text):
        if step == n_old_trajectory:
            v = np.zeros((n_atoms, 3))
            for i in range(n_atoms):
                ln = lines[k + 1 + i]
                v[i] = [float(ln[20:34]), float(ln[34:48]), float(ln[48:62])]
            return v
    raise ValueError("step %d not found in the velocity checkpoint file" % n_old_trajectory)
