"""Readers / writers for the reference's input deck and the main_ms_evb-ordered driver, in front of the
CPU oracle (no GPU needed): .gro round trip at the format's precision, simulation-parameter parsing, a short run."""
import io
import os

import numpy as np
import pytest

from reactive_pb_nn_md_b200 import inputs, run, system

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SIMPMT = """Simulation Methodology
ensemble            NVE
lj_comb_rule        opls
grid_Tang_Toennies  yes
Simulation Parameters
n_step              4
n_output            2
n_exclusions        3
temperature         300.0
delta_t             0.0005
real_space_cutoff   10.0
verlet_cutoff       12.0
na_nslist           10
nb_nslist           10
nc_nslist           10
alpha_sqrt          0.3
pme_grid            32
spline_order        6
n_threads           2
"""


def _gro_text(s):
    st = dict(xyz=s.xyz, mol_first_atom=s.mol_first_atom, mol_n_atom=s.mol_n_atom, mol_type=s.mol_type, atom_type=s.atom_type)
    buf = io.StringIO()
    inputs.write_gro_frame(buf, 0, 0.0, s.ff, st, s.box_length)
    return buf.getvalue()


def test_gro_round_trip_and_molecule_table():
    s = system.build_acid_box(10)
    text = _gro_text(s)
    names, anames, xyz, box, n_atom = inputs.read_gro(text)
    assert names[0] == "so3h" and names[1] == "h2o" and len(names) == s.n_mole
    assert np.array_equal(n_atom, s.mol_n_atom) and anames[:6] == ["C_a", "S_a", "O_a", "O_a", "O_ah", "H_a"]
    assert np.abs(xyz - s.xyz).max() <= 0.5e-2 + 1e-12                  # F8.3 in nm
    assert abs(box[0, 0] - s.box_length) <= 0.5e-3 + 1e-12 and box[0, 1] == 0.0      # F7.4 in nm
    s2 = inputs.system_from_gro(s.ff, text)
    assert np.array_equal(s2.mol_type, s.mol_type) and np.array_equal(s2.atom_type, s.atom_type) and s2.hydronium_mol == 1
    with pytest.raises(ValueError):
        inputs.system_from_gro(s.ff, text.replace(text.splitlines()[-1], "   3.1070   3.1070   3.1070   0.1000   0.0000   0.0000   0.0000   0.0000   0.0000"))


def test_simulation_parameter_file():
    sp = inputs.read_simulation_parameters(SIMPMT)
    assert sp["ensemble"] == "NVE" and sp["lj_comb_rule"] == "opls" and sp["n_step"] == 4 and sp["pme_grid"] == 32
    assert sp["delta_t"] == 0.0005 and sp["alpha_sqrt"] == 0.3
    p = inputs.force_path_parameters(sp)
    assert p.pme_grid == 32 and p.verlet_cutoff == 12.0 and p.n_threads == 2
    with pytest.raises(ValueError):
        inputs.read_simulation_parameters(SIMPMT.replace("pme_grid            32\n", ""))


def test_driver_runs_a_reference_deck_on_the_oracle(tmp_path, oracle_lib):
    s = system.build_water_box(10, with_hydronium=True)
    data = os.path.join(ROOT, "reactive_pb_nn_md_b200", "data")
    (tmp_path / "conf.gro").write_text(_gro_text(s))
    (tmp_path / "sim.pmt").write_text(SIMPMT)
    rc = run.main([str(tmp_path / "conf.gro"), os.path.join(data, "CH3SO3H.pmt"), os.path.join(data, "CH3SO3H_H2O.top"),
                   str(tmp_path / "sim.pmt"), str(tmp_path / "traj.gro"), str(tmp_path / "md.log"), "--library", oracle_lib.path])
    assert rc == 0
    log = (tmp_path / "md.log").read_text().splitlines()
    rows = [ln for ln in log if ln.strip() and ln.strip()[0].isdigit()]
    assert [int(r.split()[0]) for r in rows] == [0, 2, 4]                # step 0 + every n_output
    frames = (tmp_path / "traj.gro").read_text().count("step")
    assert frames == 3
    names, _, xyz, box, _ = inputs.read_gro((tmp_path / "traj.gro").read_text())     # first frame parses back
    assert names[0] == "h3o" and len(xyz) == s.n_atoms


def _deck(tmp_path, s, simpmt):
    data = os.path.join(ROOT, "reactive_pb_nn_md_b200", "data")
    (tmp_path / "conf.gro").write_text(_gro_text(s))
    (tmp_path / "sim.pmt").write_text(simpmt)
    return [str(tmp_path / "conf.gro"), os.path.join(data, "CH3SO3H.pmt"), os.path.join(data, "CH3SO3H_H2O.top"),
            str(tmp_path / "sim.pmt"), str(tmp_path / "traj.gro"), str(tmp_path / "md.log")]


def test_velocity_checkpoint_format_round_trip():
    """print_velocities_checkpoint / read_velocity_restart_checkpoint (general_routines.f90:148-178, 997-1026)"""
    s = system.build_water_box(10, with_hydronium=True)
    rng = np.random.default_rng(3)
    v = rng.normal(size=(s.n_atoms, 3)) * 5.0
    st = dict(xyz=s.xyz, velocity=v, mol_first_atom=s.mol_first_atom, mol_n_atom=s.mol_n_atom, mol_type=s.mol_type, atom_type=s.atom_type)
    buf = io.StringIO()
    inputs.write_velocity_checkpoint(buf, 2, s.ff, st)
    inputs.write_velocity_checkpoint(buf, 4, s.ff, dict(st, velocity=2.0 * v))
    text = buf.getvalue()
    lines = text.splitlines()
    assert lines[0].split() == ["step", "2"] and len(lines) == 2 * (s.n_atoms + 1)
    assert len(lines[1]) == 5 + 5 + 5 + 5 + 3 * 14 and lines[1][5:10].strip() == "h3o"
    assert [int(lines[1 + a][15:20]) for a in range(5)] == [1, 2, 3, 4, 1]            # atom index WITHIN the molecule
    assert np.abs(inputs.read_velocity_restart_checkpoint(text, 2, s.n_atoms) - v).max() <= 0.5e-6
    assert np.abs(inputs.read_velocity_restart_checkpoint(text, 4, s.n_atoms) - 2.0 * v).max() <= 0.5e-6
    with pytest.raises(ValueError):
        inputs.read_velocity_restart_checkpoint(text, 3, s.n_atoms)


def test_restart_detection_rules(tmp_path):
    """check_restart_trajectory (general_routines.f90:37-115)"""
    t, l, v = (str(tmp_path / n) for n in ("traj.gro", "md.log", "velocity_checkpoint"))
    assert inputs.check_restart_trajectory(t, l, v) == 0
    open(t, "w").write(" step  0 time(ps) 0.0\n 1\n    1h2o     Ow    1   0.000   0.000   0.000\n 1 1 1 0 0 0 0 0 0\n step  4 time(ps) 0.002\n 1\n    1h2o     Ow    1   0.000   0.000   0.000\n 1 1 1 0 0 0 0 0 0\n")
    open(l, "w").write("log\n")
    assert inputs.check_restart_trajectory(t, l, v) == 0                 # no velocity file: a new run
    open(v, "w").write(" step  2\n    1h2o     Ow    1      0.000000      0.000000      0.000000\n")
    with pytest.raises(ValueError):
        inputs.check_restart_trajectory(t, l, v)                         # last steps differ: the reference stops
    open(v, "a").write(" step  4\n    1h2o     Ow    1      1.000000      0.000000      0.000000\n")
    assert inputs.check_restart_trajectory(t, l, v) == 4
    assert inputs.last_gro_frame(open(t).read(), 4).splitlines()[0].split()[1] == "4"


def test_driver_checkpoints_and_restarts(tmp_path, oracle_lib):
    """N3: a run of 4 steps that is continued to 8 appends to its files, carries the step numbers on, and follows the
    uninterrupted 8-step run up to the precision the restart files hold (positions F8.3 nm, velocities F14.6 A/ps)."""
    s = system.build_water_box(10, with_hydronium=True)
    base = SIMPMT.replace("n_exclusions        3\n", "n_exclusions        3\ncheckpoint_velocity 2\n")
    a = tmp_path / "a"; b = tmp_path / "b"
    a.mkdir(); b.mkdir()
    lib = ["--library", oracle_lib.path, "--seed", "7"]
    assert run.main(_deck(a, s, base) + lib) == 0                                      # steps 0..4
    vel = (a / "velocity_checkpoint").read_text()
    assert [st for _, st in inputs._step_headings(vel)] == [2, 4]
    assert run.main(_deck(a, s, base.replace("n_step              4", "n_step              8")) + lib) == 0   # continues 5..8
    steps = [st for _, st in inputs._step_headings((a / "traj.gro").read_text())]
    assert steps == [0, 2, 4, 6, 8]
    assert [st for _, st in inputs._step_headings((a / "velocity_checkpoint").read_text())] == [2, 4, 6, 8]
    rows = [ln for ln in (a / "md.log").read_text().splitlines() if ln.strip() and ln.strip()[0].isdigit()]
    assert [int(r.split()[0]) for r in rows] == [0, 2, 4, 6, 8]
    assert run.main(_deck(b, s, base.replace("n_step              4", "n_step              8")) + lib) == 0   # uninterrupted
    fa = inputs.read_gro(inputs.last_gro_frame((a / "traj.gro").read_text(), 8))
    fb = inputs.read_gro(inputs.last_gro_frame((b / "traj.gro").read_text(), 8))
    assert fa[0] == fb[0] and np.abs(fa[2] - fb[2]).max() <= 0.03                    # 1e-3 nm print precision + 4 steps of drift
    va = inputs.read_velocity_restart_checkpoint((a / "velocity_checkpoint").read_text(), 8, s.n_atoms)
    vb = inputs.read_velocity_restart_checkpoint((b / "velocity_checkpoint").read_text(), 8, s.n_atoms)
    # the restart reads positions rounded to 1e-2 A: forces change by O(k * 1e-2 A), velocities by ~1 A/ps over 4 steps
    assert np.sqrt(((va - vb) ** 2).mean()) < 0.15 * np.sqrt((vb ** 2).mean())
    # without checkpoint_velocity in the parameters a continuation is refused (read_simulation_parameters.f90:226-236)
    with pytest.raises(SystemExit):
        run.main(_deck(a, s, SIMPMT.replace("n_step              4", "n_step              12")) + lib)
