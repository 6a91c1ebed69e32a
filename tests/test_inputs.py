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
