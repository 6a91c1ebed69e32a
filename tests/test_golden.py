"""Committed golden vectors (tools/make_golden.py, generated from the single-threaded oracle).  The oracle is re-checked
against them on CPU; the CUDA library is checked against them on the GPU (-m gpu)."""
import os

import numpy as np
import pytest

from reactive_pb_nn_md_b200 import engine
from tests.util import E_RTOL, F_RTOL, small_params, water_system

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EKEYS = ("potential_energy", "E_elec", "E_vdw", "E_bond", "E_angle", "E_dihedral", "E_recip")


def check_nonreactive(sim, etol, ftol, exact_list):
    g = np.load(os.path.join(GOLD, "water1000_nonreactive.npz"))
    sim.calculate_total_force_energy()
    e = sim.energies()
    scale = np.abs(g["energies"]).max()
    for k, ge in zip(EKEYS, g["energies"]):
        assert abs(e[k] - ge) <= etol * max(abs(ge), scale), k
    f = sim.forces()
    assert np.abs(f[:96] - g["force_head"]).max() <= ftol * np.abs(g["force_head"]).max()
    assert abs((f ** 2).sum() - g["force_sq_sum"]) <= 2 * ftol * g["force_sq_sum"]
    vp, nl, _ = sim.neighbor_list()
    assert len(nl) == int(g["n_pairs"]) and np.array_equal(vp, g["verlet_point"])
    if exact_list:
        n64 = nl.astype(np.int64)
        assert [int(n64.sum()), int((n64 * (np.arange(len(nl)) % 1009 + 1)).sum())] == list(g["nl_checksum"])
    Q, th, fr = sim.pme()
    assert abs(Q.sum() - g["Q_sum"]) < 1e-9 and abs(np.abs(Q).sum() - g["Q_abs_sum"]) < 1e-9 * g["Q_abs_sum"]
    assert abs(np.abs(th).sum() - g["theta_abs_sum"]) < 1e-10 * g["theta_abs_sum"]
    assert np.abs(fr[:96] - g["force_recip_head"]).max() <= ftol * np.abs(g["force_recip_head"]).max()


def check_msevb(sim, etol, ftol):
    g = np.load(os.path.join(GOLD, "h3o_water999_msevb.npz"))
    sim.ms_evb_calculate_total_force_energy()
    ev = sim.evb()
    assert ev["n_states"] == int(g["n_states"]) and np.array_equal(ev["proton_log"], g["proton_log"])
    assert np.array_equal(ev["coupling_matrix"], g["coupling_matrix"]) and ev["principal_diabat"] == int(g["principal_diabat"])
    scale = np.abs(np.diag(g["hamiltonian"])).max()
    assert np.abs(ev["hamiltonian"] - g["hamiltonian"]).max() <= etol * scale
    assert abs(ev["adiabatic_potential"] - float(g["adiabatic_potential"])) <= etol * abs(float(g["adiabatic_potential"]))
    c, cg = ev["eigenvector"], g["eigenvector"]
    assert np.abs(c * np.sign(np.dot(c, cg)) - cg).max() < 1e-9
    f = sim.forces()
    assert np.abs(f[:96] - g["force_head"]).max() <= ftol * np.abs(g["force_head"]).max()
    assert abs((f ** 2).sum() - g["force_sq_sum"]) <= 2 * ftol * g["force_sq_sum"]


def test_oracle_matches_golden_nonreactive(oracle_lib):
    check_nonreactive(engine.Simulation(water_system(10), small_params(), library=oracle_lib), 1e-12, 1e-11, True)


def test_oracle_matches_golden_msevb(oracle_lib):
    check_msevb(engine.Simulation(water_system(10, hydronium=True), small_params(), library=oracle_lib), 1e-12, 1e-11)


@pytest.mark.gpu
def test_cuda_matches_golden_nonreactive(cuda_lib):
    check_nonreactive(engine.Simulation(water_system(10), small_params(), library=cuda_lib), E_RTOL, F_RTOL, True)


@pytest.mark.gpu
def test_cuda_matches_golden_msevb(cuda_lib):
    check_msevb(engine.Simulation(water_system(10, hydronium=True), small_params(), library=cuda_lib), E_RTOL, F_RTOL)


def check_acid_ion_pair(sim, etol, ftol):
    """BASELINE config 1 written as CH3SO3- + H3O+: enumeration, Hamiltonian, the committed hop (re-ordered sulfonate)
    and the force array AFTER the commit -- which carries the reference's back-mapping quirk (ms_evb.f90:2608-2656)"""
    g = np.load(os.path.join(GOLD, "acid_ion_pair_msevb.npz"))
    sim.ms_evb_calculate_total_force_energy()
    ev, st, e = sim.evb(), sim.download_state(), sim.energies()
    assert ev["n_states"] == int(g["n_states"]) and np.array_equal(ev["proton_log"], g["proton_log"])
    assert ev["principal_diabat"] == int(g["principal_diabat"]) == 2 and ev["new_hydronium_mol"] == int(g["new_hydronium_mol"]) == 1
    scale = max(np.abs(np.diag(g["hamiltonian"])).max(), abs(float(g["energies"][0])))
    assert np.abs(ev["hamiltonian"] - g["hamiltonian"]).max() <= etol * scale
    assert abs(ev["adiabatic_potential"] - float(g["adiabatic_potential"])) <= etol * scale
    for k, name in enumerate(("E_elec", "E_vdw", "E_bond", "E_angle", "E_dihedral", "E_recip")):
        assert abs(e[name] - float(g["energies"][k])) <= etol * scale, name
    assert np.array_equal(st["atom_type"][:12], g["atom_type_head"]) and np.array_equal(st["mol_type"][:4], g["mol_type_head"])
    assert np.array_equal(st["mol_n_atom"][:4], g["mol_n_atom_head"])
    assert np.abs(st["xyz"][:12] - g["xyz_head"]).max() < 1e-12
    assert np.abs(st["force"][:24] - g["force_head"]).max() <= ftol * np.abs(g["force_head"]).max()
    assert abs((st["force"] ** 2).sum() - g["force_sq_sum"]) <= 2 * ftol * g["force_sq_sum"]


def test_oracle_matches_golden_acid_ion_pair(oracle_lib):
    from reactive_pb_nn_md_b200 import system
    check_acid_ion_pair(engine.Simulation(system.build_acid_box(10, ion_pair=True), small_params(), library=oracle_lib), 1e-12, 1e-11)


@pytest.mark.gpu
def test_cuda_matches_golden_acid_ion_pair(cuda_lib):
    from reactive_pb_nn_md_b200 import system
    check_acid_ion_pair(engine.Simulation(system.build_acid_box(10, ion_pair=True), small_params(), library=cuda_lib), E_RTOL, F_RTOL)
