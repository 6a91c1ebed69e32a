"""CUDA MS-EVB path vs oracle (BASELINE config 3 at reduced size), through the C-ABI."""
import numpy as np
import pytest

from reactive_pb_nn_md_b200 import engine
from tests.util import E_RTOL, F_RTOL, assert_energies_close, rel_rms, small_params, water_system

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(oracle_lib, cuda_lib):
    s = water_system(10, hydronium=True)
    p = small_params()
    so = engine.Simulation(s, p, library=oracle_lib)
    sg = engine.Simulation(s, p, library=cuda_lib)
    so.ms_evb_calculate_total_force_energy()
    sg.ms_evb_calculate_total_force_energy()
    return sg, so


def test_enumeration_bit_exact(pair):
    sg, so = pair
    eg, eo = sg.evb(), so.evb()
    assert eg["n_states"] == eo["n_states"] and eo["n_states"] >= 8
    assert np.array_equal(eg["proton_log"], eo["proton_log"])       # diabat order, donors, protons, acceptors
    assert np.array_equal(eg["coupling_matrix"], eo["coupling_matrix"])
    assert eg["principal_diabat"] == eo["principal_diabat"]
    assert eg["new_hydronium_mol"] == eo["new_hydronium_mol"]


def test_hamiltonian_elements(pair):
    sg, so = pair
    Hg, Ho = sg.evb()["hamiltonian"], so.evb()["hamiltonian"]
    scale = np.abs(np.diag(Ho)).max()
    assert np.abs(Hg - Ho).max() <= E_RTOL * scale, np.abs(Hg - Ho).max() / scale
    # per-state relative check on the diagonal and the couplings
    d = np.abs(np.diag(Hg) - np.diag(Ho)) / np.maximum(np.abs(np.diag(Ho)), 1.0)
    assert d.max() < E_RTOL
    off = np.triu(Ho, 1) != 0
    assert (np.abs(Hg[off] - Ho[off]) <= 1e-10 * np.maximum(np.abs(Ho[off]), 1.0)).all()


def test_ground_state(pair):
    sg, so = pair
    eg, eo = sg.evb(), so.evb()
    assert abs(eg["adiabatic_potential"] - eo["adiabatic_potential"]) <= E_RTOL * abs(eo["adiabatic_potential"])
    # eigenvector up to a global sign (only products c_i c_j and |c_i| are used, ms_evb.f90:292-319)
    cg, co = eg["eigenvector"], eo["eigenvector"]
    if np.dot(cg, co) < 0:
        cg = -cg
    assert np.abs(cg - co).max() < 1e-9
    assert abs(sg.energies()["potential_energy"] - so.energies()["potential_energy"]) <= E_RTOL * abs(eo["adiabatic_potential"])


def test_hellmann_feynman_forces(pair):
    sg, so = pair
    assert rel_rms(sg.forces(), so.forces()) < F_RTOL


def test_per_state_forces(pair):
    """diagonal force of every diabat: mix with c = e_s"""
    sg, so = pair
    S = so.evb()["n_states"]
    for s in range(S):
        c = np.zeros(S); c[s] = 1.0
        assert rel_rms(sg.debug_mix_forces(c), so.debug_mix_forces(c)) < F_RTOL, s


def test_coupling_forces(pair):
    """off-diagonal force of diabat s with its parent: c = (e_p + e_s)/sqrt(2) minus the two diagonal halves"""
    sg, so = pair
    ev = so.evb()
    S = ev["n_states"]
    for s in range(1, S):
        p = ev["coupling_matrix"][s] - 1
        c = np.zeros(S); c[s] = c[p] = np.sqrt(0.5)
        fg, fo = sg.debug_mix_forces(c), so.debug_mix_forces(c)
        assert rel_rms(fg, fo) < F_RTOL, s


def test_evb_trajectory(oracle_lib, cuda_lib):
    s = water_system(10, hydronium=True, seed=7)
    p = small_params()
    so = engine.Simulation(s, p, library=oracle_lib)
    sg = engine.Simulation(s, p, library=cuda_lib)
    so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
    so.md_integrate_atomic(10, ms_evb=True); sg.md_integrate_atomic(10, ms_evb=True)
    a, b = sg.download_state(), so.download_state()
    assert a["hydronium_mol"] == b["hydronium_mol"]
    assert np.array_equal(a["atom_type"], b["atom_type"])
    assert np.abs(a["xyz"] - b["xyz"]).max() < 1e-9
    assert np.abs(a["velocity"] - b["velocity"]).max() < 1e-8
    assert abs(sg.energies()["potential_energy"] - so.energies()["potential_energy"]) < 1e-9 * abs(so.energies()["potential_energy"])
