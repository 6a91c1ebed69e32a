"""Host-side setup (parsers, parameter generation, look-up tables) against hand-checked values."""
import numpy as np

from reactive_pb_nn_md_b200 import system, tables
from reactive_pb_nn_md_b200.forcefield import load_forcefield


def analytic_bspline(u, n):
    """cardinal B-spline M_n(u) by the exact recursion in float64"""
    u = np.asarray(u, float)
    if n == 2:
        return np.where((u < 0) | (u > 2), 0.0, 1.0 - np.abs(u - 1.0))
    return u / (n - 1) * analytic_bspline(u, n - 1) + (n - u) / (n - 1) * analytic_bspline(u - 1.0, n - 1)


def test_reference_constants_are_single_precision_literals():
    assert tables.CONV_E2A_KJMOL == 1389.3546142578125        # glob_v.f90:389 REAL*4 literal
    assert tables.SAFE_VERLET == 1.2000000476837158           # glob_v.f90:393
    assert tables.PI32 == 3.1415927410125732                  # pme.f90:542


def test_spline_tables_float32_storage():
    B6, B5 = tables.spline_tables()
    assert B6.shape == (100000,) and B5.shape == (100000,)
    assert np.array_equal(B6, B6.astype(np.float32).astype(np.float64))     # values are float32-representable (pme.f90:509)
    u = 6.0 / 1e5 * np.arange(1, 100001)
    err = np.abs(B6 - analytic_bspline(u, 6)).max()
    assert 1e-8 < err < 5e-5                                                 # SURVEY: up to 2.7e-5 vs analytic
    assert np.abs(B5 - analytic_bspline(5.0 / 1e5 * np.arange(1, 100001), 5)).max() < 5e-5
    # partition of unity of the tabulated weights at a few arguments (error ~1e-5)
    for frac in (0.1, 0.37, 0.93):
        idx = np.ceil((frac + np.arange(6)) / 6.0 * 1e5).astype(int)
        assert abs(B6[idx - 1].sum() - 1.0) < 5e-5


def test_ewald_tables_shifted_grid():
    dx, T, Sc = tables.ewald_tables(10.0, 0.3)
    assert dx == 10.0 / 100000 and T.shape == (100001,)
    from math import erfc, exp
    i = 50000
    x = (i * dx) * 0.3
    assert abs(T[i - 1] - erfc(x) * tables.CONV_E2A_KJMOL) < 1e-9           # T(i) holds r = i*dx (initialize_routines.f90:237-242)
    assert abs(Sc[i - 1] - (T[i - 1] + x * 2.0 / tables.PI_SQRT * exp(-x * x) * tables.CONV_E2A_KJMOL)) < 1e-9


def test_tang_toennies_tables():
    tt, dtt = tables.tang_toennies_tables()
    assert tt.shape == (4, 1000)
    assert tt[0, -1] > 0.999999 and tt[3, -1] > 0.99 and (np.diff(tt, axis=1) > -1e-15).all()
    assert (dtt >= 0).all()


def test_cb_array_structure():
    K = 16
    CB = tables.cb_array(31.07, K, 0.3)
    assert CB.shape == (K, K, K) and CB[0, 0, 0] == 0.0 and (CB.flatten()[1:] > 0).all()
    # even up to the float32 rounding of the phase factors (pme.f90:589)
    idx = (-np.arange(K)) % K
    rel = np.abs(CB - CB[np.ix_(idx, idx, idx)]) / np.maximum(CB, 1e-300)
    assert rel.max() < 1e-5
    assert np.allclose(CB, CB.transpose(1, 0, 2), rtol=1e-14)


def test_forcefield_parameters():
    ff = system.example_forcefield()
    t = ff.atype
    # opls: C12 = 4 eps sigma^12, C6 = 4 eps sigma^6; cross terms geometric means (initialize_routines.f90:598-634)
    eps, sig = 0.6502995, 3.16549
    assert np.isclose(ff.vdw_parameter[t("OW") - 1, t("OW") - 1, 0], 4 * eps * sig ** 12, rtol=1e-14)
    assert np.isclose(ff.vdw_parameter[t("OW") - 1, t("OW") - 1, 1], 4 * eps * sig ** 6, rtol=1e-14)
    c12a = ff.vdw_parameter[t("O_a") - 1, t("O_a") - 1, 0]
    assert np.isclose(ff.vdw_parameter[t("OW") - 1, t("O_a") - 1, 0], np.sqrt(c12a * 4 * eps * sig ** 12), rtol=1e-14)
    # explicit cross terms are stored (C12, C6) and typed LJ (initialize_routines.f90:390-396, 488-491)
    assert ff.vdw_parameter[t("O_h3o") - 1, t("OW") - 1, 0] == 1917990.0 and ff.vdw_parameter[t("O_h3o") - 1, t("OW") - 1, 1] == 1993.468
    assert ff.vdw_type[t("O_h3o") - 1, t("OW") - 1] == 0
    # eps = 0 types fall into the all-zero SAPT branch (type 1), B = 3.0 (initialize_routines.f90:302-303, 476-479)
    assert ff.vdw_type[t("HW") - 1, t("OW") - 1] == 1
    assert np.array_equal(ff.vdw_parameter[t("HW") - 1, t("OW") - 1], [0, 3.0, 0, 0, 0, 0])
    assert ff.vdw_type[t("HW") - 1, t("HW") - 1] == -1
    # 1-4 table = copy + pairtypes overrides (initialize_routines.f90:646-691)
    assert ff.vdw_parameter_14[t("H_a") - 1, t("O_a") - 1, 0] == 66466.2 and ff.vdw_parameter_14[t("O_a") - 1, t("H_a") - 1, 1] == 434.1
    # angles converted with the truncated pi
    assert ff.angle_parameter[t("HW") - 1, t("OW") - 1, t("HW") - 1, 0] == 113.24 * 3.141592654 / 180.0


def test_exclusion_generation():
    ff3 = system.example_forcefield(n_exclusions=3)
    for mt in ff3.molecule_types:
        n = mt.n_atom
        assert (mt.pair_exclusions[:n, :n] == 1).all()      # every pair within 3 bonds in these small molecules
    ff2 = system.example_forcefield(n_exclusions=2)
    so3h = ff2.molecule_types[ff2.mtype("so3h") - 1]
    ex = so3h.pair_exclusions
    assert ex[0, 5] == 2 and ex[5, 0] == 2        # C ... H_a : 1-4 pair (C-S-O-H)
    assert ex[2, 5] == 2 and ex[3, 5] == 2        # O_a ... H_a : 1-4
    assert ex[0, 2] == 1 and ex[1, 5] == 1 and ex[4, 5] == 1


def test_evb_topology_tables():
    ff = system.example_forcefield()
    h3o, h2o = ff.mtype("h3o"), ff.mtype("h2o")
    assert ff.evb_conjugate_pairs[h3o - 1] == h2o and ff.evb_conjugate_pairs[h2o - 1] == h3o
    assert ff.evb_proton_index[h3o - 1] == ff.atype("H_h3o") and ff.evb_heavy_acid_index[h3o - 1] == ff.atype("O_h3o")
    # "O_a O_b" comes last so that conj(O_b) = O_a (comment in the reference topology)
    assert ff.evb_conjugate_atom_index[ff.atype("O_b") - 1] == ff.atype("O_a")
    assert ff.evb_conjugate_atom_index[ff.atype("O_ah") - 1] == ff.atype("O_b")
    assert list(ff.molecule_types[h3o - 1].reactive_protons[:4]) == [0, 1, 1, 1]
    assert ff.evb_reference_energy[ff.mtype("so3h") - 1] == -643.65
    assert ff.evb_diabat_coupling_type[0] == 1 and ff.evb_diabat_coupling_parameters[0, 0] == -97.0151921


def test_synthetic_configs_shapes():
    s = system.build_water_box(10, with_hydronium=True)
    assert s.n_atoms == 3001 and s.n_mole == 1000 and s.hydronium_mol == 1
    assert abs(s.charge.sum() - 1.0) < 1e-12
    assert np.array_equal(s.xyz, np.round(s.xyz, 2))           # .gro precision
    p = (s.mass[:, None] * s.velocity).sum(axis=0)
    assert np.abs(p).max() < 1e-9


def test_ensemble_step_and_peer_entry_points_on_the_oracle(oracle_lib):
    """rpb_ensemble_step (replicas only) steps every context; the peer-memory exchange is a device feature that the
    oracle refuses with RPB_ERR_UNSUPPORTED (the engine then keeps the collective path)."""
    import ctypes as C
    import pytest
    from reactive_pb_nn_md_b200 import engine
    from reactive_pb_nn_md_b200._binding import RpbError
    from tests.util import small_params, water_system
    s = water_system(10, hydronium=True)
    p = small_params(pme_grid=32, n_threads=2)
    ens = [engine.Simulation(s, p, library=oracle_lib) for _ in range(2)]
    one = engine.Simulation(s, p, library=oracle_lib)
    for sim in ens + [one]:
        sim.ms_evb_calculate_total_force_energy()
    engine.Simulation.ensemble_step(ens, 2, ms_evb=True)
    one.md_integrate_atomic(2, ms_evb=True)
    for sim in ens:
        assert np.abs(sim.download_state()["xyz"] - one.download_state()["xyz"]).max() < 1e-10     # OpenMP reduction order
    assert one.dll.rpb_peer_enabled(one.ctx) == 0
    buf = C.create_string_buffer(64)
    assert one.dll.rpb_peer_export(one.ctx, buf) == -7
    with pytest.raises(RpbError):
        one._check(one.dll.rpb_peer_import(one.ctx, buf, 1))
