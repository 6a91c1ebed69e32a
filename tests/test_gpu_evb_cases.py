"""CUDA MS-EVB path on the hand-built clusters (hop commit, chain cut, Eigen cation), the state-sharded path
emulated on one GPU, and size-independent properties at BASELINE's full sizes."""
import ctypes as C
import os

import numpy as np
import pytest

from reactive_pb_nn_md_b200 import engine, system
from tests import test_oracle_evb as cases
from tests.util import E_RTOL, F_RTOL, rel_rms, small_params, water_system

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ff():
    return system.example_forcefield()


def compare_state(sg, so):
    a, b = sg.download_state(), so.download_state()
    for k in ("atom_type", "mol_first_atom", "mol_n_atom", "mol_type"):
        assert np.array_equal(a[k], b[k]), k
    assert a["hydronium_mol"] == b["hydronium_mol"]
    assert np.abs(a["xyz"] - b["xyz"]).max() < 1e-12 and np.abs(a["charge"] - b["charge"]).max() == 0.0
    assert rel_rms(a["force"], b["force"]) < F_RTOL
    vg, ng, _ = sg.neighbor_list(); vo, no, _ = so.neighbor_list()
    assert np.array_equal(vg, vo) and np.array_equal(ng, no)
    gi, gj, _ = sg.tile_pairs(); oi, oj, _ = so.tile_pairs()
    assert np.array_equal(gi, oi) and np.array_equal(gj, oj)


@pytest.mark.parametrize("order", ["h3o_first", "h3o_last"])
def test_hop_commit_matches_oracle(oracle_lib, cuda_lib, ff, order):
    c = np.array([15.0, 15.0, 15.0])
    h3o = [c, c + np.array([1.32, 0, 0]), c + np.array([-0.35, 0.93, 0]), c + np.array([-0.35, -0.93, 0])]
    wat = cases.water_at(c + np.array([2.42, 0, 0]), [(0.4, 0.8, 0.4), (0.4, -0.8, 0.4)])
    cases.RNG = np.random.default_rng(5)
    fill = cases.filler_waters(60, c)
    if order == "h3o_first":
        mols, names = [h3o, wat] + fill, ["h3o", "h2o"] + ["h2o"] * len(fill)
    else:
        mols, names = fill[:5] + [wat] + fill[5:] + [h3o], ["h2o"] * (len(fill) + 1) + ["h3o"]
    s = cases.build(ff, mols, names)
    so = engine.Simulation(s, small_params(), library=oracle_lib)
    sg = engine.Simulation(s, small_params(), library=cuda_lib)
    so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
    assert so.evb()["new_hydronium_mol"] != s.hydronium_mol          # the hop really happened
    compare_state(sg, so)
    # keep integrating across the committed topology
    so.md_integrate_atomic(5, ms_evb=True); sg.md_integrate_atomic(5, ms_evb=True)
    compare_state(sg, so)
    assert abs(sg.energies()["potential_energy"] - so.energies()["potential_energy"]) <= 1e-9 * abs(so.energies()["potential_energy"]) + 1e-9


def test_eigen_and_wire_clusters(oracle_lib, cuda_lib, ff):
    base = np.array([8.0, 15.0, 15.0]); step = np.array([2.55, 0, 0])
    h3o = cases.water_at(base, [(1, 0, 0), (-0.4, 0.9, 0.2), (-0.4, -0.9, 0.2)])
    wire = [cases.water_at(base + (k + 1) * step, [(1, 0.15, 0.1), (0.1, 0.6 * (-1) ** k, 0.8)]) for k in range(4)]
    cases.RNG = np.random.default_rng(5)
    fill = cases.filler_waters(60, base + 2 * step, min_dist=11.0)
    s = cases.build(ff, [h3o] + wire + fill, ["h3o"] + ["h2o"] * (4 + len(fill)))
    so = engine.Simulation(s, small_params(), library=oracle_lib)
    sg = engine.Simulation(s, small_params(), library=cuda_lib)
    so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
    eg, eo = sg.evb(), so.evb()
    assert eg["n_states"] == eo["n_states"] == 4 and np.array_equal(eg["proton_log"], eo["proton_log"])
    assert np.abs(eg["hamiltonian"] - eo["hamiltonian"]).max() <= E_RTOL * np.abs(np.diag(eo["hamiltonian"])).max()
    assert rel_rms(sg.forces(), so.forces()) < F_RTOL


def _emulated_allreduce(ranks, which):
    import torch
    torch.cuda.synchronize()
    bufs = [sim._exchange_tensor(which) for sim in ranks]
    total = bufs[0].clone()
    for b in bufs[1:]:
        total += b
    for b in bufs:
        b.copy_(total)
    torch.cuda.synchronize()


def _emulated_force(ranks):
    for which, phase in (("h", "rpb_evb_phase_build"), ("f", "rpb_evb_phase_mix")):
        for sim in ranks:
            sim._check(getattr(sim.dll, phase)(sim.ctx))
        _emulated_allreduce(ranks, which)
    for sim in ranks:
        sim._check(sim.dll.rpb_evb_phase_commit(sim.ctx))


@pytest.mark.parametrize("world", [2, 3, 8, 16])
def test_state_sharding_emulated_on_one_gpu(cuda_lib, oracle_lib, world):
    """`world` contexts (rank r of world) on the same device; the two all-reduces are done by hand on the exchange
    buffers.  world=16 exceeds the number of non-principal diabats of this box, so some ranks own no diabat at all;
    every rank still holds its slice of the principal diabat's pair forces."""
    s = water_system(10, hydronium=True)
    p = small_params()
    ref = engine.Simulation(s, p, library=oracle_lib); ref.ms_evb_calculate_total_force_energy()
    ranks = [engine.Simulation(s, p, library=cuda_lib, rank=r, world_size=world) for r in range(world)]
    _emulated_force(ranks)
    er = ref.energies()
    for sim in ranks:
        assert sim.evb()["n_states"] == ref.evb()["n_states"]
        assert np.array_equal(sim.evb()["proton_log"], ref.evb()["proton_log"])
        assert abs(sim.evb()["adiabatic_potential"] - ref.evb()["adiabatic_potential"]) <= E_RTOL * abs(ref.evb()["adiabatic_potential"])
        assert np.abs(sim.evb()["hamiltonian"] - ref.evb()["hamiltonian"]).max() <= E_RTOL * np.abs(np.diag(ref.evb()["hamiltonian"])).max()
        assert rel_rms(sim.forces(), ref.forces()) < F_RTOL
        assert abs(sim.energies()["potential_energy"] - er["potential_energy"]) <= E_RTOL * abs(er["potential_energy"])
    # a short trajectory through the split step (rpb_step_begin / phases / rpb_step_end) stays on the oracle's
    ref.md_integrate_atomic(4, ms_evb=True)
    for _ in range(4):
        for sim in ranks:
            sim._check(sim.dll.rpb_step_begin(sim.ctx))
        _emulated_force(ranks)
        for sim in ranks:
            sim._check(sim.dll.rpb_step_end(sim.ctx))
    xr = ref.download_state()
    for sim in ranks:
        st = sim.download_state()
        assert st["hydronium_mol"] == xr["hydronium_mol"]
        assert np.abs(st["xyz"] - xr["xyz"]).max() < 1e-9
        assert rel_rms(st["force"], xr["force"]) < F_RTOL


def _peer_worker(rank, world, port, outdir, n_steps):
    """one process per rank and per device, as in production"""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    from reactive_pb_nn_md_b200 import engine as eng
    from reactive_pb_nn_md_b200._binding import load_cuda
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = water_system(10, hydronium=True)
    sim = eng.Simulation(s, small_params(), library=load_cuda(), device=dev, rank=rank, world_size=world,
                         process_group=dist.group.WORLD)
    assert sim.exchange == "peer" and sim.dll.rpb_peer_enabled(sim.ctx) == 1
    sim.ms_evb_calculate_total_force_energy()
    f0, ev0, e0 = sim.forces(), sim.evb(), sim.energies()
    sim.md_integrate_atomic(n_steps, ms_evb=True)
    st = sim.download_state()
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), f0=f0, S=ev0["n_states"], log=ev0["proton_log"],
             H=ev0["hamiltonian"], ad=ev0["adiabatic_potential"], pe=e0["potential_energy"], xyz=st["xyz"],
             vel=st["velocity"], force=st["force"], hyd=st["hydronium_mol"])
    dist.barrier()
    sim.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_memory_exchange_between_processes(oracle_lib, world):
    """The sharded step with the peer-memory all-reduce kernels (kernels_peer.cu), one process per rank: IPC handles
    travel over the process group, rpb_step runs the whole step inside the library, results match the oracle and every
    rank ends with the bit-identical replicated state."""
    import tempfile
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        # kernels of different ranks wait on one another: they must never share a device (B200_PROFILING.md: Xid 109)
        pytest.skip("needs %d GPUs (one per rank)" % world)
    n_steps = 8
    s = water_system(10, hydronium=True)
    ref = engine.Simulation(s, small_params(), library=oracle_lib)
    ref.ms_evb_calculate_total_force_energy()
    er, evr, fr = ref.energies(), ref.evb(), ref.forces()
    ref.md_integrate_atomic(n_steps, ms_evb=True)
    xr = ref.download_state()
    with tempfile.TemporaryDirectory() as d:
        port = 29700 + (os.getpid() % 2000)
        mp.spawn(_peer_worker, args=(world, port, d, n_steps), nprocs=world, join=True)
        z = [np.load(os.path.join(d, "rank%d.npz" % r)) for r in range(world)]
        for q in z:
            assert int(q["S"]) == evr["n_states"] and np.array_equal(q["log"], evr["proton_log"])
            assert abs(float(q["ad"]) - evr["adiabatic_potential"]) <= E_RTOL * abs(evr["adiabatic_potential"])
            assert np.abs(q["H"] - evr["hamiltonian"]).max() <= E_RTOL * np.abs(np.diag(evr["hamiltonian"])).max()
            assert abs(float(q["pe"]) - er["potential_energy"]) <= E_RTOL * abs(er["potential_energy"])
            assert rel_rms(q["f0"], fr) < F_RTOL
            assert int(q["hyd"]) == xr["hydronium_mol"]
            assert np.abs(q["xyz"] - xr["xyz"]).max() < 1e-9
            assert rel_rms(q["force"], xr["force"]) < F_RTOL
            for k in ("xyz", "vel", "force"):
                assert np.array_equal(q[k], z[0][k]), k                    # replicated state: bit-identical


def _compare_full_state(sg, so):
    a, b = sg.download_state(), so.download_state()
    for k in ("atom_type", "mol_first_atom", "mol_n_atom", "mol_type"):
        assert np.array_equal(a[k], b[k]), k
    assert a["hydronium_mol"] == b["hydronium_mol"]
    assert np.abs(a["charge"] - b["charge"]).max() == 0.0 and np.abs(a["mass"] - b["mass"]).max() == 0.0
    assert np.abs(a["xyz"] - b["xyz"]).max() < 1e-9 and np.abs(a["velocity"] - b["velocity"]).max() < 1e-8
    assert rel_rms(a["force"], b["force"]) < F_RTOL


@pytest.mark.parametrize("ion_pair", [False, True])
def test_reference_example_system_acid_in_water(cuda_lib, oracle_lib, ion_pair):
    """BASELINE config 1, the reference's own example input (CH3SO3H in water: 6-site acid with G96 bonds, cosine
    angles, proper + improper dihedrals, three basic oxygens on the conjugate base, coupling type of the acid pair).
    ion_pair=False: the acid as principal diabat; every diabat's own force is compared (c = e_s).
    ion_pair=True: the same geometry written as CH3SO3- + H3O+: the first evaluation commits the hop onto the
    sulfonate, which protonates an oxygen that is not last in the molecule -> reorder_molecule_data_structures."""
    s = system.build_acid_box(10, ion_pair=ion_pair)
    p = small_params()
    so = engine.Simulation(s, p, library=oracle_lib)
    sg = engine.Simulation(s, p, library=cuda_lib)
    vo, lo, _ = so.neighbor_list(); vg, lg, _ = sg.neighbor_list()
    assert np.array_equal(vo, vg) and np.array_equal(lo, lg)
    so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
    eg, eo = sg.evb(), so.evb()
    assert eg["n_states"] == eo["n_states"] >= 7 and np.array_equal(eg["proton_log"], eo["proton_log"])
    assert np.array_equal(eg["coupling_matrix"], eo["coupling_matrix"])
    scale = np.abs(np.diag(eo["hamiltonian"])).max()
    assert np.abs(eg["hamiltonian"] - eo["hamiltonian"]).max() <= E_RTOL * max(scale, abs(so.energies()["E_elec"]))
    assert abs(eg["adiabatic_potential"] - eo["adiabatic_potential"]) <= E_RTOL * max(abs(eo["adiabatic_potential"]), abs(so.energies()["E_elec"]))
    assert eg["principal_diabat"] == eo["principal_diabat"] == (2 if ion_pair else 1)
    assert eg["new_hydronium_mol"] == eo["new_hydronium_mol"]
    assert_en = ("E_elec", "E_vdw", "E_bond", "E_angle", "E_dihedral", "E_recip")
    for k in assert_en:
        assert abs(sg.energies()[k] - so.energies()[k]) <= E_RTOL * max(abs(so.energies()[k]), abs(so.energies()["E_elec"])), k
    if not ion_pair:
        assert rel_rms(sg.forces(), so.forces()) < F_RTOL
        for k in range(eo["n_states"]):
            c = np.zeros(eo["n_states"]); c[k] = 1.0
            assert rel_rms(sg.debug_mix_forces(c), so.debug_mix_forces(c)) < F_RTOL, k
    _compare_full_state(sg, so)                                  # after a committed hop: permuted / retyped arrays identical
    vo, lo, _ = so.neighbor_list(); vg, lg, _ = sg.neighbor_list()
    assert np.array_equal(vo, vg) and np.array_equal(lo, lg)     # the commit rebuilds the list (ms_evb.f90:223-225)
    gi, gj, _ = sg.tile_pairs(); oi, oj, _ = so.tile_pairs()     # clusters of the 5/6-site acid and its conjugate base
    assert np.array_equal(gi, oi) and np.array_equal(gj, oj)
    so.md_integrate_atomic(12, ms_evb=True); sg.md_integrate_atomic(12, ms_evb=True)
    _compare_full_state(sg, so)
    assert sg.evb()["n_states"] == so.evb()["n_states"]


def test_state_sharding_emulated_acid_ion_pair(cuda_lib, oracle_lib):
    """three emulated ranks on the contact ion pair CH3SO3- + H3O+: the hop commit onto the sulfonate (re-ordered acceptor,
    reference force back-mapping quirk) with the diabats' pieces spread over the ranks"""
    s = system.build_acid_box(10, ion_pair=True)
    p = small_params()
    ref = engine.Simulation(s, p, library=oracle_lib); ref.ms_evb_calculate_total_force_energy()
    ranks = [engine.Simulation(s, p, library=cuda_lib, rank=r, world_size=3) for r in range(3)]
    _emulated_force(ranks)
    for sim in ranks:
        assert sim.evb()["principal_diabat"] == ref.evb()["principal_diabat"] == 2
        _compare_full_state(sim, ref)


def test_acid_trajectory_through_single_diabat_steps(cuda_lib, oracle_lib):
    """the acid box loses its hydrogen bond within ~25 steps: the diabat count drops to 1 (no acceptor in range), a
    degenerate MS-EVB step (nothing to couple, empty reciprocal-space algebra) that must still follow the oracle"""
    s = system.build_acid_box(10)
    p = small_params()
    so = engine.Simulation(s, p, library=oracle_lib)
    sg = engine.Simulation(s, p, library=cuda_lib)
    so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
    seen = set()
    for _ in range(6):
        so.md_integrate_atomic(6, ms_evb=True); sg.md_integrate_atomic(6, ms_evb=True)
        assert sg.evb()["n_states"] == so.evb()["n_states"]
        seen.add(so.evb()["n_states"])
        assert abs(sg.evb()["adiabatic_potential"] - so.evb()["adiabatic_potential"]) <= 1e-9 * max(abs(so.evb()["adiabatic_potential"]), abs(so.energies()["E_elec"]))
        assert np.abs(sg.download_state()["xyz"] - so.download_state()["xyz"]).max() < 1e-8
    assert 1 in seen, seen


def test_replica_ensemble_matches_individual_runs(cuda_lib):
    """rpb_ensemble_step (BASELINE config 5, replicas only): replicas driven concurrently by one host thread each end
    where the same replicas end when stepped one after the other."""
    n_steps = 10
    seeds = (1, 2, 3, 4)
    ens = [engine.Simulation(water_system(10, hydronium=True, seed=sd), small_params(), library=cuda_lib) for sd in seeds]
    one = [engine.Simulation(water_system(10, hydronium=True, seed=sd), small_params(), library=cuda_lib) for sd in seeds]
    for sim in ens + one:
        sim.ms_evb_calculate_total_force_energy()
    engine.Simulation.ensemble_step(ens, n_steps, ms_evb=True)
    for a, b in zip(ens, one):
        b.md_integrate_atomic(n_steps, ms_evb=True)
        sa, sb = a.download_state(), b.download_state()
        assert sa["hydronium_mol"] == sb["hydronium_mol"] and a.evb()["n_states"] == b.evb()["n_states"]
        assert np.abs(sa["xyz"] - sb["xyz"]).max() < 1e-9
        assert rel_rms(sa["force"], sb["force"]) < F_RTOL


def test_full_size_c2_properties(cuda_lib):
    """BASELINE config 2 (10 125 atoms): properties that need no oracle run -- Newton's third law for the real-space
    part, energy conservation, momentum removal, on-device rebuild reproducibility"""
    s = system.config_c2()
    sim = engine.Simulation(s, small_params(pme_grid=48), library=cuda_lib)
    sim.calculate_total_force_energy()
    f = sim.forces()
    assert np.isfinite(f).all() and np.abs(f.sum(axis=0)).max() < 1.0
    e0 = sim.energies()
    sim.md_integrate_atomic(100)
    e1 = sim.energies()
    assert abs((e1["potential_energy"] + e1["kinetic_energy"]) - (e0["potential_energy"] + e0["kinetic_energy"])) < 1e-2 * e0["kinetic_energy"]
    st = sim.download_state()
    assert np.abs((st["mass"][:, None] * st["velocity"]).sum(axis=0)).max() < 1e-6
    # a second context started from the downloaded state rebuilds the same neighbour list the first one carries
    vp, nl, flag = sim.neighbor_list()
    assert vp[-1] - 1 == len(nl) and (np.diff(vp) >= 0).all() and nl.min() >= 1 and nl.max() <= s.n_atoms


def test_full_size_c3_properties(cuda_lib):
    """BASELINE config 3 (9 001 atoms, ~20 diabats): Hamiltonian structure, normalisation, HF force = sum of weights"""
    s = system.config_c3()
    sim = engine.Simulation(s, small_params(pme_grid=48), library=cuda_lib)
    sim.ms_evb_calculate_total_force_energy()
    ev = sim.evb()
    S = ev["n_states"]
    assert 10 <= S <= 80
    assert abs((ev["eigenvector"] ** 2).sum() - 1.0) < 1e-12
    H = ev["hamiltonian"]; Hs = H + np.triu(H, 1).T
    w = np.linalg.eigvalsh(Hs)
    assert abs(w[0] - ev["adiabatic_potential"]) <= 1e-12 * abs(w[0])
    assert np.count_nonzero(np.triu(H, 1)) == S - 1                   # tree: one coupling per non-principal diabat
    # linearity of the Hellmann-Feynman mix in c_i c_j: F(c) for the ground state equals the stored adiabatic force
    assert rel_rms(sim.debug_mix_forces(ev["eigenvector"]), sim.forces()) < 1e-12


def test_recip_delta_algebra_every_diabat_force(cuda_lib, oracle_lib):
    """Reciprocal-space treatment of the diabats (charge-delta algebra on the principal grid: two convolutions per step)
    against the oracle, which keeps the reference's structure -- one patched grid and one FFT convolution per diabat
    (ms_evb.f90:1962-2248): Hamiltonian, and every diabat's own force (c = e_s), so that the chain atoms' reciprocal terms
    are exercised state by state."""
    s = water_system(10, hydronium=True)
    p = small_params()
    sa = engine.Simulation(s, p, library=cuda_lib)
    so = engine.Simulation(s, p, library=oracle_lib)
    for sim in (sa, so):
        sim.ms_evb_calculate_total_force_energy()
    ea, eo = sa.evb(), so.evb()
    assert ea["n_states"] == eo["n_states"] > 4
    scale = np.abs(np.diag(eo["hamiltonian"])).max()
    assert np.abs(ea["hamiltonian"] - eo["hamiltonian"]).max() <= E_RTOL * scale
    assert rel_rms(sa.forces(), so.forces()) < F_RTOL
    for k in range(ea["n_states"]):
        c = np.zeros(ea["n_states"]); c[k] = 1.0
        assert rel_rms(sa.debug_mix_forces(c), so.debug_mix_forces(c)) < F_RTOL, k
    for sim in (sa, so):
        sim.md_integrate_atomic(10, ms_evb=True)
    xa, xo = (sim.download_state() for sim in (sa, so))
    assert xa["hydronium_mol"] == xo["hydronium_mol"]
    assert np.abs(xa["xyz"] - xo["xyz"]).max() < 1e-9


def test_tree_solver_matches_block_jacobi(cuda_lib, oracle_lib, monkeypatch):
    """The default ground-state solver exploits the tree structure of the EVB Hamiltonian; the block-level Jacobi
    kernel (RPB_EVB_SOLVER=jacobi) and the oracle's Numerical-Recipes Jacobi (general_routines.f90:2013-2088) must
    give the same ground state, principal diabat and Hellmann-Feynman forces."""
    s = water_system(10, hydronium=True)
    p = small_params()
    monkeypatch.delenv("RPB_EVB_SOLVER", raising=False)
    st = engine.Simulation(s, p, library=cuda_lib)
    monkeypatch.setenv("RPB_EVB_SOLVER", "jacobi")
    sj = engine.Simulation(s, p, library=cuda_lib)
    monkeypatch.delenv("RPB_EVB_SOLVER", raising=False)
    so = engine.Simulation(s, p, library=oracle_lib)
    for sim in (st, sj, so):
        sim.ms_evb_calculate_total_force_energy()
    et, ej, eo = st.evb(), sj.evb(), so.evb()
    assert et["n_states"] == ej["n_states"] == eo["n_states"] > 1
    for a in (et, ej):
        assert abs(a["adiabatic_potential"] - eo["adiabatic_potential"]) <= 1e-12 * abs(eo["adiabatic_potential"])
        c = a["eigenvector"] * np.sign(np.dot(a["eigenvector"], eo["eigenvector"]))
        assert np.abs(c - eo["eigenvector"]).max() < 1e-10
        assert a["principal_diabat"] == eo["principal_diabat"]
    assert rel_rms(st.forces(), sj.forces()) < 1e-12
    # a few steps of dynamics: the warm-started solves stay on the oracle's trajectory
    for sim in (st, sj, so):
        sim.md_integrate_atomic(5, ms_evb=True)
    assert np.abs(st.download_state()["xyz"] - so.download_state()["xyz"]).max() < 1e-9
    assert np.abs(sj.download_state()["xyz"] - so.download_state()["xyz"]).max() < 1e-9
