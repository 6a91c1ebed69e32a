"""CUDA path vs oracle on the BENCHMARKED configurations themselves (BASELINE configs[2], [3]: C3 = 9 001 atoms / 48^3,
C4 = 30 001 atoms / 64^3), on the synthetic force field with non-zero SAPT / Tang-Toennies rows and a Ryckaert-Bellemans
dihedral, and -- when the box has more than one GPU -- the state-sharded step over real NVLink peers.

Bars (north_star): neighbour list, diabat enumeration and hop selection bit-exact; Hamiltonian elements and energies 1e-10
relative; forces 1e-8 relative RMS.  The trajectories cross a committed proton hop (C3, step 15) and a displacement-
triggered neighbour-list rebuild (C3 step 35, C4 step 27), found with the oracle when the tests were written."""
import os

import numpy as np
import pytest

from reactive_pb_nn_md_b200 import engine, system
from tests.synthetic_ff import sapt_rb_forcefield
from tests.util import E_RTOL, F_RTOL, assert_pair_lists_identical, rel_rms, small_params

pytestmark = pytest.mark.gpu


def _threads():
    return max(1, min(os.cpu_count() or 1, 32))


def _compare_evaluation(sg, so, n_force_states=4):
    assert_pair_lists_identical(sg, so)                                             # incl. row order
    eg, eo = sg.evb(), so.evb()
    assert eg["n_states"] == eo["n_states"]
    assert np.array_equal(eg["proton_log"], eo["proton_log"])
    assert np.array_equal(eg["coupling_matrix"], eo["coupling_matrix"])
    assert eg["principal_diabat"] == eo["principal_diabat"] and eg["new_hydronium_mol"] == eo["new_hydronium_mol"]
    en_g, en_o = sg.energies(), so.energies()
    scale = max(np.abs(np.diag(eo["hamiltonian"])).max(), abs(en_o["E_elec"]))
    assert np.abs(eg["hamiltonian"] - eo["hamiltonian"]).max() <= E_RTOL * scale
    assert abs(eg["adiabatic_potential"] - eo["adiabatic_potential"]) <= E_RTOL * scale
    for k in ("E_elec", "E_vdw", "E_bond", "E_angle", "E_dihedral", "E_recip"):
        assert abs(en_g[k] - en_o[k]) <= E_RTOL * max(abs(en_o[k]), scale), k
    assert rel_rms(sg.forces(), so.forces()) < F_RTOL
    S = eo["n_states"]
    for k in sorted(set([0, 1, S // 2, S - 1]))[:n_force_states]:             # a diabat's own force: c = e_k
        c = np.zeros(S); c[k] = 1.0
        assert rel_rms(sg.debug_mix_forces(c), so.debug_mix_forces(c)) < F_RTOL, k
    return S


def _compare_state(sg, so, xtol=1e-9):
    a, b = sg.download_state(), so.download_state()
    for k in ("atom_type", "mol_first_atom", "mol_n_atom", "mol_type"):
        assert np.array_equal(a[k], b[k]), k
    assert a["hydronium_mol"] == b["hydronium_mol"]
    assert np.abs(a["charge"] - b["charge"]).max() == 0.0
    assert np.abs(a["xyz"] - b["xyz"]).max() < xtol
    assert rel_rms(a["force"], b["force"]) < F_RTOL
    assert sg.evb()["n_states"] == so.evb()["n_states"]
    assert np.array_equal(sg.evb()["proton_log"], so.evb()["proton_log"])
    assert_pair_lists_identical(sg, so)


def test_config_c3_against_oracle(cuda_lib, oracle_lib):
    """BASELINE configs[2] as benchmarked: one evaluation, then 40 steps across the committed hop of step 15 and the
    displacement-triggered rebuild of step 35."""
    s = system.config_c3()
    so = engine.Simulation(s, small_params(pme_grid=48, n_threads=_threads()), library=oracle_lib)
    sg = engine.Simulation(s, small_params(pme_grid=48), library=cuda_lib)
    so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
    S = _compare_evaluation(sg, so)
    assert S >= 15
    hyd0 = s.hydronium_mol
    hopped = rebuilt_by_flag = False
    for _ in range(4):
        lo_before = so.neighbor_list()[1]
        hyd_before = so.download_state()["hydronium_mol"]
        so.md_integrate_atomic(10, ms_evb=True); sg.md_integrate_atomic(10, ms_evb=True)
        _compare_state(sg, so)
        hyd_now = so.download_state()["hydronium_mol"]
        hopped |= hyd_now != hyd0
        lo_now = so.neighbor_list()[1]
        rebuilt_by_flag |= (hyd_now == hyd_before) and not np.array_equal(lo_now, lo_before)
    assert hopped and rebuilt_by_flag
    scale = abs(so.energies()["E_elec"])
    assert abs(sg.evb()["adiabatic_potential"] - so.evb()["adiabatic_potential"]) <= 1e-9 * scale


def test_config_c4_against_oracle(cuda_lib, oracle_lib):
    """BASELINE configs[3] in its parity form (one excess proton, 30 001 atoms, 64^3): one evaluation, then 30 steps across
    the displacement-triggered rebuild of step 27."""
    s = system.config_c4()
    so = engine.Simulation(s, small_params(pme_grid=64, n_threads=_threads()), library=oracle_lib)
    sg = engine.Simulation(s, small_params(pme_grid=64), library=cuda_lib)
    so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
    S = _compare_evaluation(sg, so, n_force_states=3)
    assert S >= 15
    lo0 = so.neighbor_list()[1]
    so.md_integrate_atomic(30, ms_evb=True); sg.md_integrate_atomic(30, ms_evb=True)
    _compare_state(sg, so)
    assert not np.array_equal(so.neighbor_list()[1], lo0)            # the list was rebuilt on the way


def test_config_c2_against_oracle(cuda_lib, oracle_lib):
    """BASELINE configs[1] (10 125 atoms, non-reactive): one evaluation and 10 steps."""
    s = system.config_c2()
    so = engine.Simulation(s, small_params(pme_grid=48, n_threads=_threads()), library=oracle_lib)
    sg = engine.Simulation(s, small_params(pme_grid=48), library=cuda_lib)
    n_pairs, n_words = assert_pair_lists_identical(sg, so)
    assert n_words < 0.3 * n_pairs                  # water: one list word serves ~4 listed pairs (two directions stored)
    so.calculate_total_force_energy(); sg.calculate_total_force_energy()
    en_g, en_o = sg.energies(), so.energies()
    for k in ("potential_energy", "E_elec", "E_vdw", "E_bond", "E_angle", "E_recip"):
        assert abs(en_g[k] - en_o[k]) <= E_RTOL * max(abs(en_o[k]), abs(en_o["E_elec"])), k
    assert rel_rms(sg.forces(), so.forces()) < F_RTOL
    so.md_integrate_atomic(10); sg.md_integrate_atomic(10)
    a, b = sg.download_state(), so.download_state()
    assert np.abs(a["xyz"] - b["xyz"]).max() < 1e-10 and rel_rms(a["force"], b["force"]) < F_RTOL


@pytest.mark.parametrize("evb", [False, True])
def test_sapt_tang_toennies_and_rb_dihedral(cuda_lib, oracle_lib, evb):
    """Non-zero SAPT rows (Buckingham + Tang-Toennies damped C6..C12 on every pair with a hydrogen) and a
    Ryckaert-Bellemans C-S-O-H torsion, on the reference's example molecule in water: Verlet pair kernel,
    intramolecular pairs, per-diabat real-space deltas and the bonded kernel all take the branches the example
    parameters leave at zero."""
    ff = sapt_rb_forcefield()
    nT = ff.n_atom_type
    assert (ff.vdw_type[:nT, :nT] == 1).sum() > 40 and ff.dihedral_type[0, 1, 3, 4] == 3
    assert np.abs(ff.vdw_parameter[8, 9]).min() > 0.0                     # OW-HW row: A, B, C6, C8, C10, C12 all non-zero
    s = system.build_acid_box(10, ff=ff)
    so = engine.Simulation(s, small_params(n_threads=_threads()), library=oracle_lib)
    sg = engine.Simulation(s, small_params(), library=cuda_lib)
    if evb:
        so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
        S = _compare_evaluation(sg, so, n_force_states=4)
        assert S >= 4
    else:
        so.calculate_total_force_energy(); sg.calculate_total_force_energy()
        en_g, en_o = sg.energies(), so.energies()
        for k in ("potential_energy", "E_elec", "E_vdw", "E_bond", "E_angle", "E_dihedral"):
            assert abs(en_g[k] - en_o[k]) <= E_RTOL * max(abs(en_o[k]), abs(en_o["E_elec"])), k
        assert abs(en_o["E_dihedral"]) > 1.0 and abs(en_o["E_vdw"]) > 1e3
        assert rel_rms(sg.forces(), so.forces()) < F_RTOL
    so.md_integrate_atomic(12, ms_evb=evb); sg.md_integrate_atomic(12, ms_evb=evb)
    a, b = sg.download_state(), so.download_state()
    assert a["hydronium_mol"] == b["hydronium_mol"]
    assert np.abs(a["xyz"] - b["xyz"]).max() < 1e-9 and rel_rms(a["force"], b["force"]) < F_RTOL


# ---- state sharding over real peers ---------------------------------------------------------------------------------
def _sharded_worker(rank, world, port, outdir, workload, n_steps):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    from reactive_pb_nn_md_b200 import engine as eng, system as sy
    from reactive_pb_nn_md_b200._binding import load_cuda
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    s, K = (sy.config_c3(), 48) if workload == "c3" else (sy.build_water_box(10, with_hydronium=True), 32)
    sim = eng.Simulation(s, small_params(pme_grid=K), library=load_cuda(), device=rank, rank=rank, world_size=world,
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


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,workload", [(2, "small"), (2, "c3"), (4, "c3"), (8, "c3")])
def test_state_sharding_on_real_peers(oracle_lib, world, workload):
    """One process per GPU, NCCL process group for the handle exchange, the two per-step exchanges through peer memory over
    NVLink inside rpb_step.  Skipped on boxes with fewer GPUs than ranks (ranks that spin on one another must never share
    a device)."""
    if _device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import tempfile
    import torch.multiprocessing as mp
    n_steps = 20
    s, K = (system.config_c3(), 48) if workload == "c3" else (system.build_water_box(10, with_hydronium=True), 32)
    ref = engine.Simulation(s, small_params(pme_grid=K, n_threads=_threads()), library=oracle_lib)
    ref.ms_evb_calculate_total_force_energy()
    er, evr, fr = ref.energies(), ref.evb(), ref.forces()
    ref.md_integrate_atomic(n_steps, ms_evb=True)
    xr = ref.download_state()
    with tempfile.TemporaryDirectory() as d:
        port = 29900 + (os.getpid() % 2000)
        mp.spawn(_sharded_worker, args=(world, port, d, workload, n_steps), nprocs=world, join=True)
        z = [np.load(os.path.join(d, "rank%d.npz" % r)) for r in range(world)]
    scale = max(np.abs(np.diag(evr["hamiltonian"])).max(), abs(er["E_elec"]))
    for q in z:
        assert int(q["S"]) == evr["n_states"] and np.array_equal(q["log"], evr["proton_log"])
        assert abs(float(q["ad"]) - evr["adiabatic_potential"]) <= E_RTOL * scale
        assert np.abs(q["H"] - evr["hamiltonian"]).max() <= E_RTOL * scale
        assert abs(float(q["pe"]) - er["potential_energy"]) <= E_RTOL * scale
        assert rel_rms(q["f0"], fr) < F_RTOL
        assert int(q["hyd"]) == xr["hydronium_mol"]
        assert np.abs(q["xyz"] - xr["xyz"]).max() < 1e-9
        assert rel_rms(q["force"], xr["force"]) < F_RTOL
        for k in ("xyz", "vel", "force"):
            assert np.array_equal(q[k], z[0][k]), k                    # replicated state: bit-identical on every rank


# ---- small boxes: the pair kernel's per-pair minimum-image instantiation ------------------------------------------------
def _small_box_params(**kw):
    p = small_params(pme_grid=32, na_nslist=18, nb_nslist=18, nc_nslist=18, **kw)
    p.real_space_cutoff = 9.0
    p.verlet_cutoff = 11.0
    return p


def test_small_box_uses_per_pair_minimum_image(cuda_lib, oracle_lib):
    """L/2 - r_cutoff = 3.4 A leaves no room for one minimum-image shift per (cluster, atom): the library launches the
    pair kernel that shifts every pair (k_pair_tiles<.., false>).  512 molecules, L = 24.9 A: one evaluation + 20 steps."""
    s = system.build_water_box(8, with_hydronium=True)
    assert 0.5 * s.box_length - 9.0 < 4.0
    so = engine.Simulation(s, _small_box_params(n_threads=_threads()), library=oracle_lib)
    sg = engine.Simulation(s, _small_box_params(), library=cuda_lib)
    so.ms_evb_calculate_total_force_energy(); sg.ms_evb_calculate_total_force_energy()
    _compare_evaluation(sg, so, n_force_states=2)
    so.md_integrate_atomic(20, ms_evb=True); sg.md_integrate_atomic(20, ms_evb=True)
    _compare_state(sg, so)
    # and the non-reactive evaluation of the same box
    so.calculate_total_force_energy(); sg.calculate_total_force_energy()
    assert rel_rms(sg.forces(), so.forces()) < F_RTOL


def test_oversized_cluster_fails_loudly(cuda_lib):
    """One shift per (cluster, atom) needs r_cutoff + (extent of three consecutive atoms of a molecule) < L/2; the choice is
    made from the box and verified on the device -- a molecule that violates it stops the step with an error instead of
    being computed with a wrong image."""
    from reactive_pb_nn_md_b200._binding import RpbError
    s = system.build_water_box(10, with_hydronium=False)
    assert 0.5 * s.box_length - 10.0 > 4.0
    sim = engine.Simulation(s, small_params(pme_grid=32), library=cuda_lib)
    sim.calculate_total_force_energy()
    xyz = s.xyz.copy()
    xyz[1] = xyz[0] + np.array([0.5 * s.box_length - 10.0 + 0.5, 0.0, 0.0])       # the first water's H, far from its O
    sim.upload_state(xyz, s.velocity)
    with pytest.raises(RpbError, match="farther apart than the box allows"):
        sim.calculate_total_force_energy()
