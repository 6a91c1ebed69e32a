"""Diabatic-state sharding over torch.distributed (gloo, world_size 2, CPU): the same engine code that drives the CUDA
library over NCCL, here in front of the oracle.  Two real processes, rendezvous on 127.0.0.1."""
import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, outdir, acid=False):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from reactive_pb_nn_md_b200 import engine
    from reactive_pb_nn_md_b200._binding import Library
    from tests.util import small_params, water_system
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = Library(os.path.join(ROOT, "oracle", "librpbmd_oracle.so"))
    if acid:
        from reactive_pb_nn_md_b200 import system
        s = system.build_acid_box(10, ion_pair=True)
    else:
        s = water_system(10, hydronium=True)
    sim = engine.Simulation(s, small_params(), library=lib, rank=rank, world_size=world, process_group=dist.group.WORLD)
    sim.ms_evb_calculate_total_force_energy()
    sim.md_integrate_atomic(3, ms_evb=True)
    st = sim.download_state()
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), xyz=st["xyz"], vel=st["velocity"], force=st["force"], atype=st["atom_type"],
             pe=sim.energies()["potential_energy"], S=sim.evb()["n_states"], hyd=st["hydronium_mol"])
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank(oracle_lib):
    import torch.multiprocessing as mp
    from reactive_pb_nn_md_b200 import engine
    from tests.util import small_params, water_system
    s = water_system(10, hydronium=True)
    ref = engine.Simulation(s, small_params(), library=oracle_lib)
    ref.ms_evb_calculate_total_force_energy()
    ref.md_integrate_atomic(3, ms_evb=True)
    r = ref.download_state()
    with tempfile.TemporaryDirectory() as d:
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_worker, args=(2, port, d), nprocs=2, join=True)
        for rank in range(2):
            z = np.load(os.path.join(d, "rank%d.npz" % rank))
            assert int(z["S"]) == ref.evb()["n_states"] and int(z["hyd"]) == r["hydronium_mol"]
            assert np.abs(z["xyz"] - r["xyz"]).max() < 1e-11
            assert np.abs(z["vel"] - r["velocity"]).max() < 1e-10
            assert np.abs(z["force"] - r["force"]).max() < 1e-8
            assert abs(float(z["pe"]) - ref.energies()["potential_energy"]) < 1e-9


def test_two_rank_gloo_acid_ion_pair_commit(oracle_lib):
    """the hop commit onto the sulfonate (BASELINE config 1 as the contact ion pair) with the diabats split over two
    ranks: the permuted / retyped arrays and the trajectory equal the single-rank run"""
    import torch.multiprocessing as mp
    from reactive_pb_nn_md_b200 import engine, system
    from tests.util import small_params
    s = system.build_acid_box(10, ion_pair=True)
    ref = engine.Simulation(s, small_params(), library=oracle_lib)
    ref.ms_evb_calculate_total_force_energy()
    ref.md_integrate_atomic(3, ms_evb=True)
    r = ref.download_state()
    assert r["hydronium_mol"] == 1 and r["mol_type"][0] == s.ff.mtype("so3h")
    with tempfile.TemporaryDirectory() as d:
        port = 31500 + (os.getpid() % 2000)
        mp.spawn(_worker, args=(2, port, d, True), nprocs=2, join=True)
        for rank in range(2):
            z = np.load(os.path.join(d, "rank%d.npz" % rank))
            assert int(z["hyd"]) == r["hydronium_mol"] and np.array_equal(z["atype"], r["atom_type"])
            assert np.abs(z["xyz"] - r["xyz"]).max() < 1e-11
            assert np.abs(z["force"] - r["force"]).max() < 1e-8
