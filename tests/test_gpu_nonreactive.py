"""CUDA path vs oracle, non-reactive force path (BASELINE config 2 at reduced size), through the C-ABI."""
import numpy as np
import pytest

from reactive_pb_nn_md_b200 import engine
from tests.util import F_RTOL, assert_energies_close, rel_rms, small_params, water_system

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(oracle_lib, cuda_lib):
    s = water_system(10)
    p = small_params()
    so = engine.Simulation(s, p, library=oracle_lib)
    sg = engine.Simulation(s, p, library=cuda_lib)
    so.calculate_total_force_energy()
    sg.calculate_total_force_energy()
    return sg, so


def test_backend_is_cuda(cuda_lib):
    assert cuda_lib.backend == "cuda-sm100a"


def test_neighbor_list_bit_exact(pair):
    sg, so = pair
    vpg, nlg, fg = sg.neighbor_list()
    vpo, nlo, fo = so.neighbor_list()
    assert np.array_equal(vpg, vpo)
    assert np.array_equal(nlg, nlo)      # same rows AND same order inside each row
    assert fg == fo
    gi, gj, _ = sg.tile_pairs(); oi, oj, _ = so.tile_pairs()      # the cluster-pair list the pair kernel consumes: same set
    assert np.array_equal(gi, oi) and np.array_equal(gj, oj)


def test_com_and_wrap(pair):
    sg, so = pair
    assert np.array_equal(sg.r_com(), so.r_com())
    assert np.array_equal(sg.download_state()["xyz"], so.download_state()["xyz"])


def test_energies(pair):
    sg, so = pair
    assert_energies_close(sg.energies(), so.energies())


def test_forces(pair):
    sg, so = pair
    assert rel_rms(sg.forces(), so.forces()) < F_RTOL


def test_pme_grids(pair):
    sg, so = pair
    Qg, thg, frg = sg.pme()
    Qo, tho, fro = so.pme()
    assert np.abs(Qg - Qo).max() < 1e-13 * max(1.0, np.abs(Qo).max())
    assert rel_rms(thg, tho) < 1e-12
    assert rel_rms(frg, fro) < F_RTOL


def test_short_trajectory(oracle_lib, cuda_lib):
    s = water_system(10, seed=7)
    p = small_params(pme_grid=32)
    so = engine.Simulation(s, p, library=oracle_lib)
    sg = engine.Simulation(s, p, library=cuda_lib)
    so.calculate_total_force_energy(); sg.calculate_total_force_energy()
    so.md_integrate_atomic(20); sg.md_integrate_atomic(20)
    a, b = sg.download_state(), so.download_state()
    assert np.abs(a["xyz"] - b["xyz"]).max() < 1e-9
    assert np.abs(a["velocity"] - b["velocity"]).max() < 1e-8
    assert_energies_close(sg.energies(), so.energies(), rtol=1e-9)
    assert abs(sg.energies()["kinetic_energy"] - so.energies()["kinetic_energy"]) < 1e-8 * so.energies()["kinetic_energy"]
    # neighbour lists still identical after the on-device rebuild schedule
    vg, ng, fg = sg.neighbor_list(); vo, no, fo = so.neighbor_list()
    assert np.array_equal(vg, vo) and np.array_equal(ng, no) and fg == fo
    gi, gj, _ = sg.tile_pairs(); oi, oj, _ = so.tile_pairs()
    assert np.array_equal(gi, oi) and np.array_equal(gj, oj)


def test_launch_counter(pair):
    sg, _ = pair
    own, fft = sg.launch_counts()
    assert own > 0 and fft == 0      # 32^3 grid: the hand-written FFT convolution runs, cuFFT is not called


@pytest.mark.parametrize("K", [36, 40, 48, 64])
def test_reciprocal_space_all_grid_sizes(oracle_lib, cuda_lib, K, monkeypatch):
    """The hand-written batched FFT convolution (grid sizes 2^a 3^b: 36, 48, 64), the cuFFT fallback (40 = 2^3 5) and
    the cuFFT cross-check of the hand-written path all give the oracle's theta grid, E_rec and reciprocal forces."""
    s = water_system(10)
    p = small_params(pme_grid=K)
    so = engine.Simulation(s, p, library=oracle_lib)
    so.calculate_total_force_energy()
    Qo, tho, fro = so.pme()
    sims = []
    monkeypatch.delenv("RPB_FFT", raising=False)
    sims.append(engine.Simulation(s, p, library=cuda_lib))
    monkeypatch.setenv("RPB_FFT", "cufft")
    sims.append(engine.Simulation(s, p, library=cuda_lib))
    monkeypatch.delenv("RPB_FFT", raising=False)
    for sg in sims:
        sg.calculate_total_force_energy()
        Qg, thg, frg = sg.pme()
        assert rel_rms(thg, tho) < 1e-12
        assert rel_rms(frg, fro) < F_RTOL
        assert abs(sg.energies()["E_recip"] - so.energies()["E_recip"]) <= 1e-11 * abs(so.energies()["E_recip"])
    assert rel_rms(sims[0].pme()[1], sims[1].pme()[1]) < 1e-13
