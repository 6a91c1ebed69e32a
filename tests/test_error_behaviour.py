"""The reference reports failures on the path with `stop "message"`; both libraries return the same negative status and
carry the reference's message text.  Same cases for the CPU oracle (run here) and the CUDA library
(`-m gpu`)."""
import numpy as np
import pytest

from reactive_pb_nn_md_b200 import engine
from reactive_pb_nn_md_b200._binding import RpbError
from tests.util import small_params, water_system


def _cases(lib):
    # spline_order /= 6: pme.f90:247 divides by 6.D0 whatever the order -- only 6 is self-consistent in the reference
    with pytest.raises(RpbError) as ei:
        engine.Simulation(water_system(10), small_params(spline_order=4), library=lib)
    assert ei.value.code == -7 and "spline_order" in str(ei.value)
    # non-orthorhombic box: main_ms_evb.f90:62-68
    s = water_system(10)
    s.box[0, 1] = 0.5
    with pytest.raises(RpbError) as ei:
        engine.Simulation(s, small_params(), library=lib)
    assert ei.value.code == -7 and "orthorhombic" in str(ei.value)
    # |F| > 1e5 on some atom: "force on atom ... is too big" md_integration.f90:515-527 (two oxygens 1.3 A apart: ~1e6)
    s = water_system(10)
    d = s.xyz[3:6] - s.xyz[3]
    s.xyz[3:6] = s.xyz[0] + np.array([1.3, 0.0, 0.0]) + d
    sim = engine.Simulation(s, small_params(), library=lib)
    sim.calculate_total_force_energy()
    with pytest.raises(RpbError) as ei:
        sim.md_integrate_atomic(1)
    assert ei.value.code == -4 and "too big" in str(ei.value)
    # more diabats than evb_max_states: ms_evb.f90:3107-3121
    s = water_system(10, hydronium=True)
    sim = engine.Simulation(s, small_params(), library=lib, evb_max_states=4)
    with pytest.raises(RpbError) as ei:
        sim.ms_evb_calculate_total_force_energy()
    assert ei.value.code == -6 and "evb_max_states" in str(ei.value)
    # neighbour list larger than the allocated array: general_routines.f90:1562
    with pytest.raises(RpbError) as ei:
        sim = engine.Simulation(water_system(10), small_params(), library=lib, verlet_capacity=100000)
        sim.calculate_total_force_energy()
    assert ei.value.code == -5 and "verlet neighbor list" in str(ei.value)


def test_error_behaviour_oracle(oracle_lib):
    _cases(oracle_lib)


@pytest.mark.gpu
def test_error_behaviour_cuda(cuda_lib):
    _cases(cuda_lib)
    # and the library stays usable afterwards
    sim = engine.Simulation(water_system(10), small_params(), library=cuda_lib)
    sim.calculate_total_force_energy()
    assert np.isfinite(sim.energies()["potential_energy"])
