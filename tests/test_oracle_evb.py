"""MS-EVB bookkeeping of the oracle on hand-worked clusters (known-answer enumeration cases)."""
import numpy as np
import pytest

from reactive_pb_nn_md_b200 import engine, system
from tests.util import rel_rms, small_params, water_system

L = 31.07
RNG = np.random.default_rng(5)


def water_at(o, h_dirs, r=1.0):
    return [o] + [o + r * np.asarray(d, float) / np.linalg.norm(d) for d in h_dirs]


def filler_waters(n, exclude_centre, min_dist=8.0):
    """far-away waters so that the cell grid / PME have something to chew on"""
    out = []
    while len(out) < n:
        o = RNG.uniform(0, L, 3)
        if np.linalg.norm((o - exclude_centre + L / 2) % L - L / 2) < min_dist:
            continue
        if any(np.linalg.norm((o - q[0] + L / 2) % L - L / 2) < 2.6 for q in out):
            continue
        out.append(water_at(o, [(1, 0.2, 0), (-0.3, 1, 0)]))
    return out


def build(ff, mols, names):
    xyz = np.array([a for m in mols for a in m])
    return system.System(ff, L, names, xyz)


@pytest.fixture(scope="module")
def ff():
    return system.example_forcefield()


def test_eigen_cation_four_states(oracle_lib, ff):
    """H9O4+: hydronium with three H-bonded waters and nothing else reactive -> S = 4, all children of diabat 1"""
    c = np.array([15.0, 15.0, 15.0])
    dirs = [np.array([1, 0, 0.0]), np.array([-0.5, 0.85, 0.0]), np.array([-0.5, -0.85, 0.0])]
    h3o = water_at(c, dirs)
    shell = [water_at(c + 2.55 * d / np.linalg.norm(d), [d + np.array([0, 0, 0.9]), d - np.array([0, 0, 0.9])]) for d in dirs]
    fill = filler_waters(60, c)
    s = build(ff, [h3o] + shell + fill, ["h3o"] + ["h2o"] * (3 + len(fill)))
    sim = engine.Simulation(s, small_params(), library=oracle_lib)
    sim.ms_evb_calculate_total_force_energy()
    ev = sim.evb()
    assert ev["n_states"] == 4
    log = ev["proton_log"]
    assert (log[0] == -1).all()
    for k in range(3):
        # (donor mol, proton index in donor, heavy atom it is bonded to, acceptor mol, acceptor atom), 1-based
        assert list(log[k + 1, 0]) == [1, k + 2, 1, k + 2, 1]
        assert (log[k + 1, 1:] == -1).all()
    assert list(ev["coupling_matrix"]) == [-1, 1, 1, 1]
    H = ev["hamiltonian"]
    assert (np.diag(H)[1:] > H[0, 0] - 400).all()
    # tree-structured: only (parent, child) couplings
    off = np.triu(H, 1)
    assert (off[0, 1:] != 0).all() and np.count_nonzero(off) == 3
    w, v = np.linalg.eigh(H + np.triu(H, 1).T)
    assert abs(w[0] - ev["adiabatic_potential"]) < 1e-9 * abs(w[0])
    c0 = ev["eigenvector"] * np.sign(ev["eigenvector"][0]) * np.sign(v[0, 0])
    assert np.abs(c0 - v[:, 0]).max() < 1e-9


def test_chain_is_cut_at_evb_max_chain(oracle_lib, ff):
    """a water wire h3o -> w1 -> w2 -> w3 -> w4 : diabats stop after 3 hops (evb_max_chain, glob_v.f90:65)"""
    base = np.array([8.0, 15.0, 15.0])
    step = np.array([2.55, 0, 0])
    up = np.array([0, 0.8, 0.5])
    h3o = water_at(base, [(1, 0, 0), (-0.4, 0.9, 0.2), (-0.4, -0.9, 0.2)])
    wire = [water_at(base + (k + 1) * step, [(1, 0.15, 0.1), (0.1, 0.6 * (-1) ** k, 0.8)]) for k in range(4)]
    fill = filler_waters(60, base + 2 * step, min_dist=11.0)
    s = build(ff, [h3o] + wire + fill, ["h3o"] + ["h2o"] * (4 + len(fill)))
    sim = engine.Simulation(s, small_params(), library=oracle_lib)
    sim.ms_evb_calculate_total_force_energy()
    ev = sim.evb()
    assert ev["n_states"] == 4
    assert [int((ev["proton_log"][k, :, 0] > 0).sum()) for k in range(4)] == [0, 1, 2, 3]
    assert list(ev["coupling_matrix"]) == [-1, 1, 2, 3]
    assert [int(ev["proton_log"][k, k - 1, 3]) for k in (1, 2, 3)] == [2, 3, 4]


def test_hop_commit_permutes_topology(oracle_lib, ff):
    """proton sitting closer to the acceptor oxygen: the acceptor diabat dominates and the hop is committed
    (evb_change_diabat_data_structure_topology, ms_evb.f90:806-932, 2677-2840)"""
    c = np.array([15.0, 15.0, 15.0])
    w_o = c + np.array([2.42, 0, 0])
    h3o = [c, c + np.array([1.32, 0, 0]), c + np.array([-0.35, 0.93, 0]), c + np.array([-0.35, -0.93, 0])]
    wat = water_at(w_o, [(0.4, 0.8, 0.4), (0.4, -0.8, 0.4)])
    fill = filler_waters(60, c)
    for order in ("h3o_first", "h3o_last"):
        if order == "h3o_first":
            mols, names, hyd = [h3o, wat] + fill, ["h3o", "h2o"] + ["h2o"] * len(fill), 1
        else:
            mols, names, hyd = fill[:5] + [wat] + fill[5:] + [h3o], ["h2o"] * (len(fill) + 1) + ["h3o"], len(fill) + 2
        s = build(ff, mols, names)
        assert s.hydronium_mol == hyd
        sim = engine.Simulation(s, small_params(), library=oracle_lib)
        before = sim.download_state()
        sim.ms_evb_calculate_total_force_energy()
        ev = sim.evb()
        assert ev["principal_diabat"] == 2 and ev["new_hydronium_mol"] != hyd
        after = sim.download_state()
        assert after["hydronium_mol"] == ev["new_hydronium_mol"]
        new_h, old_h = after["hydronium_mol"] - 1, hyd - 1
        assert after["mol_n_atom"][new_h] == 4 and after["mol_n_atom"][old_h] == 3
        assert after["mol_type"][new_h] == ff.mtype("h3o") and after["mol_type"][old_h] == ff.mtype("h2o")
        # contiguous ascending molecule ranges are preserved
        assert np.array_equal(after["mol_first_atom"], np.concatenate(([1], 1 + np.cumsum(after["mol_n_atom"])[:-1])))
        f = after["mol_first_atom"][new_h] - 1
        assert list(after["atom_type"][f:f + 4]) == [ff.atype("O_h3o")] + [ff.atype("H_h3o")] * 3
        f2 = after["mol_first_atom"][old_h] - 1
        assert list(after["atom_type"][f2:f2 + 3]) == [ff.atype("OW"), ff.atype("HW"), ff.atype("HW")]
        assert np.allclose(after["charge"][f:f + 4], [-0.5, 0.5, 0.5, 0.5]) and abs(after["charge"].sum() - 1.0) < 1e-12
        # the multiset of positions is unchanged (the proton record moved, nothing was lost)
        key = lambda a: np.array(sorted(map(tuple, np.round(a, 9))))
        assert np.array_equal(key(before["xyz"]), key(after["xyz"]))
        # the transferred proton is the last atom of the new hydronium
        assert np.allclose(after["xyz"][f + 3], h3o[1])


def test_sharded_phases_equal_single_rank(oracle_lib):
    """world_size=2 emulated in one process: ranks build disjoint diabats, exchange buffers are summed by hand"""
    import ctypes as C
    s = water_system(10, hydronium=True)
    p = small_params()
    ref = engine.Simulation(s, p, library=oracle_lib)
    ref.ms_evb_calculate_total_force_energy()
    ranks = [engine.Simulation(s, p, library=oracle_lib, rank=r, world_size=2) for r in range(2)]

    def buf(sim, which):
        ptr, n = C.c_void_p(), C.c_int()
        fn = sim.dll.rpb_evb_exchange_h if which == "h" else sim.dll.rpb_evb_exchange_f
        sim._check(fn(sim.ctx, C.byref(ptr), C.byref(n)))
        return np.frombuffer((C.c_double * n.value).from_address(ptr.value), dtype=np.float64)

    for which, phase in (("h", "rpb_evb_phase_build"), ("f", "rpb_evb_phase_mix")):
        for sim in ranks:
            sim._check(getattr(sim.dll, phase)(sim.ctx))
        bufs = [buf(sim, which) for sim in ranks]
        total = bufs[0] + bufs[1]
        for b in bufs:
            b[:] = total
    for sim in ranks:
        sim._check(sim.dll.rpb_evb_phase_commit(sim.ctx))
        assert sim.evb()["n_states"] == ref.evb()["n_states"]
        assert abs(sim.energies()["potential_energy"] - ref.energies()["potential_energy"]) < 1e-9
        assert rel_rms(sim.forces(), ref.forces()) < 1e-12
