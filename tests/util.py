"""Shared helpers for the parity tests (oracle vs CUDA library through the same C-ABI)."""
import numpy as np

from reactive_pb_nn_md_b200 import engine, system

E_RTOL = 1e-10     # north_star: per-state and ground-state energies, relative
F_RTOL = 1e-8      # north_star: forces, relative RMS


def rel_rms(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def small_params(**kw):
    p = dict(pme_grid=32, na_nslist=10, nb_nslist=10, nc_nslist=10, n_threads=1)
    p.update(kw)
    return engine.SimulationParameters(**p)


def water_system(n_side=10, hydronium=False, n_molecules=None, seed=20171017):
    return system.build_water_box(n_side, with_hydronium=hydronium, n_molecules=n_molecules, seed=seed)


def assert_energies_close(eg, eo, keys=("potential_energy", "E_elec", "E_vdw", "E_bond", "E_angle", "E_dihedral", "E_recip"),
                          rtol=E_RTOL):
    scale = max(abs(eo["potential_energy"]), abs(eo["E_elec"]), 1.0)
    for k in keys:
        assert abs(eg[k] - eo[k]) <= rtol * max(abs(eo[k]), scale), (k, eg[k], eo[k])


def assert_pair_lists_identical(sg, so):
    """neighbour list: the reference-ordered half list (accessor) is bit-identical incl. row order, AND the cluster-pair
    list the CUDA pair kernel actually consumes encodes exactly the same set of pairs"""
    vo, lo, fo = so.neighbor_list(); vg, lg, fg = sg.neighbor_list()
    assert np.array_equal(vo, vg) and np.array_equal(lo, lg) and fo == fg
    gi, gj, n_words = sg.tile_pairs()
    oi, oj, _ = so.tile_pairs()
    assert len(gi) == len(lo) and np.array_equal(gi, oi) and np.array_equal(gj, oj)
    return len(lo), n_words
