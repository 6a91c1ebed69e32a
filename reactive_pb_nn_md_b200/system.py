"""Synthetic inputs for the configurations BASELINE.json names, plus the flat
host-side description of a system (what the reference keeps in atom_data / molecule_data /
system_data, src/glob_v.f90:125-176).  Input generation only -- no force-path arithmetic.
"""
import os

import numpy as np

from . import tables
from .forcefield import load_forcefield

_HERE = os.path.dirname(os.path.abspath(__file__))
DATA_DIR = os.path.join(_HERE, "data")


def example_forcefield(n_exclusions=3, molecule_type_order=("h3o", "h2o")):
    """The reference's example force field (example_input_files/CH3SO3H.pmt, CH3SO3H_H2O.top),
    carried as text fixtures under data/ (parameters only)."""
    pmt = open(os.path.join(DATA_DIR, "CH3SO3H.pmt")).read()
    top = open(os.path.join(DATA_DIR, "CH3SO3H_H2O.top")).read()
    return load_forcefield(pmt, top, "opls", n_exclusions, molecule_type_order)


class System:
    """Flat arrays in the reference's layout: xyz/velocity (3,N) column-major == C [N][3]."""

    def __init__(self, ff, box_length, mol_names, xyz, velocity=None):
        self.ff = ff
        self.box_length = float(box_length)
        self.box = np.zeros((3, 3), order="F")
        for i in range(3):
            self.box[i, i] = self.box_length
        self.mol_type = np.array([ff.mtype(n) for n in mol_names], np.int32)
        self.mol_n_atom = np.array([ff.molecule_types[t - 1].n_atom for t in self.mol_type], np.int32)
        self.mol_first_atom = (np.concatenate(([0], np.cumsum(self.mol_n_atom)[:-1])) + 1).astype(np.int32)
        self.n_mole = len(mol_names)
        self.n_atoms = int(self.mol_n_atom.sum())
        self.atom_type = np.concatenate([ff.molecule_types[t - 1].atom_types for t in self.mol_type]).astype(np.int32)
        self.charge = ff.atype_chg[self.atom_type - 1].copy()            # gen_param :549-550
        self.mass = ff.atype_mass[self.atom_type - 1].copy()             # fill_mass
        self.xyz = np.ascontiguousarray(xyz, np.float64).reshape(self.n_atoms, 3)
        self.velocity = np.zeros_like(self.xyz) if velocity is None else np.ascontiguousarray(velocity, np.float64)
        # update_hydronium_molecule_index ms_evb.f90:98-135
        acid = [i + 1 for i, t in enumerate(self.mol_type) if ff.has_evb and ff.evb_proton_index[t - 1] > 0]
        if len(acid) > 1:
            raise ValueError("can't have more than 1 hydronium, see code comments")
        self.hydronium_mol = acid[0] if acid else 0


def _water_geometry(rng, a, b, r_oh=1.012, theta_deg=113.24):
    """two O-H vectors in the plane of unit vectors a,b (perpendicular), HOH angle theta,
    bisector along (a+b)."""
    half = np.radians(theta_deg) / 2.0
    bis = (a + b) / np.sqrt(2.0)
    perp = (a - b) / np.sqrt(2.0)
    return r_oh * (np.cos(half) * bis + np.sin(half) * perp), r_oh * (np.cos(half) * bis - np.sin(half) * perp)


def _h3o_geometry(axes, r_oh=1.0, s=0.1879):
    n = axes[0] + axes[1] + axes[2]
    out = []
    for e in axes:
        v = e - s * n
        out.append(r_oh * v / np.linalg.norm(v))
    return out


def build_water_box(n_side, with_hydronium=False, n_molecules=None, seed=20171017, density=0.03334,
                    temperature=300.0, jitter=0.05, ff=None):
    """Waters (and optionally one H3O+) on a jittered simple-cubic lattice with every O-H bond
    pointing at a lattice neighbour (so that proton-hop chains exist), positions rounded to
    .gro precision (1e-3 nm), Maxwell-Boltzmann velocities with COM motion removed."""
    rng = np.random.default_rng(seed)
    if ff is None:
        ff = example_forcefield(molecule_type_order=("h3o", "h2o") if with_hydronium else ("h2o",))
    n_sites = n_side ** 3
    if n_molecules is None:
        n_molecules = n_sites
    L = (n_molecules / density) ** (1.0 / 3.0)
    L = round(L, 2)
    a0 = L / n_side
    sites = np.array([(i, j, k) for i in range(n_side) for j in range(n_side) for k in range(n_side)], float)
    centre = np.array([n_side // 2] * 3, float)
    if n_molecules < n_sites:
        # keep the sites closest to the centre occupied, drop a random subset of the far ones
        d = np.abs(sites - centre).sum(axis=1)
        far = np.where(d > 4)[0]
        drop = rng.choice(far, n_sites - n_molecules, replace=False)
        keep = np.ones(n_sites, bool); keep[drop] = False
        sites = sites[keep]
    eye = np.eye(3)
    names, coords = [], []
    order = np.arange(len(sites))
    if with_hydronium:
        ic = int(np.where((sites == centre).all(axis=1))[0][0])
        order = np.concatenate(([ic], np.delete(order, ic)))
    for n, idx in enumerate(order):
        o = (sites[idx] + 0.5) * a0 + rng.normal(0.0, jitter, 3)
        if with_hydronium and n == 0:
            axes = [eye[k] * rng.choice([-1.0, 1.0]) for k in range(3)]
            names.append("h3o")
            coords.append(o)
            coords += [o + h for h in _h3o_geometry(axes)]
        else:
            ax = rng.choice(3, 2, replace=False)
            a = eye[ax[0]] * rng.choice([-1.0, 1.0]); b = eye[ax[1]] * rng.choice([-1.0, 1.0])
            h1, h2 = _water_geometry(rng, a, b)
            names.append("h2o")
            coords += [o, o + h1, o + h2]
    xyz = np.round(np.array(coords), 2)          # .gro F8.3 in nm == 0.01 A
    sysm = System(ff, L, names, xyz)
    sigma = np.sqrt(tables.BOLTZMANN * temperature / sysm.mass * tables.CONV_KJMOL)
    v = rng.normal(size=(sysm.n_atoms, 3)) * sigma[:, None]
    p = (sysm.mass[:, None] * v).sum(axis=0)
    v -= p / sysm.mass.sum()
    sysm.velocity = np.ascontiguousarray(v)
    return sysm


def build_acid_box(n_side=10, seed=20171017, density=0.03334, temperature=300.0, jitter=0.05, ff=None, ion_pair=False):
    """BASELINE config 1 -- the reference's own example system: one CH3SO3H (united-atom methyl, 6 sites) in water, the
    acid at the centre of the lattice with its O-H pointing at the +x lattice neighbour (so that proton-transfer
    diabats so3h + h2o -> so3 + h3o exist), the lattice sites within 3.3 A of its heavy atoms left out to make room.  Molecule types
    in the order of the example topology: so3h, so3, h3o, h2o.  ion_pair: the same geometry written as the contact ion pair
    CH3SO3- + H3O+ (molecules so3, h3o, waters; the H3O+ donates to the sulfonate oxygen): the acid diabat is the ground
    state, so the first evaluation commits the hop h3o + so3 -> h2o + so3h and re-orders the acceptor to its template."""
    rng = np.random.default_rng(seed)
    if ff is None:
        ff = example_forcefield(molecule_type_order=("so3h", "so3", "h3o", "h2o"))
    n_sites = n_side ** 3
    L = round((n_sites / density) ** (1.0 / 3.0), 2)
    a0 = L / n_side
    sites = np.array([(i, j, k) for i in range(n_side) for j in range(n_side) for k in range(n_side)], float)
    centre = np.array([n_side // 2] * 3, float)
    # acid geometry: tetrahedral S, O_ah along +x, C / O_a / O_a on the cone at 109.47 degrees
    eye = np.eye(3)
    sS = (centre + 0.5) * a0 + np.array([-1.2, 0.0, 0.0])
    cone = lambda phi: np.array([-1.0 / 3.0, 2.0 * np.sqrt(2.0) / 3.0 * np.cos(phi), 2.0 * np.sqrt(2.0) / 3.0 * np.sin(phi)])
    C = sS + 1.77 * cone(np.radians(90.0))
    Oa1 = sS + 1.45 * cone(np.radians(210.0))
    Oa2 = sS + 1.45 * cone(np.radians(330.0))
    Oah = sS + 1.60 * eye[0]
    Ha = Oah + 0.97 * np.array([np.cos(np.radians(40.0)), 0.0, np.sin(np.radians(40.0))])   # S-O-H = 140 degrees: dihedral C-S-O-H defined
    acceptor = centre + np.array([1.0, 0.0, 0.0])
    acid = np.array([C, sS, Oa1, Oa2, Oah, Ha])
    pos = (sites + 0.5) * a0
    dmin = np.sqrt(((pos[:, None, :] - acid[None, :5, :]) ** 2).sum(axis=2)).min(axis=1)
    keep = (dmin > 3.3) | (sites == centre).all(axis=1) | (sites == acceptor).all(axis=1)
    sites = sites[keep]
    ic = int(np.where((sites == centre).all(axis=1))[0][0])
    order = np.concatenate(([ic], np.delete(np.arange(len(sites)), ic)))
    eye = np.eye(3)
    names, coords = [], []
    for n, idx in enumerate(order):
        o = (sites[idx] + 0.5) * a0 + rng.normal(0.0, jitter, 3)
        if n == 0 and ion_pair:
            names.append("so3")
            coords += [C, sS, Oah, Oa1, Oa2]                    # the oxygen that will be protonated is NOT last: exercises the re-ordering
        elif n == 0:
            names.append("so3h")
            coords += list(acid)                                # atom order of [ moleculetype ] so3h
        elif (sites[idx] == acceptor).all() and ion_pair:
            o = (sites[idx] + 0.5) * a0
            names.append("h3o")
            coords.append(o)
            coords += [o + h for h in _h3o_geometry([-eye[0], eye[1], eye[2]])]
        elif (sites[idx] == acceptor).all():
            # the acceptor: lone pair towards the acid, both O-H pointing away from it
            h1, h2 = _water_geometry(rng, eye[0], eye[1])
            names.append("h2o")
            coords += [o, o + h1, o + h2]
        else:
            ax = rng.choice(3, 2, replace=False)
            a = eye[ax[0]] * rng.choice([-1.0, 1.0]); b = eye[ax[1]] * rng.choice([-1.0, 1.0])
            h1, h2 = _water_geometry(rng, a, b)
            names.append("h2o")
            coords += [o, o + h1, o + h2]
    xyz = np.round(np.array(coords), 2)
    sysm = System(ff, L, names, xyz)
    sigma = np.sqrt(tables.BOLTZMANN * temperature / sysm.mass * tables.CONV_KJMOL)
    v = rng.normal(size=(sysm.n_atoms, 3)) * sigma[:, None]
    p = (sysm.mass[:, None] * v).sum(axis=0)
    v -= p / sysm.mass.sum()
    sysm.velocity = np.ascontiguousarray(v)
    return sysm


def config_c1(**kw):
    """BASELINE config 1: CH3SO3H + 993 H2O (the reference's example input), MS-EVB, 2985 atoms, L~31.1 A."""
    return build_acid_box(10, **kw)


def config_c2(n_side=15, **kw):
    """BASELINE config 2: non-reactive flexible water box, 3375 H2O = 10125 atoms, L~46.6 A."""
    return build_water_box(n_side, with_hydronium=False, **kw)


def config_c3(**kw):
    """BASELINE config 3: H3O+ + 2999 H2O, MS-EVB, 9001 atoms, L~44.8 A."""
    return build_water_box(15, with_hydronium=True, n_molecules=3000, **kw)


def config_c4(**kw):
    """BASELINE config 4 (parity form): one H3O+ + 9999 H2O, 30001 atoms, L~66.9 A, K=64."""
    return build_water_box(22, with_hydronium=True, n_molecules=10000, **kw)
