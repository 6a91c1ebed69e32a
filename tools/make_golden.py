#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (single thread, source-order arithmetic).
The reference itself cannot be run in this image (no Fortran compiler / MKL), so these vectors pin the ORACLE,
not the reference: parity stays "unpinned" in the sense of SURVEY 8c.   usage: python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reactive_pb_nn_md_b200 import engine  # noqa: E402
from reactive_pb_nn_md_b200._binding import Library  # noqa: E402
from tests.util import small_params, water_system  # noqa: E402

lib = Library(os.path.join(ROOT, "oracle", "librpbmd_oracle.so"))
out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)

s = water_system(10)
sim = engine.Simulation(s, small_params(), library=lib)
sim.calculate_total_force_energy()
e = sim.energies(); f = sim.forces(); vp, nl, _ = sim.neighbor_list(); Q, th, fr = sim.pme()
np.savez_compressed(os.path.join(out, "water1000_nonreactive.npz"),
                    energies=np.array([e[k] for k in ("potential_energy", "E_elec", "E_vdw", "E_bond", "E_angle", "E_dihedral", "E_recip")]),
                    force_head=f[:96], force_sq_sum=(f ** 2).sum(), n_pairs=len(nl), verlet_point=vp,
                    nl_checksum=np.array([int(nl.astype(np.int64).sum()), int((nl.astype(np.int64) * (np.arange(len(nl)) % 1009 + 1)).sum())]),
                    Q_sum=Q.sum(), Q_abs_sum=np.abs(Q).sum(), theta_abs_sum=np.abs(th).sum(), force_recip_head=fr[:96])

s = water_system(10, hydronium=True)
sim = engine.Simulation(s, small_params(), library=lib)
sim.ms_evb_calculate_total_force_energy()
ev = sim.evb(); f = sim.forces()
np.savez_compressed(os.path.join(out, "h3o_water999_msevb.npz"), n_states=ev["n_states"], proton_log=ev["proton_log"],
                    coupling_matrix=ev["coupling_matrix"], hamiltonian=ev["hamiltonian"], eigenvector=ev["eigenvector"],
                    adiabatic_potential=ev["adiabatic_potential"], principal_diabat=ev["principal_diabat"],
                    force_head=f[:96], force_sq_sum=(f ** 2).sum())
# BASELINE config 1 as the contact ion pair CH3SO3- + H3O+: the first evaluation commits the hop onto the sulfonate
from reactive_pb_nn_md_b200 import system  # noqa: E402
s = system.build_acid_box(10, ion_pair=True)
sim = engine.Simulation(s, small_params(), library=lib)
sim.ms_evb_calculate_total_force_energy()
ev = sim.evb(); st = sim.download_state(); e = sim.energies()
np.savez_compressed(os.path.join(out, "acid_ion_pair_msevb.npz"), n_states=ev["n_states"], proton_log=ev["proton_log"],
                    hamiltonian=ev["hamiltonian"], eigenvector=ev["eigenvector"], adiabatic_potential=ev["adiabatic_potential"],
                    principal_diabat=ev["principal_diabat"], new_hydronium_mol=ev["new_hydronium_mol"],
                    energies=np.array([e[k] for k in ("E_elec", "E_vdw", "E_bond", "E_angle", "E_dihedral", "E_recip")]),
                    force_head=st["force"][:24], force_sq_sum=(st["force"] ** 2).sum(), atom_type_head=st["atom_type"][:12],
                    mol_type_head=st["mol_type"][:4], mol_n_atom_head=st["mol_n_atom"][:4], xyz_head=st["xyz"][:12])
print("golden vectors written to", out)
