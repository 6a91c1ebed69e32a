"""A synthetic variant of the example force field that switches on the two branches of the path the example
parameters never reach with non-zero numbers:

  * SAPT / Tang-Toennies pair terms (pairwise_real_space_sapt, src/pair_int_real_space.f90:651-690, parameters from the
    `custom_sapt_parameters` block of the .pmt, src/initialize_routines.f90:312-366): every pair with a hydrogen gets a
    Buckingham wall and damped C6/C8/C10/C12 dispersion with non-zero coefficients;
  * a Ryckaert-Bellemans dihedral (dihedral type 3, src/intra_bonded_interactions.f90:511-547): the C-S-O-H torsion
    of CH3SO3H.

Test input only; the numbers are chosen to be of ordinary magnitude, not to be physical."""
import os

from reactive_pb_nn_md_b200 import system
from reactive_pb_nn_md_b200.forcefield import load_forcefield

# name  A_exch A_elec A_ind A_dhf  exponent  C6 C8 C10 C12     (kJ/mol, 1/A, kJ/mol A^n)
_SAPT_ROWS = {
    "C_a":   (60000.0, 9000.0, 500.0, 300.0, 3.30, 1500.0, 9000.0, 50000.0, 300000.0),
    "S_a":   (90000.0, 12000.0, 700.0, 400.0, 3.10, 2500.0, 15000.0, 80000.0, 500000.0),
    "O_a":   (150000.0, 30000.0, 900.0, 500.0, 3.80, 1200.0, 6000.0, 30000.0, 160000.0),
    "O_ah":  (150000.0, 30000.0, 900.0, 500.0, 3.80, 1200.0, 6000.0, 30000.0, 160000.0),
    "H_a":   (4000.0, 900.0, 60.0, 30.0, 3.60, 90.0, 300.0, 1000.0, 3000.0),
    "C_b":   (60000.0, 9000.0, 500.0, 300.0, 3.30, 1500.0, 9000.0, 50000.0, 300000.0),
    "S_b":   (90000.0, 12000.0, 700.0, 400.0, 3.10, 2500.0, 15000.0, 80000.0, 500000.0),
    "O_b":   (150000.0, 30000.0, 900.0, 500.0, 3.80, 1200.0, 6000.0, 30000.0, 160000.0),
    "OW":    (160000.0, 32000.0, 950.0, 520.0, 3.75, 1300.0, 6500.0, 32000.0, 170000.0),
    "HW":    (3500.0, 800.0, 50.0, 25.0, 3.55, 80.0, 280.0, 900.0, 2800.0),
    "O_h3o": (140000.0, 28000.0, 850.0, 480.0, 3.85, 1100.0, 5500.0, 28000.0, 150000.0),
    "H_h3o": (3000.0, 700.0, 45.0, 22.0, 3.65, 70.0, 250.0, 800.0, 2500.0),
}


def sapt_rb_forcefield(molecule_type_order=("so3h", "so3", "h3o", "h2o")):
    pmt = open(os.path.join(system.DATA_DIR, "CH3SO3H.pmt")).read()
    top = open(os.path.join(system.DATA_DIR, "CH3SO3H_H2O.top")).read()
    head, tail = pmt.split("cross_terms", 1)
    names = [ln.split()[0] for ln in head.splitlines()[3:15]]
    block = "custom_sapt_parameters\nname A_exch A_elec A_ind A_dhf exponent C6 C8 C10 C12\n"
    for n in names:
        block += n + " " + " ".join("%r" % v for v in _SAPT_ROWS[n]) + "\n"
    pmt = head + block + "\ncross_terms" + tail
    old = "C_a  S_a   O_ah  H_a   1  180.0     2.92       3.0"
    assert old in top
    top = top.replace(old, "C_a  S_a   O_ah  H_a   3  9.28  12.16  -13.12  -3.06  26.24  -31.5")
    ff = load_forcefield(pmt, top, "opls", 3, molecule_type_order)
    return ff
