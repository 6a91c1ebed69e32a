"""Host-side force-field setup: the parts of the reference's initialisation whose RESULTS the
force path consumes as flat arrays (SURVEY.md 8(f) N2 -- readers themselves are not the hot path).

Restates, for the example force field format:
  read_param / gen_param / read_generate_14_interaction_parameters
        src/initialize_routines.f90:281-430, 448-634, 646-691
  read_topology_file (bondtypes / angletypes / dihedraltypes / moleculetype blocks)
        src/intra_bonded_interactions.f90:696-1464
  generate_intramolecular_exclusions  src/intra_bonded_interactions.f90:574-679
  read_evb_parameters / read_evb_topology  src/ms_evb.f90:3170-3623

All arrays use the Fortran shapes of src/glob_v.f90 (column-major, 1-based index VALUES) so
they can be handed to the C-ABI unchanged.  Deviations from the shipped (stale) example files,
which the reference's own readers reject, are tolerated here and listed in DESIGN.md:
6-column solute_species rows, missing [ exclusions ] blocks, one-line [ geometry_factor ] rows.
"""
import numpy as np

from ._binding import (MAX_INTERACTION_TYPE, MAX_MOLE_ATOMS, MAX_N_ATOM_TYPE, MAX_N_MOLE_TYPE)

T = MAX_N_ATOM_TYPE
MT = MAX_N_MOLE_TYPE
MI = MAX_INTERACTION_TYPE
MA = MAX_MOLE_ATOMS


class MoleculeType:
    def __init__(self, name):
        self.name = name
        self.atom_types = []      # 1-based atom-type indices
        self.bonds, self.angles, self.dihedrals = [], [], []   # 1-based atom indices in molecule
        self.explicit_exclusions = []
        self.reactive_protons = None
        self.reactive_basic_atoms = None
        self.pair_exclusions = None

    @property
    def n_atom(self):
        return len(self.atom_types)


class ForceField:
    """Global force-field arrays of src/glob_v.f90:319-337 and 77-120."""

    def __init__(self):
        self.atype_name = []
        self.atype_chg = np.zeros(T)
        self.atype_freeze = np.zeros(T, np.int32)
        self.atype_mass = -np.ones(T)
        self.vdw_parameter = np.zeros((T, T, 6), order="F")
        self.vdw_parameter_14 = np.zeros((T, T, 6), order="F")
        self.vdw_tmp = np.zeros((T, T, 9), order="F")
        self.vdw_type = np.zeros((T, T), np.int32, order="F")
        self.bond_type = np.zeros((T, T), np.int32, order="F")
        self.bond_parameter = np.zeros((T, T, 3), order="F")
        self.angle_type = np.zeros((T, T, T), np.int32, order="F")
        self.angle_parameter = np.zeros((T, T, T, 2), order="F")
        self.dihedral_type = np.zeros((T, T, T, T), np.int32, order="F")
        self.dihedral_parameter = np.zeros((T, T, T, T, 6), order="F")
        self.molecule_types = []          # list[MoleculeType], index+1 == molecule_type_index
        self.lj_comb_rule = "opls"
        self.n_exclusions = 3
        self.pi = 3.141592654             # constants%pi, glob_v.f90:386
        # MS-EVB
        self.evb_donor_acceptor_interaction = np.zeros((MI, 3), np.int32, order="F")
        self.evb_donor_acceptor_parameters = np.zeros((MI, 6), order="F")
        self.evb_proton_acceptor_interaction = np.zeros((MI, 2), np.int32, order="F")
        self.evb_proton_acceptor_parameters = np.zeros((MI, 5), order="F")
        self.evb_diabat_coupling_interaction = np.zeros((MI, 3), np.int32, order="F")
        self.evb_diabat_coupling_parameters = np.zeros((MI, 10), order="F")
        self.evb_diabat_coupling_type = np.zeros(MI, np.int32)
        self.evb_exchange_charge_atomic = np.zeros(T)
        self.evb_exchange_charge_proton = np.zeros((MT, MT), order="F")
        self.evb_acid_molecule = np.zeros(MT, np.int32)
        self.evb_basic_molecule = np.zeros(MT, np.int32)
        self.evb_conjugate_pairs = np.zeros(MT, np.int32)
        self.evb_conjugate_atom_index = np.zeros(T, np.int32)
        self.evb_reference_energy = np.zeros(MT)
        self.evb_proton_index = np.zeros(MT, np.int32)
        self.evb_heavy_acid_index = np.zeros(MT, np.int32)
        self.has_evb = False

    @property
    def n_atom_type(self):
        return len(self.atype_name)

    def atype(self, name):
        """atype_name_reverse_lookup (general_routines.f90:1679-1700), 1-based."""
        try:
            return self.atype_name.index(name) + 1
        except ValueError:
            raise ValueError("atom type %r doesn't have force field parameters!" % name)

    def mtype(self, name, create=False):
        for i, m in enumerate(self.molecule_types):
            if m.name == name:
                return i + 1
        if create:
            self.molecule_types.append(MoleculeType(name))
            return len(self.molecule_types)
        raise ValueError("unknown molecule type %r" % name)


def _clean_lines(text):
    return text.splitlines()


# ------------------------------------------------------------------------------------------------
# .pmt  (initialize_routines.f90:281-430, 646-691)
# ------------------------------------------------------------------------------------------------
def read_param(ff, text):
    lines = _clean_lines(text)
    i = 0
    gen_cross_terms = np.zeros((T, T), np.int32)
    ff.vdw_tmp[:] = 0.0
    ff.vdw_tmp[:, :, 4] = 3.0          # exp_init, :302-303
    while i < len(lines):
        line = lines[i]
        if "solute_species" in line:
            i += 2                      # heading line is skipped (:328)
            n = int(lines[i].split()[0])
            for _ in range(n):
                i += 1
                a = lines[i].split()
                if len(a) not in (5, 6):
                    raise ValueError("should have 5 input arguments under solute_species section")
                k = len(ff.atype_name)
                ff.atype_name.append(a[0][:5])
                ff.atype_chg[k] = float(a[1])
                ff.vdw_parameter[k, k, 0] = float(a[2])
                ff.vdw_parameter[k, k, 1] = float(a[3])
                ff.atype_freeze[k] = int(float(a[4]))
        elif "custom_sapt_parameters" in line:
            i += 1
            for k in range(ff.n_atom_type):
                i += 1
                a = lines[i].split()
                if len(a) != 10:
                    raise ValueError("should have 10 input arguments under custom_sapt_parameters")
                for q in range(9):
                    ff.vdw_tmp[k, k, q] = float(a[1 + q])
        elif "cross_terms" in line:
            i += 1
            n = int(lines[i].split()[0])
            for _ in range(n):
                i += 1
                a = lines[i].split()
                it, jt = int(a[0]) - 1, int(a[1]) - 1
                s1, s2, s3 = float(a[2]), float(a[3]), float(a[4])
                if ff.lj_comb_rule == "opls":
                    ff.vdw_parameter[it, jt, 0] = ff.vdw_parameter[jt, it, 0] = s2   # C12 first
                    ff.vdw_parameter[it, jt, 1] = ff.vdw_parameter[jt, it, 1] = s1
                else:
                    for q, v in enumerate((s1, s2, s3)):
                        ff.vdw_parameter[it, jt, q] = ff.vdw_parameter[jt, it, q] = v
                    if s1 > 1000.0 or s2 > 1000.0:
                        raise ValueError("looks like combination rule should be opls")
                gen_cross_terms[it, jt] = gen_cross_terms[jt, it] = 1
        i += 1
    if len(set(ff.atype_name)) != len(ff.atype_name):
        raise ValueError("atomic parameters defined more than once")
    if ff.n_atom_type > T:
        raise ValueError("number of atom types g.t. MAX_N_ATOM_TYPE")
    return gen_cross_terms


def _gen_c12_c6(p, i, j):
    eps, sig = p[i, j, 0], p[i, j, 1]
    p[i, j, 0] = 4.0 * eps * sig ** 12
    p[i, j, 1] = 4.0 * eps * sig ** 6


def _combination_rule(ff, i, j):
    p, tmp = ff.vdw_parameter, ff.vdw_tmp
    if ff.vdw_type[i, j] == 1:
        a_ex = np.sqrt(tmp[i, i, 0] * tmp[j, j, 0]); a_el = np.sqrt(tmp[i, i, 1] * tmp[j, j, 1])
        a_ind = np.sqrt(tmp[i, i, 2] * tmp[j, j, 2]); a_dhf = np.sqrt(tmp[i, i, 3] * tmp[j, j, 3])
        p[i, j, 0] = a_ex - a_el - a_ind - a_dhf
        bi, bj = tmp[i, i, 4], tmp[j, j, 4]
        p[i, j, 1] = (bi + bj) * bi * bj / (bi ** 2 + bj ** 2)
        for q in range(4):
            p[i, j, 2 + q] = np.sqrt(tmp[i, i, 5 + q] * tmp[j, j, 5 + q])
    elif ff.lj_comb_rule == "standard":
        p[i, j, 0] = np.sqrt(p[i, i, 0] * p[j, j, 0])
        p[i, j, 1] = (p[i, i, 1] + p[j, j, 1]) / 2.0
    elif ff.lj_comb_rule == "opls":
        p[i, j, 0] = np.sqrt(p[i, i, 0] * p[j, j, 0])
        p[i, j, 1] = np.sqrt(p[i, i, 1] * p[j, j, 1])
    else:
        raise ValueError("lj_comb_rule parameter isn't recognized")


def gen_param(ff, gen_cross_terms):
    """initialize_routines.f90:448-558 (without the per-atom index fill, done by the system builder)."""
    small = 1e-6
    n = ff.n_atom_type
    p, tmp = ff.vdw_parameter, ff.vdw_tmp
    if ff.lj_comb_rule == "opls":
        for i in range(n):
            _gen_c12_c6(p, i, i)
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            if gen_cross_terms[i, j] == 0:
                if p[i, i, 0] > small and p[j, j, 0] > small:
                    ff.vdw_type[i, j] = 0
                elif tmp[i, i, 4] > small and tmp[j, j, 4] > small:
                    ff.vdw_type[i, j] = 1
                else:
                    ff.vdw_type[i, j] = -1
                _combination_rule(ff, i, j)
            else:
                ff.vdw_type[i, j] = 0
    for i in range(n):
        if gen_cross_terms[i, i] == 0:
            if p[i, i, 0] > small:
                ff.vdw_type[i, i] = 0
            elif tmp[i, i, 0] > small:
                ff.vdw_type[i, i] = 1
            else:
                ff.vdw_type[i, i] = -1
            _combination_rule(ff, i, i)
        else:
            ff.vdw_type[i, i] = 0
    if ff.lj_comb_rule == "standard":
        for i in range(n):
            for j in range(n):
                if ff.vdw_type[i, j] == 0:
                    _gen_c12_c6(p, i, j)


def read_14_parameters(ff, text):
    """read_generate_14_interaction_parameters initialize_routines.f90:646-691."""
    ff.vdw_parameter_14[:] = ff.vdw_parameter
    lines = _clean_lines(text)
    for i, line in enumerate(lines):
        if "pairtypes" in line:
            n = int(lines[i + 1].split()[0])
            for k in range(n):
                a = lines[i + 2 + k].split()
                i1, i2 = ff.atype(a[0]) - 1, ff.atype(a[1]) - 1
                c6, c12 = float(a[2]), float(a[3])
                ff.vdw_parameter_14[i1, i2, 0] = ff.vdw_parameter_14[i2, i1, 0] = c12
                ff.vdw_parameter_14[i1, i2, 1] = ff.vdw_parameter_14[i2, i1, 1] = c6
            break


# ------------------------------------------------------------------------------------------------
# .top  (intra_bonded_interactions.f90:696-1464 ; ms_evb.f90:3170-3623)
# ------------------------------------------------------------------------------------------------
class _TopReader:
    """read_topology_line semantics: ';' lines are comments, a blank line ends a section."""

    def __init__(self, text):
        self.lines = text.splitlines()
        self.pos = 0

    def find_heading(self, heading, start=None):
        if start is not None:
            self.pos = start
        while self.pos < len(self.lines):
            line = self.lines[self.pos]
            self.pos += 1
            if heading in line and not line.lstrip().startswith(";"):
                return True
        return False

    def section_lines(self):
        out = []
        while self.pos < len(self.lines):
            raw = self.lines[self.pos]
            self.pos += 1
            s = raw.strip()
            if s.startswith(";") or s.startswith("!"):
                continue
            if s == "":
                break
            if s.startswith("["):
                self.pos -= 1
                break
            out.append(s.split(";")[0].split())
        return out


def read_topology(ff, text, molecule_type_order=()):
    for name in molecule_type_order:      # types present in the coordinate file come first
        ff.mtype(name, create=True)
    r = _TopReader(text)
    if not r.find_heading("[ bondtypes ]"):
        raise ValueError("couldn't find '[ bondtypes ]' section in topology file!")
    for a in r.section_lines():
        i1, i2, bt = ff.atype(a[0]) - 1, ff.atype(a[1]) - 1, int(a[2])
        ff.bond_type[i1, i2] = ff.bond_type[i2, i1] = bt
        npar = 3 if bt == 3 else 2
        for q in range(npar):
            ff.bond_parameter[i1, i2, q] = ff.bond_parameter[i2, i1, q] = float(a[3 + q])
    if not r.find_heading("[ angletypes ]"):
        raise ValueError("couldn't find '[ angletypes ]' section in topology file!")
    for a in r.section_lines():
        i1, i2, i3 = (ff.atype(x) - 1 for x in a[:3])
        at = int(a[3])
        th0 = float(a[4]) * ff.pi / 180.0
        cth = float(a[5])
        ff.angle_type[i1, i2, i3] = ff.angle_type[i3, i2, i1] = at
        ff.angle_parameter[i1, i2, i3, 0] = ff.angle_parameter[i3, i2, i1, 0] = th0
        ff.angle_parameter[i1, i2, i3, 1] = ff.angle_parameter[i3, i2, i1, 1] = cth
    if not r.find_heading("[ dihedraltypes ]"):
        raise ValueError("couldn't find '[ dihedraltypes ]' section in topology file!")
    for a in r.section_lines():
        i1, i2, i3, i4 = (ff.atype(x) - 1 for x in a[:4])
        dt = int(a[4])
        ff.dihedral_type[i1, i2, i3, i4] = ff.dihedral_type[i4, i3, i2, i1] = dt
        if dt == 3:
            vals = [float(x) for x in a[5:11]]
        else:
            vals = [float(a[5]) * ff.pi / 180.0, float(a[6])]
            if dt == 1:
                vals.append(float(a[7]))
        for q, v in enumerate(vals):
            ff.dihedral_parameter[i1, i2, i3, i4, q] = ff.dihedral_parameter[i4, i3, i2, i1, q] = v
    # molecule types
    while r.find_heading("[ moleculetype ]"):
        head = r.section_lines()
        mt = ff.molecule_types[ff.mtype(head[0][0][:5], create=True) - 1]
        mt.atom_types, mt.bonds, mt.angles, mt.dihedrals, mt.explicit_exclusions = [], [], [], [], []
        # sub-sections until the next [ moleculetype ] / evb section
        while r.pos < len(r.lines):
            save = r.pos
            line = r.lines[r.pos].strip()
            r.pos += 1
            if not line.startswith("["):
                continue
            key = line.strip("[] ").strip()
            if key == "atoms":
                for a in r.section_lines():
                    t = ff.atype(a[1][:5])
                    mt.atom_types.append(t)
                    ff.atype_mass[t - 1] = float(a[2])
            elif key == "bonds":
                mt.bonds = [[int(a[0]), int(a[1])] for a in r.section_lines()]
            elif key == "angles":
                mt.angles = [[int(a[0]), int(a[1]), int(a[2])] for a in r.section_lines()]
            elif key == "dihedrals":
                mt.dihedrals = [[int(a[0]), int(a[1]), int(a[2]), int(a[3])] for a in r.section_lines()]
            elif key == "exclusions":
                mt.explicit_exclusions = [[int(x) for x in a] for a in r.section_lines()]
            else:
                r.pos = save
                break
        if mt.n_atom + 1 > MA:
            raise ValueError("molecule type %s too large for RPB_MAX_MOLE_ATOMS" % mt.name)
    if len(ff.molecule_types) > MT:
        raise ValueError("too many molecule types")
    for mt in ff.molecule_types:
        if mt.n_atom == 0:
            raise ValueError("couldn't find '[ moleculetype ]' section for moleculetype %s" % mt.name)


def generate_intramolecular_exclusions(ff):
    """intra_bonded_interactions.f90:574-679 (recursive bond walk, depth max(n_exclusions,3))."""
    nex = ff.n_exclusions
    max_search = max(nex, 3)
    for mt in ff.molecule_types:
        n = mt.n_atom
        ex = np.zeros((MA, MA), np.int32)
        for e in mt.explicit_exclusions:
            for b in e[1:]:
                ex[e[0] - 1, b - 1] = ex[b - 1, e[0] - 1] = 1
        bonded = np.zeros((n, n), bool)
        for i, j in mt.bonds:
            bonded[i - 1, j - 1] = bonded[j - 1, i - 1] = True

        def search(i_atom, j_atom, n_bonds_in, traj):
            for local in range(n):
                if bonded[j_atom, local] and local not in traj[:n_bonds_in]:
                    if n_bonds_in == 3 and nex < 3 and ex[i_atom, local] != 1:
                        ex[i_atom, local] = 2
                    else:
                        ex[i_atom, local] = 1
                    n_out = n_bonds_in + 1
                    if n_out <= max_search:
                        search(i_atom, local, n_out, traj[:n_bonds_in] + [local])

        for i_atom in range(n):
            ex[i_atom, i_atom] = 1
            search(i_atom, i_atom, 1, [i_atom])
        mt.pair_exclusions = ex


def read_evb(ff, text):
    """read_evb_parameters / read_evb_topology  ms_evb.f90:3170-3623."""
    r = _TopReader(text)
    if not r.find_heading("[ evb_parameters ]", 0):
        return
    if not r.find_heading("[ reference_energy ]"):
        raise ValueError("missing [ reference_energy ]")
    for a in r.section_lines():
        ff.evb_reference_energy[ff.mtype(a[0][:5]) - 1] = float(a[1])
    r.find_heading("[ donor_acceptor ]")
    for k, a in enumerate(r.section_lines()):
        if len(a) != 9:
            raise ValueError("must have 9 arguments in 'donor_acceptor' section")
        for q in range(3):
            ff.evb_donor_acceptor_interaction[k, q] = ff.atype(a[q])
        for q in range(6):
            ff.evb_donor_acceptor_parameters[k, q] = float(a[3 + q])
    r.find_heading("[ proton_acceptor ]")
    for k, a in enumerate(r.section_lines()):
        if len(a) != 7:
            raise ValueError("must have 7 arguments in 'proton_acceptor' section")
        for q in range(2):
            ff.evb_proton_acceptor_interaction[k, q] = ff.atype(a[q])
        for q in range(5):
            ff.evb_proton_acceptor_parameters[k, q] = float(a[2 + q])
    r.find_heading("[ geometry_factor ]")
    rows = r.section_lines()
    k = 0
    q = 0
    while q < len(rows):
        a = rows[q]
        if len(a) == 4:           # reference format: 3 types + function type, parameters on the next line
            ftype, pars = int(a[3]), [float(x) for x in rows[q + 1]]
            q += 2
        else:                     # shipped example: 3 types + 10 MS-EVB3 parameters on one line
            ftype, pars = 1, [float(x) for x in a[3:]]
            q += 1
        if (ftype == 1 and len(pars) != 10) or (ftype == 2 and len(pars) != 4):
            raise ValueError("wrong number of diabat_coupling parameters")
        for z in range(3):
            ff.evb_diabat_coupling_interaction[k, z] = ff.atype(a[z])
        ff.evb_diabat_coupling_type[k] = ftype
        ff.evb_diabat_coupling_parameters[k, :len(pars)] = pars
        k += 1
    r.find_heading("[ exchange_charge_atomic ]")
    for a in r.section_lines():
        ff.evb_exchange_charge_atomic[ff.atype(a[0]) - 1] = float(a[1])
    r.find_heading("[ exchange_charge_proton ]")
    for a in r.section_lines():
        i1, i2 = ff.mtype(a[0][:5]) - 1, ff.mtype(a[1][:5]) - 1
        ff.evb_exchange_charge_proton[i1, i2] = ff.evb_exchange_charge_proton[i2, i1] = float(a[2])
    # topology
    if not r.find_heading("[ evb_topology ]", 0):
        raise ValueError("missing [ evb_topology ]")
    while r.find_heading("[ evb_pairs ]"):
        a = r.section_lines()[0]
        acid, base = ff.mtype(a[0][:5]), ff.mtype(a[1][:5])
        ff.evb_acid_molecule[acid - 1] = 1
        ff.evb_basic_molecule[base - 1] = 1
        ff.evb_conjugate_pairs[acid - 1] = base
        ff.evb_conjugate_pairs[base - 1] = acid
        ff.evb_proton_index[acid - 1] = ff.atype(a[2])
        ff.evb_heavy_acid_index[acid - 1] = ff.atype(a[3])
        ma, mb = ff.molecule_types[acid - 1], ff.molecule_types[base - 1]

        def flags(mt):
            f = np.zeros(MA, np.int32)
            for row in r.section_lines():
                f[int(row[0]) - 1] = int(row[1])
            return f
        r.find_heading("[ acid_reactive_protons ]"); ma.reactive_protons = flags(ma)
        r.find_heading("[ base_reactive_protons ]"); mb.reactive_protons = flags(mb)
        r.find_heading("[ acid_acceptor_atoms ]"); ma.reactive_basic_atoms = flags(ma)
        r.find_heading("[ base_acceptor_atoms ]"); mb.reactive_basic_atoms = flags(mb)
        r.find_heading("[ conjugate_atoms ]")
        for row in r.section_lines():
            i1, i2 = ff.atype(row[0]), ff.atype(row[1])
            ff.evb_conjugate_atom_index[i1 - 1] = i2
            ff.evb_conjugate_atom_index[i2 - 1] = i1
    # evb_consistency_checks ms_evb.f90:144-167: acidic protons last
    for i, mt in enumerate(ff.molecule_types):
        if ff.evb_acid_molecule[i] == 1:
            seen = False
            for a in range(mt.n_atom):
                if mt.reactive_protons[a] == 1:
                    seen = True
                elif seen:
                    raise ValueError("acidic protons must be defined last in the molecule topology")
    ff.has_evb = True


def load_forcefield(pmt_text, top_text, lj_comb_rule="opls", n_exclusions=3, molecule_type_order=()):
    ff = ForceField()
    ff.lj_comb_rule = lj_comb_rule
    ff.n_exclusions = n_exclusions
    gct = read_param(ff, pmt_text)
    gen_param(ff, gct)
    read_14_parameters(ff, pmt_text)
    read_topology(ff, top_text, molecule_type_order)
    generate_intramolecular_exclusions(ff)
    read_evb(ff, top_text)
    return ff


def flatten_molecule_types(ff):
    """Arrays for rpb_set_molecule_types (include/rpbmd.h)."""
    nt = len(ff.molecule_types)
    n_atom = np.zeros(MT, np.int32)
    atom_type = np.zeros(MT * MA, np.int32)
    n_bond = np.zeros(MT, np.int32); n_angle = np.zeros(MT, np.int32); n_dih = np.zeros(MT, np.int32)
    bonds, angles, dihs = [], [], []
    excl = np.zeros(MT * MA * MA, np.int32)
    rp = np.zeros(MT * MA, np.int32); rb = np.zeros(MT * MA, np.int32)
    for t, mt in enumerate(ff.molecule_types):
        n_atom[t] = mt.n_atom
        atom_type[t * MA:t * MA + mt.n_atom] = mt.atom_types
        n_bond[t], n_angle[t], n_dih[t] = len(mt.bonds), len(mt.angles), len(mt.dihedrals)
        bonds += [x for b in mt.bonds for x in b]
        angles += [x for b in mt.angles for x in b]
        dihs += [x for b in mt.dihedrals for x in b]
        excl[t * MA * MA:(t + 1) * MA * MA] = mt.pair_exclusions.flatten(order="F")
        if mt.reactive_protons is not None:
            rp[t * MA:(t + 1) * MA] = mt.reactive_protons
        if mt.reactive_basic_atoms is not None:
            rb[t * MA:(t + 1) * MA] = mt.reactive_basic_atoms
    as_i = lambda x: np.asarray(x if len(x) else [0], np.int32)
    return dict(n_atom=n_atom, atom_type=atom_type, n_bond=n_bond, bonds=as_i(bonds), n_angle=n_angle,
                angles=as_i(angles), n_dihedral=n_dih, dihedrals=as_i(dihs), pair_exclusions=excl,
                reactive_protons=rp, reactive_basic_atoms=rb, n_types=nt)
