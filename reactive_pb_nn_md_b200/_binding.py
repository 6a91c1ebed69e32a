"""ctypes binding of the C-ABI declared in include/rpbmd.h.

The same binding drives any shared library that exports that ABI.  The product library is
``csrc/librpbmd.so`` (CUDA, sm_100a); :func:`load_cuda` fails loudly when it is missing --
there is no CPU fallback in this package.  (tests/ additionally load the CPU oracle through
:class:`Library` by explicit path; nothing in this package does.)
"""
import ctypes as C
import os

import numpy as np

MAX_N_ATOM_TYPE = 25      # src/glob_v.f90:34
MAX_N_MOLE_TYPE = 10      # src/glob_v.f90:34
MAX_INTERACTION_TYPE = 15  # src/glob_v.f90:72
EVB_MAX_STATES = 80       # src/glob_v.f90:60
EVB_MAX_CHAIN = 3         # src/glob_v.f90:65
PEER_HANDLE_BYTES = 64    # RPB_PEER_HANDLE_BYTES, include/rpbmd.h
EVB_MAX_NEIGHBORS = 10    # src/glob_v.f90:56
MAX_MOLE_ATOMS = 8

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.path.join(_HERE, "csrc", "librpbmd.so")


class RpbConfig(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int), ("n_mole", C.c_int), ("n_atom_type", C.c_int), ("n_mole_type", C.c_int),
        ("pme_grid", C.c_int), ("spline_order", C.c_int), ("spline_grid", C.c_int), ("erfc_grid", C.c_int),
        ("tt_grid", C.c_int), ("na_nslist", C.c_int), ("nb_nslist", C.c_int), ("nc_nslist", C.c_int),
        ("verlet_capacity", C.c_int), ("device", C.c_int), ("rank", C.c_int), ("world_size", C.c_int),
        ("n_threads", C.c_int), ("evb_max_chain", C.c_int), ("evb_max_states", C.c_int),
        ("reserved_i", C.c_int * 4),
        ("box", C.c_double * 9), ("alpha_sqrt", C.c_double), ("real_space_cutoff", C.c_double),
        ("verlet_cutoff", C.c_double), ("delta_t", C.c_double), ("erfc_dx", C.c_double), ("tt_max", C.c_double),
        ("pi", C.c_double), ("pi_sqrt", C.c_double), ("conv_e2A_kJmol", C.c_double),
        ("conv_kJmol_ang2ps2gmol", C.c_double), ("safe_verlet", C.c_double), ("verlet_thresh", C.c_double),
        ("evb_first_solvation_cutoff", C.c_double), ("evb_reactive_pair_distance", C.c_double),
        ("ewald_self", C.c_double), ("reserved_d", C.c_double * 4),
    ]


class RpbEnergies(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "potential_energy", "kinetic_energy", "E_elec", "E_vdw", "E_bond", "E_angle", "E_dihedral", "E_recip")]


class RpbError(RuntimeError):
    """Raised where the reference would `stop "message"`; carries the rpb_status code."""

    def __init__(self, code, msg):
        super().__init__("rpb error %d: %s" % (code, msg))
        self.code = code


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

_PROTOS = {
    "rpb_last_error": (C.c_char_p, [_vp]),
    "rpb_backend": (C.c_char_p, []),
    "rpb_create": (C.c_int, [C.POINTER(_vp), C.POINTER(RpbConfig)]),
    "rpb_destroy": (None, [_vp]),
    "rpb_set_tables": (C.c_int, [_vp] + [_dp] * 7),
    "rpb_set_forcefield": (C.c_int, [_vp, _dp, _ip, _dp, _dp, _ip, _ip, _dp, _ip, _dp, _ip, _dp]),
    "rpb_set_molecule_types": (C.c_int, [_vp] + [_ip] * 11),
    "rpb_set_evb": (C.c_int, [_vp, _ip, _dp, _ip, _dp, _ip, _dp, _ip, _dp, _dp, _ip, _ip, _ip, _ip, _dp, _ip, _ip]),
    "rpb_upload_state": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _ip, _ip, _ip, _ip, C.c_int]),
    "rpb_initialize": (C.c_int, [_vp]),
    "rpb_force_energy": (C.c_int, [_vp, C.c_int]),
    "rpb_step": (C.c_int, [_vp, C.c_int, C.c_int]),
    "rpb_ensemble_step": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int]),
    "rpb_step_begin": (C.c_int, [_vp]),
    "rpb_step_end": (C.c_int, [_vp]),
    "rpb_evb_phase_build": (C.c_int, [_vp]),
    "rpb_evb_phase_mix": (C.c_int, [_vp]),
    "rpb_evb_phase_commit": (C.c_int, [_vp]),
    "rpb_evb_exchange_h": (C.c_int, [_vp, C.POINTER(_vp), _ip]),
    "rpb_evb_exchange_f": (C.c_int, [_vp, C.POINTER(_vp), _ip]),
    "rpb_peer_export": (C.c_int, [_vp, _vp]),
    "rpb_peer_import": (C.c_int, [_vp, _vp, C.c_int]),
    "rpb_peer_attach_local": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "rpb_peer_enabled": (C.c_int, [_vp]),
    "rpb_get_energies": (C.c_int, [_vp, C.POINTER(RpbEnergies)]),
    "rpb_download_state": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _ip, _ip, _ip]),
    "rpb_get_r_com": (C.c_int, [_vp, _dp]),
    "rpb_get_neighbor_list": (C.c_int, [_vp, _ip, _ip, C.c_int, _ip, _ip]),
    "rpb_debug_tile_pairs": (C.c_int, [_vp, _ip, _ip, C.c_longlong, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "rpb_get_pme": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp]),
    "rpb_get_evb": (C.c_int, [_vp, _ip, _dp, _dp, _ip, _ip, _ip, _ip, _dp]),
    "rpb_debug_mix_forces": (C.c_int, [_vp, _dp, _dp]),
    "rpb_get_launch_counts": (C.c_int, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "rpb_timers_enable": (C.c_int, [_vp, C.c_int]),
    "rpb_timers_reset": (C.c_int, [_vp]),
    "rpb_timer_count": (C.c_int, []),
    "rpb_timer_name": (C.c_char_p, [C.c_int]),
    "rpb_timers_get": (C.c_int, [_vp, _dp, C.POINTER(C.c_longlong)]),
    "rpb_get_stream": (_vp, [_vp]),
    "rpb_measure_fp64_peak": (C.c_int, [_vp, _dp]),
}

ABI_SYMBOLS = tuple(sorted(_PROTOS))


class Library:
    """A loaded shared library exporting the rpbmd C-ABI."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s not found -- build it first (python -c 'import __graft_entry__ as g; g.build()')" % path)
        self.path = path
        self.dll = C.CDLL(path, mode=C.RTLD_LOCAL)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(self.dll, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        self.backend = self.dll.rpb_backend().decode()


def load_cuda():
    """Load the CUDA product library; never falls back to anything else."""
    return Library(CUDA_LIB_PATH)


def dptr(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and (a.flags.c_contiguous or a.flags.f_contiguous)
    return a.ctypes.data_as(_dp)


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and (a.flags.c_contiguous or a.flags.f_contiguous)
    return a.ctypes.data_as(_ip)
