"""Host-side mirror of the reference's force-path interface, on top of the C-ABI.

Method names follow the reference routines they stand for:
  calculate_total_force_energy          src/total_energy_forces.f90:19-99
  ms_evb_calculate_total_force_energy   src/ms_evb.f90:181-235
  md_integrate_atomic                   src/md_integration.f90:438-541 (NVE)
  calculate_kinetic_energy              src/total_energy_forces.f90:106-121
Errors the reference reports with `stop "..."` surface as RpbError.

Diabatic-state sharding: with a torch.distributed process group the two
exchange steps of the sharded MS-EVB step (Hamiltonian elements, Hellmann-Feynman partial forces)
are set up here.  CUDA library: the ranks' exchange arenas are mapped into each other over CUDA IPC
once, after which the whole step -- including the two all-reduces, one peer-memory kernel each --
runs inside rpb_step (exchange == "peer").  Otherwise (RPB_EXCHANGE=collective, no peer access, or
the CPU oracle under gloo) the collectives run here on the library's exchange buffers.
"""
import ctypes as C
import os
import sys

import numpy as np

from . import tables
from ._binding import (EVB_MAX_CHAIN, EVB_MAX_STATES, PEER_HANDLE_BYTES, Library, RpbConfig, RpbEnergies, RpbError, dptr, iptr,
                       load_cuda)
from .forcefield import flatten_molecule_types


class SimulationParameters:
    """The numeric keywords of the simulation-parameter file (read_simulation_parameters.f90:46-146)
    that the force path consumes."""

    def __init__(self, delta_t=0.0005, real_space_cutoff=10.0, verlet_cutoff=12.0, na_nslist=10, nb_nslist=10,
                 nc_nslist=10, alpha_sqrt=0.3, pme_grid=60, spline_order=6, n_threads=1):
        self.delta_t = delta_t
        self.real_space_cutoff = real_space_cutoff
        self.verlet_cutoff = verlet_cutoff
        self.na_nslist, self.nb_nslist, self.nc_nslist = na_nslist, nb_nslist, nc_nslist
        self.alpha_sqrt = alpha_sqrt
        self.pme_grid = pme_grid
        self.spline_order = spline_order
        self.n_threads = n_threads


_TABLE_CACHE = {}


def _cached(key, fn):
    if key not in _TABLE_CACHE:
        _TABLE_CACHE[key] = fn()
    return _TABLE_CACHE[key]


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias a library-owned device buffer."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class Simulation:
    def __init__(self, system, params, library=None, device=0, rank=0, world_size=1, process_group=None,
                 evb_max_states=EVB_MAX_STATES, verlet_capacity=0):
        self.lib = library if library is not None else load_cuda()
        assert isinstance(self.lib, Library)
        self.dll = self.lib.dll
        self.system = system
        self.params = params
        self.rank, self.world_size, self.pg = rank, world_size, process_group
        ff = system.ff
        cfg = RpbConfig()
        cfg.n_atoms, cfg.n_mole = system.n_atoms, system.n_mole
        cfg.n_atom_type, cfg.n_mole_type = ff.n_atom_type, len(ff.molecule_types)
        cfg.pme_grid, cfg.spline_order = params.pme_grid, params.spline_order
        cfg.spline_grid, cfg.erfc_grid, cfg.tt_grid = tables.SPLINE_GRID, tables.ERFC_GRID, tables.TT_GRID
        cfg.na_nslist, cfg.nb_nslist, cfg.nc_nslist = params.na_nslist, params.nb_nslist, params.nc_nslist
        cfg.verlet_capacity = int(verlet_capacity)      # 0: the reference's size formula (general_routines.f90:1231-1239)
        cfg.device, cfg.rank, cfg.world_size = device, rank, world_size
        cfg.n_threads = params.n_threads
        cfg.evb_max_chain, cfg.evb_max_states = EVB_MAX_CHAIN, int(evb_max_states)    # glob_v.f90:60,65
        for i, v in enumerate(system.box.flatten(order="F")):
            cfg.box[i] = v
        cfg.alpha_sqrt, cfg.real_space_cutoff = params.alpha_sqrt, params.real_space_cutoff
        cfg.verlet_cutoff, cfg.delta_t = params.verlet_cutoff, params.delta_t
        erfc_dx, erfc_t, scale_t = _cached(("ew", params.real_space_cutoff, params.alpha_sqrt),
                                           lambda: tables.ewald_tables(params.real_space_cutoff, params.alpha_sqrt))
        cfg.erfc_dx, cfg.tt_max = erfc_dx, tables.TT_MAX
        cfg.pi, cfg.pi_sqrt = tables.PI, tables.PI_SQRT
        cfg.conv_e2A_kJmol, cfg.conv_kJmol_ang2ps2gmol = tables.CONV_E2A_KJMOL, tables.CONV_KJMOL
        cfg.safe_verlet, cfg.verlet_thresh = tables.SAFE_VERLET, tables.VERLET_THRESH
        cfg.evb_first_solvation_cutoff, cfg.evb_reactive_pair_distance = 5.0, 2.5   # glob_v.f90:54-55
        cfg.ewald_self = tables.ewald_self(system.charge, params.alpha_sqrt)      # frozen at init
        self.cfg = cfg
        self.ctx = C.c_void_p()
        rc = self.dll.rpb_create(C.byref(self.ctx), C.byref(cfg))
        self._check(rc)
        B6, B5 = _cached(("bs",), tables.spline_tables)
        tt, dtt = _cached(("tt",), tables.tang_toennies_tables)
        CB = _cached(("cb", system.box_length, params.pme_grid, params.alpha_sqrt),
                     lambda: tables.cb_array(system.box_length, params.pme_grid, params.alpha_sqrt))
        self._check(self.dll.rpb_set_tables(self.ctx, dptr(B6), dptr(B5), dptr(erfc_t), dptr(scale_t), dptr(tt),
                                            dptr(dtt), dptr(CB)))
        self._check(self.dll.rpb_set_forcefield(
            self.ctx, dptr(ff.vdw_parameter), iptr(ff.vdw_type), dptr(ff.vdw_parameter_14), dptr(ff.atype_chg),
            iptr(ff.atype_freeze), iptr(ff.bond_type), dptr(ff.bond_parameter), iptr(ff.angle_type),
            dptr(ff.angle_parameter), iptr(ff.dihedral_type), dptr(ff.dihedral_parameter)))
        m = flatten_molecule_types(ff)
        self._check(self.dll.rpb_set_molecule_types(
            self.ctx, iptr(m["n_atom"]), iptr(m["atom_type"]), iptr(m["n_bond"]), iptr(m["bonds"]), iptr(m["n_angle"]),
            iptr(m["angles"]), iptr(m["n_dihedral"]), iptr(m["dihedrals"]), iptr(m["pair_exclusions"]),
            iptr(m["reactive_protons"]), iptr(m["reactive_basic_atoms"])))
        if ff.has_evb:
            self._check(self.dll.rpb_set_evb(
                self.ctx, iptr(ff.evb_donor_acceptor_interaction), dptr(ff.evb_donor_acceptor_parameters),
                iptr(ff.evb_proton_acceptor_interaction), dptr(ff.evb_proton_acceptor_parameters),
                iptr(ff.evb_diabat_coupling_interaction), dptr(ff.evb_diabat_coupling_parameters),
                iptr(ff.evb_diabat_coupling_type), dptr(ff.evb_exchange_charge_atomic),
                dptr(ff.evb_exchange_charge_proton), iptr(ff.evb_acid_molecule), iptr(ff.evb_basic_molecule),
                iptr(ff.evb_conjugate_pairs), iptr(ff.evb_conjugate_atom_index), dptr(ff.evb_reference_energy),
                iptr(ff.evb_proton_index), iptr(ff.evb_heavy_acid_index)))
        self.upload_state(system.xyz, system.velocity)
        self._check(self.dll.rpb_initialize(self.ctx))
        self._xh = self._xf = None
        self._ext_stream = None
        self.exchange = "none" if world_size == 1 else "collective"
        if (world_size > 1 and ff.has_evb and self.lib.backend.startswith("cuda") and process_group is not None
                and os.environ.get("RPB_EXCHANGE", "peer") == "peer"):
            self._setup_peer_exchange()

    # -- peer-memory exchange (NVLink / NVSwitch) --------------------------------------------
    def _setup_peer_exchange(self):
        """One process per GPU: all-gather the 64-byte IPC handles of the ranks' exchange arenas over the process group
        and map the peers' arenas.  From then on md_integrate_atomic / ms_evb_calculate_total_force_energy run inside
        the library (rpb_step / rpb_force_energy); torch.distributed is no longer on the step's path.  If the arenas
        cannot be mapped (no peer access between the devices), every rank keeps the collective path."""
        import torch
        import torch.distributed as dist
        hb = C.create_string_buffer(PEER_HANDLE_BYTES)
        self._check(self.dll.rpb_peer_export(self.ctx, hb))
        # the handles travel on whatever the process group moves: device tensors for NCCL, host tensors otherwise
        dev = torch.device("cuda", self.cfg.device) if dist.get_backend(self.pg) == "nccl" else torch.device("cpu")
        mine = torch.frombuffer(bytearray(hb.raw), dtype=torch.uint8).to(dev)
        every = [torch.empty_like(mine) for _ in range(self.world_size)]
        dist.all_gather(every, mine, group=self.pg)
        blob = b"".join(t.cpu().numpy().tobytes() for t in every)
        rc = self.dll.rpb_peer_import(self.ctx, C.c_char_p(blob), self.world_size)
        ok = torch.tensor([1 if rc == 0 else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.pg)
        if int(ok.item()) == 1:
            self.exchange = "peer"
        else:
            msg = self.dll.rpb_last_error(self.ctx)
            sys.stderr.write("[rpbmd] peer-memory exchange unavailable (%s); using torch.distributed all-reduce\n"
                             % (msg.decode() if msg else "a peer failed"))

    @staticmethod
    def attach_local(sims):
        """Several ranks inside ONE process, each context on its OWN device (contexts in rank order): map the arenas
        directly.  (Ranks sharing a device inside one process would deadlock: a kernel waiting for its peer keeps the
        peer's cooperative neighbour-list kernel from ever becoming co-resident.)"""
        arr = (C.c_void_p * len(sims))(*[s.ctx for s in sims])
        sims[0]._check(sims[0].dll.rpb_peer_attach_local(arr, len(sims)))
        for s in sims:
            s.exchange = "peer"

    @staticmethod
    def ensemble_step(sims, n_steps=1, ms_evb=False):
        """md_integrate_atomic on independent replicas sharing one device (BASELINE config 5): one host thread per
        replica inside the library, so their kernels overlap."""
        arr = (C.c_void_p * len(sims))(*[s.ctx for s in sims])
        rc = sims[0].dll.rpb_ensemble_step(arr, len(sims), int(n_steps), int(bool(ms_evb)))
        if rc != 0:
            for s in sims:
                msg = s.dll.rpb_last_error(s.ctx)
                if msg:
                    raise RpbError(rc, msg.decode())
            raise RpbError(rc, "")

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self.dll.rpb_last_error(self.ctx)
            raise RpbError(rc, msg.decode() if msg else "")

    def close(self):
        if self.ctx:
            self.dll.rpb_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    _STATE_KEYS = ("xyz", "velocity", "force", "mass", "charge", "atom_type", "mol_first_atom", "mol_n_atom", "mol_type")

    def _state_pointers(self, holder):
        """ctypes pointers of a state dict's arrays, cached per dict while it keeps holding the SAME array objects (a
        caller that owns its arrays passes the same ones every step; building nine ctypes pointers costs ~15 us)."""
        cache = self.__dict__.setdefault("_ptr_cache", {})
        ent = cache.get(id(holder))
        if ent is not None and ent[0] is holder and all(holder.get(k) is a for k, a in zip(self._STATE_KEYS, ent[1])):
            return ent[2]
        arrs = tuple(holder.get(k) for k in self._STATE_KEYS)
        ptrs = tuple(None if a is None else (dptr(a) if a.dtype == np.float64 else iptr(a)) for a in arrs)
        if len(cache) >= 4:
            cache.clear()
        cache[id(holder)] = (holder, arrs, ptrs)      # (strong references: the ids cannot be recycled while cached)
        return ptrs

    def upload_state(self, xyz, velocity, topology=None):
        """atom_data / molecule_data -> library.  `topology` (a dict as returned by download_state) carries the
        per-atom and per-molecule arrays after proton hops have permuted them; default: the initial system."""
        s = self.system
        t = topology if topology is not None else dict(
            mass=s.mass, charge=s.charge, atom_type=s.atom_type, mol_first_atom=s.mol_first_atom,
            mol_n_atom=s.mol_n_atom, mol_type=s.mol_type, hydronium_mol=s.hydronium_mol)
        if topology is not None and xyz is topology.get("xyz") and velocity is topology.get("velocity") and "force" in topology:
            p = self._state_pointers(topology)        # a dict from download_state handed back unchanged in identity
            self._check(self.dll.rpb_upload_state(self.ctx, p[0], p[1], p[3], p[4], p[5], p[6], p[7], p[8], int(t["hydronium_mol"])))
            return
        xyz = np.ascontiguousarray(xyz, np.float64)
        velocity = np.ascontiguousarray(velocity, np.float64)
        self._check(self.dll.rpb_upload_state(
            self.ctx, dptr(xyz), dptr(velocity), dptr(t["mass"]), dptr(t["charge"]), iptr(t["atom_type"]),
            iptr(t["mol_first_atom"]), iptr(t["mol_n_atom"]), iptr(t["mol_type"]), int(t["hydronium_mol"])))

    # -- the reference interface ------------------------------------------------------------
    def calculate_total_force_energy(self):
        self._check(self.dll.rpb_force_energy(self.ctx, 0))

    def ms_evb_calculate_total_force_energy(self):
        if self.world_size == 1 or self.exchange == "peer":
            self._check(self.dll.rpb_force_energy(self.ctx, 1))
            return
        self._check(self.dll.rpb_evb_phase_build(self.ctx))
        self._allreduce("h")
        self._check(self.dll.rpb_evb_phase_mix(self.ctx))
        self._allreduce("f")
        self._check(self.dll.rpb_evb_phase_commit(self.ctx))

    def md_integrate_atomic(self, n_steps=1, ms_evb=False):
        if self.world_size == 1 or self.exchange == "peer" or not ms_evb:
            self._check(self.dll.rpb_step(self.ctx, int(n_steps), int(bool(ms_evb))))
            return
        for _ in range(n_steps):
            self._check(self.dll.rpb_step_begin(self.ctx))
            if ms_evb:
                self.ms_evb_calculate_total_force_energy()
            else:
                self.calculate_total_force_energy()
            self._check(self.dll.rpb_step_end(self.ctx))

    def _exchange_tensor(self, which):
        import torch
        ptr, n = C.c_void_p(), C.c_int()
        fn = self.dll.rpb_evb_exchange_h if which == "h" else self.dll.rpb_evb_exchange_f
        self._check(fn(self.ctx, C.byref(ptr), C.byref(n)))
        if self.lib.backend.startswith("cuda"):
            return torch.as_tensor(_DevArray(ptr.value, n.value), device="cuda:%d" % self.cfg.device)
        buf = (C.c_double * n.value).from_address(ptr.value)
        return torch.from_numpy(np.frombuffer(buf, dtype=np.float64))

    def _allreduce(self, which):
        import torch.distributed as dist
        t = self._exchange_tensor(which)
        if t.is_cuda:
            # Stream-ordered: with the library's main stream as torch's current stream, NCCL's collective waits for the
            # kernels that filled the buffer and the library's next kernels wait for the collective -- no host sync.
            import torch
            if self._ext_stream is None:
                self._ext_stream = torch.cuda.ExternalStream(self.dll.rpb_get_stream(self.ctx),
                                                             device=torch.device("cuda", self.cfg.device))
            with torch.cuda.stream(self._ext_stream):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)

    # -- results ----------------------------------------------------------------------------
    def energies(self):
        e = RpbEnergies()
        self._check(self.dll.rpb_get_energies(self.ctx, C.byref(e)))
        return {n: getattr(e, n) for n, _ in RpbEnergies._fields_}

    def calculate_kinetic_energy(self):
        return self.energies()["kinetic_energy"]

    def download_state(self, out=None):
        """atom_data / molecule_data <- library.  `out`: a dict returned by an earlier call, whose arrays are reused
        (a caller that owns its arrays, like the Fortran driver, never allocates per step)."""
        N, M = self.system.n_atoms, self.system.n_mole
        if out is None:
            out = dict(xyz=np.zeros((N, 3)), velocity=np.zeros((N, 3)), force=np.zeros((N, 3)), mass=np.zeros(N),
                       charge=np.zeros(N), atom_type=np.zeros(N, np.int32), mol_first_atom=np.zeros(M, np.int32),
                       mol_n_atom=np.zeros(M, np.int32), mol_type=np.zeros(M, np.int32))
        h = C.c_int()
        p = self._state_pointers(out)
        self._check(self.dll.rpb_download_state(self.ctx, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], C.byref(h)))
        out["hydronium_mol"] = h.value
        return out

    def forces(self):
        f = np.zeros((self.system.n_atoms, 3))
        self._check(self.dll.rpb_download_state(self.ctx, None, None, dptr(f), None, None, None, None, None, None, None))
        return f

    def r_com(self):
        r = np.zeros((self.system.n_mole, 3))
        self._check(self.dll.rpb_get_r_com(self.ctx, dptr(r)))
        return r

    def neighbor_list(self):
        N = self.system.n_atoms
        vp = np.zeros(N + 1, np.int32)
        n, flag = C.c_int(), C.c_int()
        self._check(self.dll.rpb_get_neighbor_list(self.ctx, iptr(vp), None, 0, C.byref(n), C.byref(flag)))
        nl = np.zeros(max(n.value, 1), np.int32)
        self._check(self.dll.rpb_get_neighbor_list(self.ctx, iptr(vp), iptr(nl), n.value, C.byref(n), C.byref(flag)))
        return vp, nl[:n.value], flag.value

    def tile_pairs(self):
        """the pair list the force kernel consumes, as sorted (i, j) pairs, i < j, 1-based, and its size in list words"""
        n, nt = C.c_longlong(), C.c_longlong()
        self._check(self.dll.rpb_debug_tile_pairs(self.ctx, None, None, 0, C.byref(n), C.byref(nt)))
        pi = np.zeros(max(n.value, 1), np.int32); pj = np.zeros(max(n.value, 1), np.int32)
        self._check(self.dll.rpb_debug_tile_pairs(self.ctx, iptr(pi), iptr(pj), n.value, C.byref(n), C.byref(nt)))
        return pi[:n.value], pj[:n.value], nt.value

    def pme(self, state=1):
        K, N = self.params.pme_grid, self.system.n_atoms
        Q = np.zeros((K, K, K), order="F"); th = np.zeros((K, K, K), order="F"); fr = np.zeros((N, 3))
        self._check(self.dll.rpb_get_pme(self.ctx, int(state), dptr(Q), dptr(th), dptr(fr)))
        return Q, th, fr

    def evb(self):
        S = C.c_int(); pd = C.c_int(); nh = C.c_int(); ad = C.c_double()
        H = np.zeros((EVB_MAX_STATES, EVB_MAX_STATES), order="F")
        c = np.zeros(EVB_MAX_STATES)
        log = np.zeros((EVB_MAX_STATES, EVB_MAX_CHAIN, 5), np.int32, order="F")
        cm = np.zeros(EVB_MAX_STATES, np.int32)
        self._check(self.dll.rpb_get_evb(self.ctx, C.byref(S), dptr(H), dptr(c), iptr(log), iptr(cm), C.byref(pd),
                                         C.byref(nh), C.byref(ad)))
        n = S.value
        return dict(n_states=n, hamiltonian=H[:n, :n].copy(), eigenvector=c[:n].copy(), proton_log=log[:n].copy(),
                    coupling_matrix=cm[:n].copy(), principal_diabat=pd.value, new_hydronium_mol=nh.value,
                    adiabatic_potential=ad.value)

    def debug_mix_forces(self, coeff):
        coeff = np.ascontiguousarray(coeff, np.float64)
        f = np.zeros((self.system.n_atoms, 3))
        self._check(self.dll.rpb_debug_mix_forces(self.ctx, dptr(coeff), dptr(f)))
        return f

    # -- measurement ------------------------------------------------------------------------
    def launch_counts(self):
        a, b = C.c_longlong(), C.c_longlong()
        self._check(self.dll.rpb_get_launch_counts(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def timers_enable(self, on=True):
        self._check(self.dll.rpb_timers_enable(self.ctx, int(on)))

    def fp64_peak_tflops(self):
        t = C.c_double()
        self._check(self.dll.rpb_measure_fp64_peak(self.ctx, C.byref(t)))
        return t.value

    def timers(self, reset=False):
        n = self.dll.rpb_timer_count()
        ms = np.zeros(max(n, 1)); calls = (C.c_longlong * max(n, 1))()
        self._check(self.dll.rpb_timers_get(self.ctx, dptr(ms), calls))
        out = {self.dll.rpb_timer_name(i).decode(): (float(ms[i]), int(calls[i])) for i in range(n)}
        if reset:
            self._check(self.dll.rpb_timers_reset(self.ctx))
        return out
