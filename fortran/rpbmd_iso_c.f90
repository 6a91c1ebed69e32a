!*******************************************************************************
! rpbmd_iso_c.f90 -- ISO_C_BINDING shim that binds librpbmd.so (include/rpbmd.h) into the
! reference driver jmcdaniel43/Reactive_PB_NN_MD.
!
! Add this file to the reference's src/ (after glob_v.f90 in the link order) and replace the
! bodies of the two force routines by the ones in fortran/patched_force_routines.f90:
!     calculate_total_force_energy          src/total_energy_forces.f90:19-99
!     ms_evb_calculate_total_force_energy   src/ms_evb.f90:181-235
! Nothing else of the reference changes: input formats (.gro/.pmt/.top, simulation parameters),
! the step loop (main_ms_evb.f90:100-119), output cadence and restart logic stay in Fortran.
!
! NOT COMPILED IN THE BUILD IMAGE (no Fortran compiler there).  What IS checked without one:
!   * tests/test_c_abi_driver.py compiles tests/c/abi_driver.c (gcc, no ctypes) which drives the ABI in
!     exactly the call order of rpb_setup / rpb_push_state / rpb_force_energy / rpb_pull_results below, and
!     static_asserts sizeof(rpb_config) and every field offset against the table this file's
!     `type, bind(C) :: rpb_config` implies (4-byte c_int, 8-byte c_double, natural alignment);
!   * tests/test_abi.py: every name bound below is exported by both libraries.
! gfortran notes (the reference is ifort dialect): see fortran/README.md for the patch list.
!*******************************************************************************
module rpbmd_iso_c
  use iso_c_binding
  use global_variables
  implicit none

  integer, parameter :: RPB_MA = 8           ! RPB_MAX_MOLE_ATOMS
  integer, parameter :: RPB_MAXS = 80        ! RPB_EVB_MAX_STATES == evb_max_states (glob_v.f90:60)
  integer, parameter :: RPB_MAXC = 3         ! RPB_EVB_MAX_CHAIN  == evb_max_chain  (glob_v.f90:65)

  type, bind(C) :: rpb_config                 ! include/rpbmd.h  rpb_config, field for field
     integer(c_int) :: n_atoms, n_mole, n_atom_type, n_mole_type
     integer(c_int) :: pme_grid, spline_order, spline_grid, erfc_grid, tt_grid
     integer(c_int) :: na_nslist, nb_nslist, nc_nslist, verlet_capacity
     integer(c_int) :: device, rank, world_size, n_threads, evb_max_chain, evb_max_states
     integer(c_int) :: reserved_i(4)
     real(c_double) :: box(9), alpha_sqrt, real_space_cutoff, verlet_cutoff, delta_t, erfc_dx, tt_max
     real(c_double) :: pi, pi_sqrt, conv_e2A_kJmol, conv_kJmol_ang2ps2gmol, safe_verlet, verlet_thresh
     real(c_double) :: evb_first_solvation_cutoff, evb_reactive_pair_distance, ewald_self, reserved_d(4)
  end type rpb_config

  type, bind(C) :: rpb_energies               ! include/rpbmd.h  rpb_energies
     real(c_double) :: potential_energy, kinetic_energy, E_elec, E_vdw, E_bond, E_angle, E_dihedral, E_recip
  end type rpb_energies

  interface
     function rpb_last_error(ctx) bind(C, name="rpb_last_error")
       import; type(c_ptr), value :: ctx; type(c_ptr) :: rpb_last_error
     end function
     function rpb_backend() bind(C, name="rpb_backend")
       import; type(c_ptr) :: rpb_backend
     end function
     integer(c_int) function rpb_create(ctx, cfg) bind(C, name="rpb_create")
       import; type(c_ptr) :: ctx; type(rpb_config) :: cfg
     end function
     subroutine rpb_destroy(ctx) bind(C, name="rpb_destroy")
       import; type(c_ptr), value :: ctx
     end subroutine
     integer(c_int) function rpb_set_tables(ctx, B6, B5, erfc_t, scale_t, tt, dtt, CB) bind(C, name="rpb_set_tables")
       import; type(c_ptr), value :: ctx; real(c_double) :: B6(*), B5(*), erfc_t(*), scale_t(*), tt(*), dtt(*), CB(*)
     end function
     integer(c_int) function rpb_set_forcefield(ctx, vdw_p, vdw_t, vdw_p14, chg, freeze, bt, bp, at, ap, dt, dp) &
          bind(C, name="rpb_set_forcefield")
       import; type(c_ptr), value :: ctx
       real(c_double) :: vdw_p(*), vdw_p14(*), chg(*), bp(*), ap(*), dp(*)
       integer(c_int) :: vdw_t(*), freeze(*), bt(*), at(*), dt(*)
     end function
     integer(c_int) function rpb_set_molecule_types(ctx, n_atom, atom_type, n_bond, bonds, n_angle, angles, n_dihedral, &
          dihedrals, pair_exclusions, reactive_protons, reactive_basic_atoms) bind(C, name="rpb_set_molecule_types")
       import; type(c_ptr), value :: ctx
       integer(c_int) :: n_atom(*), atom_type(*), n_bond(*), bonds(*), n_angle(*), angles(*), n_dihedral(*), dihedrals(*)
       integer(c_int) :: pair_exclusions(*), reactive_protons(*), reactive_basic_atoms(*)
     end function
     integer(c_int) function rpb_set_evb(ctx, da_i, da_p, pa_i, pa_p, dc_i, dc_p, dc_t, ex_a, ex_p, acid, basic, conj_pairs, &
          conj_atom, ref_e, proton_index, heavy_acid_index) bind(C, name="rpb_set_evb")
       import; type(c_ptr), value :: ctx
       integer(c_int) :: da_i(*), pa_i(*), dc_i(*), dc_t(*), acid(*), basic(*), conj_pairs(*), conj_atom(*), proton_index(*), heavy_acid_index(*)
       real(c_double) :: da_p(*), pa_p(*), dc_p(*), ex_a(*), ex_p(*), ref_e(*)
     end function
     integer(c_int) function rpb_upload_state(ctx, xyz, vel, mass, chg, atype, mfirst, mnatom, mtype, hyd) &
          bind(C, name="rpb_upload_state")
       import; type(c_ptr), value :: ctx; real(c_double) :: xyz(3,*), vel(3,*), mass(*), chg(*)
       integer(c_int) :: atype(*), mfirst(*), mnatom(*), mtype(*); integer(c_int), value :: hyd
     end function
     integer(c_int) function rpb_initialize(ctx) bind(C, name="rpb_initialize")
       import; type(c_ptr), value :: ctx
     end function
     integer(c_int) function rpb_force_energy(ctx, ms_evb) bind(C, name="rpb_force_energy")
       import; type(c_ptr), value :: ctx; integer(c_int), value :: ms_evb
     end function
     integer(c_int) function rpb_step(ctx, n_steps, ms_evb) bind(C, name="rpb_step")
       import; type(c_ptr), value :: ctx; integer(c_int), value :: n_steps, ms_evb
     end function
     integer(c_int) function rpb_get_energies(ctx, e) bind(C, name="rpb_get_energies")
       import; type(c_ptr), value :: ctx; type(rpb_energies) :: e
     end function
     integer(c_int) function rpb_download_state(ctx, xyz, vel, force, mass, chg, atype, mfirst, mnatom, mtype, hyd) &
          bind(C, name="rpb_download_state")
       import; type(c_ptr), value :: ctx; real(c_double) :: xyz(3,*), vel(3,*), force(3,*), mass(*), chg(*)
       integer(c_int) :: atype(*), mfirst(*), mnatom(*), mtype(*), hyd
     end function
     integer(c_int) function rpb_get_r_com(ctx, r_com) bind(C, name="rpb_get_r_com")
       import; type(c_ptr), value :: ctx; real(c_double) :: r_com(3,*)
     end function
     integer(c_int) function rpb_get_neighbor_list(ctx, verlet_point, neighbor_list, capacity, n_pairs, flag) &
          bind(C, name="rpb_get_neighbor_list")
       import; type(c_ptr), value :: ctx; integer(c_int) :: verlet_point(*), neighbor_list(*), n_pairs, flag
       integer(c_int), value :: capacity
     end function
     integer(c_int) function rpb_get_evb(ctx, n_states, hamiltonian, eigenvector, proton_log, coupling_matrix, &
          principal_diabat, new_hydronium_mol, adiabatic_potential) bind(C, name="rpb_get_evb")
       import; type(c_ptr), value :: ctx; integer(c_int) :: n_states, proton_log(*), coupling_matrix(*), principal_diabat, new_hydronium_mol
       real(c_double) :: hamiltonian(*), eigenvector(*), adiabatic_potential
     end function
     ! diabatic-state sharding over several GPUs of one node: the 64-byte handles travel over MPI_Allgather
     integer(c_int) function rpb_peer_export(ctx, handle) bind(C, name="rpb_peer_export")
       import; type(c_ptr), value :: ctx; character(kind=c_char) :: handle(64)
     end function
     integer(c_int) function rpb_peer_import(ctx, handles, world_size) bind(C, name="rpb_peer_import")
       import; type(c_ptr), value :: ctx; character(kind=c_char) :: handles(*); integer(c_int), value :: world_size
     end function
     integer(c_int) function rpb_peer_enabled(ctx) bind(C, name="rpb_peer_enabled")
       import; type(c_ptr), value :: ctx
     end function
  end interface

  type(c_ptr), save :: rpb_handle = c_null_ptr
  ! flat molecule table handed to / returned by the library (molecules are contiguous ascending atom ranges,
  ! general_routines.f90:670-671)
  integer(c_int), allocatable, save :: rpb_mol_first(:), rpb_mol_natom(:), rpb_mol_type(:)
  real(c_double), allocatable, save :: rpb_r_com_store(:,:)      ! centres of mass as the library returns them, (3, n_mole)

contains

  !---------------------------------------------------------------------------
  ! the reference's only error mechanism is `stop "message"` (108 sites)
  !---------------------------------------------------------------------------
  subroutine rpb_check(rc)
    integer(c_int), intent(in) :: rc
    character(kind=c_char), pointer :: msg(:)
    integer :: i
    if (rc /= 0) then
       call c_f_pointer(rpb_last_error(rpb_handle), msg, [512])
       i = 1
       do while (i <= 512)
          if (msg(i) == c_null_char) exit
          write(*,'(A)',advance='no') msg(i)
          i = i + 1
       end do
       write(*,*) ""
       stop "rpbmd: force path failed"
    end if
  end subroutine rpb_check

  !---------------------------------------------------------------------------
  ! one-time set-up.  Call from initialize_energy_force (initialize_routines.f90) after the tables have been built
  ! (:212-264) and periodic_box_change has filled PME_data%CB (pme.f90:634-664), instead of the first force call
  ! at :268-273 -- i.e. as   call rpb_setup(...) ; call (ms_evb_)calculate_total_force_energy(...)
  !---------------------------------------------------------------------------
  subroutine rpb_setup(system_data, molecule_data, atom_data, verlet_list_data, PME_data, integrator_data, real_space_cutoff, device)
    type(system_data_type), intent(in) :: system_data
    type(molecule_data_type), dimension(:), intent(in) :: molecule_data
    type(atom_data_type), intent(in) :: atom_data
    type(verlet_list_data_type), intent(in) :: verlet_list_data
    type(PME_data_type), intent(in) :: PME_data
    type(integrator_data_type), intent(in) :: integrator_data
    real*8, intent(in) :: real_space_cutoff
    integer, intent(in) :: device
    type(rpb_config) :: cfg
    integer :: t, a, b, k, ob, oa, od, nb, na, nd
    integer(c_int), allocatable :: mt_natom(:), mt_atype(:), mt_nbond(:), mt_bonds(:), mt_nangle(:), mt_angles(:), &
         mt_ndih(:), mt_dihs(:), mt_excl(:), mt_rp(:), mt_rb(:)

    cfg%n_atoms = system_data%total_atoms ; cfg%n_mole = system_data%n_mole
    cfg%n_atom_type = n_atom_type ; cfg%n_mole_type = n_molecule_type
    cfg%pme_grid = PME_data%pme_grid ; cfg%spline_order = PME_data%spline_order
    cfg%spline_grid = PME_data%spline_grid ; cfg%erfc_grid = PME_data%erfc_grid ; cfg%tt_grid = Tang_Toennies_grid
    cfg%na_nslist = verlet_list_data%na_nslist ; cfg%nb_nslist = verlet_list_data%nb_nslist ; cfg%nc_nslist = verlet_list_data%nc_nslist
    cfg%verlet_capacity = size(verlet_list_data%neighbor_list)
    cfg%device = device ; cfg%rank = 0 ; cfg%world_size = 1 ; cfg%n_threads = n_threads
    cfg%evb_max_chain = evb_max_chain ; cfg%evb_max_states = evb_max_states
    cfg%reserved_i = 0 ; cfg%reserved_d = 0d0
    cfg%box = reshape(system_data%box, [9])
    cfg%alpha_sqrt = PME_data%alpha_sqrt ; cfg%real_space_cutoff = real_space_cutoff
    cfg%verlet_cutoff = verlet_list_data%verlet_cutoff ; cfg%delta_t = integrator_data%delta_t
    cfg%erfc_dx = PME_data%erfc_dx ; cfg%tt_max = Tang_Toennies_max
    cfg%pi = constants%pi ; cfg%pi_sqrt = constants%pi_sqrt
    cfg%conv_e2A_kJmol = constants%conv_e2A_kJmol ; cfg%conv_kJmol_ang2ps2gmol = constants%conv_kJmol_ang2ps2gmol
    cfg%safe_verlet = verlet_list_data%safe_verlet ; cfg%verlet_thresh = verlet_list_data%verlet_thresh
    cfg%evb_first_solvation_cutoff = evb_first_solvation_cutoff ; cfg%evb_reactive_pair_distance = evb_reactive_pair_distance
    cfg%ewald_self = PME_data%Ewald_self

    call rpb_check( rpb_create(rpb_handle, cfg) )
    call rpb_check( rpb_set_tables(rpb_handle, PME_data%B6_spline, PME_data%B5_spline, PME_data%erfc_table, &
         PME_data%ewaldscale_table, Tang_Toennies_table, dTang_Toennies_table, PME_data%CB) )
    call rpb_check( rpb_set_forcefield(rpb_handle, atype_vdw_parameter, atype_vdw_type, atype_vdw_parameter_14, atype_chg, &
         atype_freeze, atype_bond_type, atype_bond_parameter, atype_angle_type, atype_angle_parameter, &
         atype_dihedral_type, atype_dihedral_parameter) )

    ! molecule_type_data (glob_v.f90:299-317): allocatable components -> flat arrays of include/rpbmd.h
    nb = 0 ; na = 0 ; nd = 0
    do t = 1, n_molecule_type
       nb = nb + size(molecule_type_data(t)%bond_list)
       na = na + size(molecule_type_data(t)%angle_list)
       nd = nd + size(molecule_type_data(t)%dihedral_list)
    end do
    allocate( mt_natom(MAX_N_MOLE_TYPE), mt_atype(MAX_N_MOLE_TYPE*RPB_MA), mt_nbond(MAX_N_MOLE_TYPE), mt_bonds(2*max(nb,1)), &
         mt_nangle(MAX_N_MOLE_TYPE), mt_angles(3*max(na,1)), mt_ndih(MAX_N_MOLE_TYPE), mt_dihs(4*max(nd,1)), &
         mt_excl(MAX_N_MOLE_TYPE*RPB_MA*RPB_MA), mt_rp(MAX_N_MOLE_TYPE*RPB_MA), mt_rb(MAX_N_MOLE_TYPE*RPB_MA) )
    mt_natom = 0 ; mt_atype = 0 ; mt_nbond = 0 ; mt_nangle = 0 ; mt_ndih = 0 ; mt_excl = 0 ; mt_rp = 0 ; mt_rb = 0
    ob = 0 ; oa = 0 ; od = 0
    do t = 1, n_molecule_type
       mt_natom(t) = molecule_type_data(t)%n_atom
       if ( mt_natom(t) + 1 > RPB_MA ) stop "rpbmd: molecule type larger than RPB_MAX_MOLE_ATOMS-1"
       do a = 1, mt_natom(t)
          mt_atype((t-1)*RPB_MA + a) = molecule_type_data(t)%atom_type_index(a)
          if ( allocated(molecule_type_data(t)%evb_reactive_protons) ) mt_rp((t-1)*RPB_MA + a) = molecule_type_data(t)%evb_reactive_protons(a)
          if ( allocated(molecule_type_data(t)%evb_reactive_basic_atoms) ) mt_rb((t-1)*RPB_MA + a) = molecule_type_data(t)%evb_reactive_basic_atoms(a)
          do b = 1, mt_natom(t)
             mt_excl((t-1)*RPB_MA*RPB_MA + a + RPB_MA*(b-1)) = molecule_type_data(t)%pair_exclusions(a,b)
          end do
       end do
       mt_nbond(t) = size(molecule_type_data(t)%bond_list)
       do k = 1, mt_nbond(t)
          mt_bonds(2*(ob+k)-1) = molecule_type_data(t)%bond_list(k)%i_atom ; mt_bonds(2*(ob+k)) = molecule_type_data(t)%bond_list(k)%j_atom
       end do
       ob = ob + mt_nbond(t)
       mt_nangle(t) = size(molecule_type_data(t)%angle_list)
       do k = 1, mt_nangle(t)
          mt_angles(3*(oa+k)-2) = molecule_type_data(t)%angle_list(k)%i_atom ; mt_angles(3*(oa+k)-1) = molecule_type_data(t)%angle_list(k)%j_atom
          mt_angles(3*(oa+k)) = molecule_type_data(t)%angle_list(k)%k_atom
       end do
       oa = oa + mt_nangle(t)
       mt_ndih(t) = size(molecule_type_data(t)%dihedral_list)
       do k = 1, mt_ndih(t)
          mt_dihs(4*(od+k)-3) = molecule_type_data(t)%dihedral_list(k)%i_atom ; mt_dihs(4*(od+k)-2) = molecule_type_data(t)%dihedral_list(k)%j_atom
          mt_dihs(4*(od+k)-1) = molecule_type_data(t)%dihedral_list(k)%k_atom ; mt_dihs(4*(od+k)) = molecule_type_data(t)%dihedral_list(k)%l_atom
       end do
       od = od + mt_ndih(t)
    end do
    call rpb_check( rpb_set_molecule_types(rpb_handle, mt_natom, mt_atype, mt_nbond, mt_bonds, mt_nangle, mt_angles, mt_ndih, &
         mt_dihs, mt_excl, mt_rp, mt_rb) )

    if ( ms_evb_simulation == "yes" ) then
       call rpb_check( rpb_set_evb(rpb_handle, evb_donor_acceptor_interaction, evb_donor_acceptor_parameters, &
            evb_proton_acceptor_interaction, evb_proton_acceptor_parameters, evb_diabat_coupling_interaction, &
            evb_diabat_coupling_parameters, evb_diabat_coupling_type, evb_exchange_charge_atomic, evb_exchange_charge_proton, &
            evb_acid_molecule, evb_basic_molecule, evb_conjugate_pairs, evb_conjugate_atom_index, evb_reference_energy, &
            evb_proton_index, evb_heavy_acid_index) )
    end if

    allocate( rpb_mol_first(system_data%n_mole), rpb_mol_natom(system_data%n_mole), rpb_mol_type(system_data%n_mole) )
    call rpb_push_state(system_data, molecule_data, atom_data)
    call rpb_check( rpb_initialize(rpb_handle) )     ! update_r_com, shift_molecules_into_box, construct_verlet_list (initialize_routines.f90:121-134)
  end subroutine rpb_setup

  !---------------------------------------------------------------------------
  ! atom_data / molecule_data -> library.  The (3,N) / (N) component arrays are contiguous: the column-major layout IS
  ! the ABI's layout, so they are passed as they are.
  !---------------------------------------------------------------------------
  subroutine rpb_push_state(system_data, molecule_data, atom_data)
    type(system_data_type), intent(in) :: system_data
    type(molecule_data_type), dimension(:), intent(in) :: molecule_data
    type(atom_data_type), intent(in) :: atom_data
    integer :: i, hyd
    do i = 1, system_data%n_mole
       rpb_mol_first(i) = molecule_data(i)%atom_index(1)
       rpb_mol_natom(i) = molecule_data(i)%n_atom
       rpb_mol_type(i)  = molecule_data(i)%molecule_type_index
    end do
    hyd = 0
    if ( ms_evb_simulation == "yes" .and. n_hydronium_molecules > 0 ) hyd = hydronium_molecule_index(1)
    call rpb_check( rpb_upload_state(rpb_handle, atom_data%xyz, atom_data%velocity, atom_data%mass, atom_data%charge, &
         atom_data%atom_type_index, rpb_mol_first, rpb_mol_natom, rpb_mol_type, int(hyd, c_int)) )
  end subroutine rpb_push_state

  !---------------------------------------------------------------------------
  ! library -> atom_data%force, system_data energies, and (after a committed proton hop) the permuted atom arrays, the
  ! molecule table and hydronium_molecule_index(1), exactly as evb_change_diabat_data_structure_topology leaves them
  ! (ms_evb.f90:806-932).  Atom NAMES follow the types (atype_name), as in evb_change_data_structures_proton_transfer.
  !---------------------------------------------------------------------------
  subroutine rpb_pull_results(system_data, molecule_data, atom_data, PME_data)
    type(system_data_type), intent(inout) :: system_data
    type(molecule_data_type), dimension(:), intent(inout) :: molecule_data
    type(atom_data_type), intent(inout) :: atom_data
    type(PME_data_type), intent(inout) :: PME_data
    type(rpb_energies) :: e
    integer(c_int) :: hyd
    integer :: i, a
    call rpb_check( rpb_download_state(rpb_handle, atom_data%xyz, atom_data%velocity, atom_data%force, atom_data%mass, &
         atom_data%charge, atom_data%atom_type_index, rpb_mol_first, rpb_mol_natom, rpb_mol_type, hyd) )
    call rpb_check( rpb_get_energies(rpb_handle, e) )
    system_data%potential_energy = e%potential_energy
    system_data%E_elec = e%E_elec ; system_data%E_vdw = e%E_vdw
    system_data%E_bond = e%E_bond ; system_data%E_angle = e%E_angle ; system_data%E_dihedral = e%E_dihedral
    PME_data%E_recip = e%E_recip
    do i = 1, system_data%n_mole
       if ( molecule_data(i)%n_atom /= rpb_mol_natom(i) .or. molecule_data(i)%atom_index(1) /= rpb_mol_first(i) .or. &
            molecule_data(i)%molecule_type_index /= rpb_mol_type(i) ) then
          molecule_data(i)%n_atom = rpb_mol_natom(i)
          molecule_data(i)%molecule_type_index = rpb_mol_type(i)
          molecule_data(i)%mname = molecule_type_data(rpb_mol_type(i))%mname
          do a = 1, rpb_mol_natom(i)                       ! atom_index is allocated n_atom+1 (general_routines.f90:246-248)
             molecule_data(i)%atom_index(a) = rpb_mol_first(i) + a - 1
             atom_data%aname(rpb_mol_first(i) + a - 1) = atype_name( atom_data%atom_type_index(rpb_mol_first(i) + a - 1) )
          end do
       end if
    end do
    if ( ms_evb_simulation == "yes" .and. hyd > 0 ) hydronium_molecule_index(1) = hyd
    if ( .not. allocated(rpb_r_com_store) ) allocate( rpb_r_com_store(3, system_data%n_mole) )
    call rpb_check( rpb_get_r_com(rpb_handle, rpb_r_com_store) )      ! (the array itself: a function result would be copied)
    do i = 1, system_data%n_mole
       molecule_data(i)%r_com(:) = rpb_r_com_store(:, i)
    end do
  end subroutine rpb_pull_results

end module rpbmd_iso_c
