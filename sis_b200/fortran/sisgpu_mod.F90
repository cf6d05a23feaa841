! sisgpu_mod.F90 -- ISO_C_BINDING interface of libsisgpu.so (include/sisgpu.h) for the SIS
! Fortran driver.  Add this file to src/CMakeLists.txt, compile with -DSIS_GPU and link with
! -lsisgpu; INTEGRATION.md lists the call sites that are redirected.
!
! The module holds one opaque handle per process (the reference's state is module-global
! too) and thin wrappers with the argument lists of the reference routines they replace:
!   gpu_neighbor_list(irep)                     <- neighbor_list(irep)        src/neighbor_list.F90:20
!   gpu_force(irep, forces)                     <- force(irep, forces)        src/force.F90:14-15
!   gpu_energy_sumup(irep, tempK, energies)     <- energy_sumup(...)          src/energy_sumup.F90:11-13
!   gpu_energy_kinetic(irep, Ekin)              <- energy_kinetic             src/energy_kinetic.F90:17-29
!   gpu_md_steps(nsteps)                        <- loop body                  src/job_md.F90:374-460
!   gpu_energy_replica / gpu_replica_exchange   <- src/energy_replica.F90:13-16, src/replica_exchange.F90:19-21
! A non-zero status prints sisgpu_last_error() and calls sis_abort (src/sis_abort.F90:12-19).
!
! No Fortran compiler exists in the build image of this repository, so this file is checked by
! tests/test_abi.py only for consistency with the header (every bind(C) name must be exported
! with the same argument count); it is deliberately mechanical.
module sisgpu_mod

   use, intrinsic :: iso_c_binding
   implicit none
   private

   type(c_ptr), save :: h = c_null_ptr

   public :: gpu_init, gpu_finalize, gpu_upload_state, gpu_download_state
   public :: gpu_neighbor_list, gpu_full_lists, gpu_force, gpu_energy_sumup, gpu_energy_kinetic
   public :: gpu_md_steps, gpu_energy_replica, gpu_replica_exchange, gpu_get_bp_state
   public :: gpu_set_replica_params, gpu_set_box, gpu_check

   interface
      function sisgpu_last_error() bind(C, name='sisgpu_last_error') result(msg)
         import :: c_ptr
         type(c_ptr) :: msg
      end function
      function sisgpu_create(out, device) bind(C, name='sisgpu_create') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), intent(out) :: out
         integer(c_int), value :: device
         integer(c_int) :: rc
      end function
      function sisgpu_destroy(h) bind(C, name='sisgpu_destroy') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int) :: rc
      end function
      function sisgpu_set_topology(h, nmp, nchains, nmp_chain, moltypes, seq, has_charge, is_frozen, mass) &
            bind(C, name='sisgpu_set_topology') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: nmp, nchains
         integer(c_int), intent(in) :: nmp_chain(*), moltypes(*), seq(*), has_charge(*), is_frozen(*)
         real(c_double), intent(in) :: mass(*)
         integer(c_int) :: rc
      end function
      function sisgpu_set_replicas(h, nrep_proc) bind(C, name='sisgpu_set_replicas') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int), value :: nrep_proc
         integer(c_int) :: rc
      end function
      function sisgpu_set_ff_bonded(h, bond_k, bond_r0, angl_type, angl_k, angl_t0, dih_flags, dih_k, dih_w, dih_p0) &
            bind(C, name='sisgpu_set_ff_bonded') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         real(c_double), value :: bond_k, bond_r0, angl_k, angl_t0, dih_k, dih_w, dih_p0
         integer(c_int), value :: angl_type, dih_flags
         integer(c_int) :: rc
      end function
      function sisgpu_set_ff_wca(h, sigma, eps) bind(C, name='sisgpu_set_ff_wca') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         real(c_double), value :: sigma, eps
         integer(c_int) :: rc
      end function
      function sisgpu_set_ff_bp(h, model, min_loop, max_bp_per_nt, nstep_MC, cutoff_energy, bp_paras) &
            bind(C, name='sisgpu_set_ff_bp') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: model, min_loop, max_bp_per_nt, nstep_MC
         real(c_double), value :: cutoff_energy
         real(c_double), intent(in) :: bp_paras(15, 3)
         integer(c_int) :: rc
      end function
      function sisgpu_set_ff_nn(h, NN_dH, NN_dS, dH0, dS0, flg_NNend, NNend_dH, NNend_dS, dHend0, dSend0) &
            bind(C, name='sisgpu_set_ff_nn') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         real(c_double), intent(in) :: NN_dH(21), NN_dS(21), NNend_dH(6), NNend_dS(6)
         real(c_double), value :: dH0, dS0, dHend0, dSend0
         integer(c_int), value :: flg_NNend
         integer(c_int) :: rc
      end function
      function sisgpu_set_ele(h, flg_ele, cutoff_type, cutoff_inp, excl_bond, excl_angle, length_per_charge, &
                              temp_independent, tempK_ref) bind(C, name='sisgpu_set_ele') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: flg_ele, cutoff_type, excl_bond, excl_angle, temp_independent
         real(c_double), value :: cutoff_inp, length_per_charge, tempK_ref
         integer(c_int) :: rc
      end function
      function sisgpu_set_pbc(h, flg_pbc, box) bind(C, name='sisgpu_set_pbc') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: flg_pbc
         real(c_double), intent(in) :: box(3)
         integer(c_int) :: rc
      end function
      function sisgpu_set_md(h, integrator, dt, viscosity_Pas, nl_margin, seed) &
            bind(C, name='sisgpu_set_md') result(rc)
         import :: c_ptr, c_int, c_double, c_int64_t
         type(c_ptr), value :: h
         integer(c_int), value :: integrator
         real(c_double), value :: dt, viscosity_Pas, nl_margin
         integer(c_int64_t), value :: seed
         integer(c_int) :: rc
      end function
      function sisgpu_set_bias_ss(h, flg, force) bind(C, name='sisgpu_set_bias_ss') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: flg
         real(c_double), value :: force
         integer(c_int) :: rc
      end function
      function sisgpu_set_stage(h, flg, sigma, eps) bind(C, name='sisgpu_set_stage') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: flg
         real(c_double), value :: sigma, eps
         integer(c_int) :: rc
      end function
      function sisgpu_set_tweezers_dcf(h, n, pairs, direction) bind(C, name='sisgpu_set_tweezers_dcf') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: n
         integer(c_int), intent(in) :: pairs(2, *)
         real(c_double), intent(in) :: direction(3, *)
         integer(c_int) :: rc
      end function
      function sisgpu_set_tweezers_fr(h, n, pairs, k, init, velo) bind(C, name='sisgpu_set_tweezers_fr') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: n
         integer(c_int), intent(in) :: pairs(2, *)
         real(c_double), intent(in) :: k(2, *), init(3, 2, *), velo(3, 2, *)
         integer(c_int) :: rc
      end function
      function sisgpu_set_bias_rg(h, flg, pott, k) bind(C, name='sisgpu_set_bias_rg') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: flg, pott
         real(c_double), value :: k
         integer(c_int) :: rc
      end function
      function sisgpu_set_restraints(h, n, id, rcut, para) bind(C, name='sisgpu_set_restraints') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: n
         integer(c_int), intent(in) :: id(3, *)
         real(c_double), intent(in) :: rcut(*), para(3, *)
         integer(c_int) :: rc
      end function
      function sisgpu_set_replica_params(h, irep, tempK, ionic_strength, twz_force, rg0) &
            bind(C, name='sisgpu_set_replica_params') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: irep
         real(c_double), value :: tempK, ionic_strength, twz_force, rg0
         integer(c_int) :: rc
      end function
      function sisgpu_finalize_setup(h) bind(C, name='sisgpu_finalize_setup') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int) :: rc
      end function
      function sisgpu_set_state(h, irep, xyz, velos, accels) bind(C, name='sisgpu_set_state') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h, xyz, velos, accels     ! c_loc(array) or c_null_ptr to skip
         integer(c_int), value :: irep
         integer(c_int) :: rc
      end function
      function sisgpu_get_state(h, irep, xyz, velos, accels) bind(C, name='sisgpu_get_state') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h, xyz, velos, accels
         integer(c_int), value :: irep
         integer(c_int) :: rc
      end function
      function sisgpu_set_step(h, istep) bind(C, name='sisgpu_set_step') result(rc)
         import :: c_ptr, c_int, c_int64_t
         type(c_ptr), value :: h
         integer(c_int64_t), value :: istep
         integer(c_int) :: rc
      end function
      function sisgpu_neighbor_list(h, irep) bind(C, name='sisgpu_neighbor_list') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int), value :: irep
         integer(c_int) :: rc
      end function
      function sisgpu_full_lists(h) bind(C, name='sisgpu_full_lists') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int) :: rc
      end function
      function sisgpu_get_list_sizes(h, irep, nwca, nele, nbp) bind(C, name='sisgpu_get_list_sizes') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int), value :: irep
         integer(c_int), intent(out) :: nwca, nele, nbp
         integer(c_int) :: rc
      end function
      function sisgpu_get_lists(h, irep, wca_mp, ele_mp, bp_mp, bp_coef) bind(C, name='sisgpu_get_lists') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h, wca_mp, ele_mp, bp_mp, bp_coef
         integer(c_int), value :: irep
         integer(c_int) :: rc
      end function
      function sisgpu_set_bp_mc(h, flg_bp_MC) bind(C, name='sisgpu_set_bp_mc') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int), value :: flg_bp_MC
         integer(c_int) :: rc
      end function
      function sisgpu_force(h, irep, forces) bind(C, name='sisgpu_force') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: irep
         real(c_double), intent(out) :: forces(3, *)
         integer(c_int) :: rc
      end function
      function sisgpu_energy_sumup(h, irep, tempK, recompute_bp, energies) &
            bind(C, name='sisgpu_energy_sumup') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: irep, recompute_bp
         real(c_double), value :: tempK
         real(c_double), intent(out) :: energies(0:10)
         integer(c_int) :: rc
      end function
      function sisgpu_energy_kinetic(h, irep, ekin) bind(C, name='sisgpu_energy_kinetic') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: irep
         real(c_double), intent(out) :: ekin
         integer(c_int) :: rc
      end function
      function sisgpu_set_bp_restrict(h, n, pairs) bind(C, name='sisgpu_set_bp_restrict') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int), value :: n
         integer(c_int), intent(in) :: pairs(2, *)
         integer(c_int) :: rc
      end function
      function sisgpu_snapshot_begin(h, irep, tempK, flags, slot) bind(C, name='sisgpu_snapshot_begin') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: irep, flags
         real(c_double), value :: tempK
         integer(c_int), intent(out) :: slot
         integer(c_int) :: rc
      end function
      function sisgpu_snapshot_end(h, slot, xyz32, nbp, bp_i, bp_j, bp_e, cap, energies, ekin) &
            bind(C, name='sisgpu_snapshot_end') result(rc)
         import :: c_ptr, c_int, c_double, c_float
         type(c_ptr), value :: h
         integer(c_int), value :: slot, cap
         real(c_float), intent(out) :: xyz32(*), bp_e(*)
         integer(c_int), intent(out) :: nbp, bp_i(*), bp_j(*)
         real(c_double), intent(out) :: energies(*), ekin
         integer(c_int) :: rc
      end function
      function sisgpu_dcd_energies(h, nframes, xyz32, tempK, energies) bind(C, name='sisgpu_dcd_energies') result(rc)
         import :: c_ptr, c_int, c_double, c_float
         type(c_ptr), value :: h
         integer(c_int), value :: nframes
         real(c_float), intent(in) :: xyz32(*)
         real(c_double), value :: tempK
         real(c_double), intent(out) :: energies(*)
         integer(c_int) :: rc
      end function
      function sisgpu_get_bp_state(h, irep, status, ene) bind(C, name='sisgpu_get_bp_state') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: irep
         integer(c_int), intent(out) :: status(*)
         real(c_double), intent(out) :: ene(*)
         integer(c_int) :: rc
      end function
      function sisgpu_init_velocities(h, irep) bind(C, name='sisgpu_init_velocities') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int), value :: irep
         integer(c_int) :: rc
      end function
      function sisgpu_md_steps(h, nsteps, nrebuild) bind(C, name='sisgpu_md_steps') result(rc)
         import :: c_ptr, c_int, c_int64_t
         type(c_ptr), value :: h
         integer(c_int64_t), value :: nsteps
         integer(c_int64_t), intent(out) :: nrebuild
         integer(c_int) :: rc
      end function
      function sisgpu_check_force(h, tempK, f_energy, f_force, maxdiff) bind(C, name='sisgpu_check_force') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         real(c_double), value :: tempK
         real(c_double), intent(out) :: f_energy(3, *), f_force(3, *), maxdiff
         integer(c_int) :: rc
      end function
      function sisgpu_remd_setup(h, nrep, values, nstep_exchange, rank, nranks) &
            bind(C, name='sisgpu_remd_setup') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), intent(in) :: nrep(4)
         real(c_double), intent(in) :: values(256, 4)
         integer(c_int), value :: nstep_exchange, rank, nranks
         integer(c_int) :: rc
      end function
      function sisgpu_nccl_unique_id(id) bind(C, name='sisgpu_nccl_unique_id') result(rc)
         import :: c_int, c_signed_char
         integer(c_signed_char), intent(out) :: id(128)
         integer(c_int) :: rc
      end function
      function sisgpu_comm_init(h, nranks, rank, id) bind(C, name='sisgpu_comm_init') result(rc)
         import :: c_ptr, c_int, c_signed_char
         type(c_ptr), value :: h
         integer(c_int), value :: nranks, rank
         integer(c_signed_char), intent(in) :: id(128)
         integer(c_int) :: rc
      end function
      function sisgpu_energy_replica(h, energies, replica_energies) bind(C, name='sisgpu_energy_replica') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         real(c_double), intent(out) :: energies(0:10, *), replica_energies(2, *)
         integer(c_int) :: rc
      end function
      function sisgpu_replica_exchange(h, replica_energies, uniforms, n_attempt, n_accept) &
            bind(C, name='sisgpu_replica_exchange') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h, uniforms             ! c_null_ptr: library stream
         real(c_double), intent(in) :: replica_energies(2, *)
         integer(c_int), intent(out) :: n_attempt, n_accept
         integer(c_int) :: rc
      end function
      function sisgpu_remd_state(h, rep2lab, lab2rep, lab2val) bind(C, name='sisgpu_remd_state') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h, rep2lab, lab2rep, lab2val
         integer(c_int) :: rc
      end function
      function sisgpu_get_rng_state(h, state, cap, n) bind(C, name='sisgpu_get_rng_state') result(rc)
         import :: c_ptr, c_int, c_int64_t
         type(c_ptr), value :: h
         integer(c_int64_t), intent(out) :: state(*)
         integer(c_int), value :: cap
         integer(c_int), intent(out) :: n
         integer(c_int) :: rc
      end function
      function sisgpu_set_rng_state(h, state, n) bind(C, name='sisgpu_set_rng_state') result(rc)
         import :: c_ptr, c_int, c_int64_t
         type(c_ptr), value :: h
         integer(c_int64_t), intent(in) :: state(*)
         integer(c_int), value :: n
         integer(c_int) :: rc
      end function
      function sisgpu_remd_set_labels(h, rep2lab, idx_pair) bind(C, name='sisgpu_remd_set_labels') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int), intent(in) :: rep2lab(*)
         integer(c_int), value :: idx_pair
         integer(c_int) :: rc
      end function
      function sisgpu_peer_export(h, handle) bind(C, name='sisgpu_peer_export') result(rc)
         import :: c_ptr, c_int, c_char
         type(c_ptr), value :: h
         character(kind=c_char), intent(out) :: handle(*)
         integer(c_int) :: rc
      end function
      function sisgpu_peer_import(h, nranks, handles) bind(C, name='sisgpu_peer_import') result(rc)
         import :: c_ptr, c_int, c_char
         type(c_ptr), value :: h
         integer(c_int), value :: nranks
         character(kind=c_char), intent(in) :: handles(*)
         integer(c_int) :: rc
      end function
      function sisgpu_bp_status_to_mc(h) bind(C, name='sisgpu_bp_status_to_mc') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int) :: rc
      end function
      function sisgpu_bpcoef(h, cap, n, imp, jmp, dH, dS) bind(C, name='sisgpu_bpcoef') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: cap
         integer(c_int), intent(out) :: n
         integer(c_int), intent(out) :: imp(*), jmp(*)
         real(c_double), intent(out) :: dH(*), dS(*)
         integer(c_int) :: rc
      end function
   end interface

contains

   ! status -> sis_abort, the reference's error convention (src/sis_abort.F90:12-19)
   subroutine gpu_check(rc, where)
      integer(c_int), intent(in) :: rc
      character(len=*), intent(in) :: where
      character(kind=c_char), pointer :: msg(:)
      integer :: i
      if (rc == 0) return
      call c_f_pointer(sisgpu_last_error(), msg, [1024])
      write(*, '(3a)', advance='no') 'Error in libsisgpu (', where, '): '
      do i = 1, 1024
         if (msg(i) == c_null_char) exit
         write(*, '(a)', advance='no') msg(i)
      enddo
      write(*, *)
      call sis_abort()
   end subroutine

   ! Called from main.F90 after list_local (src/main.F90:140): hands the outputs of the
   ! reference's read_* / init_* routines to the device.  Module globals are read directly.
   subroutine gpu_init(myrank, nprocs)
      use const, only : PREC
      use const_idx, only : MOLT, BPT, INTGRT, REPT
      use pbc, only : flg_pbc, pbc_box
      use var_top, only : nmp, nchains, nmp_chain, moltypes, seq, has_charge, is_frozen, mass, imp_chain
      use var_state, only : viscosity_Pas, tempK, ionic_strength, dt, nl_margin, integrator, rng_seed, &
                            nstep_bp_MC, flg_bp_MC, bp_status
      use var_potential
      use var_replica, only : nrep_proc, nrep, flg_replica, nstep_rep_exchange, replica_values, &
                              irep2grep, rep2lab, lab2val, flg_repvar
      integer, intent(in) :: myrank, nprocs
      integer(c_int) :: ndev_rank, c, l, imp, irep, dihf, anglt
      integer(c_int), allocatable :: seq_mp(:), chg(:), frz(:)
      real(c_double) :: bpp(15, 3)
      integer(c_signed_char) :: nid(128)
      integer :: t, ierr

      ndev_rank = myrank      ! sisgpu_create takes rank mod #devices itself when out of range
      call gpu_check(sisgpu_create(h, int(mod(ndev_rank, 8), c_int)), 'create')

      allocate(seq_mp(nmp), chg(nmp), frz(nmp))
      do c = 1, nchains
         do l = 1, nmp_chain(c)
            imp = imp_chain(l, c)
            seq_mp(imp) = seq(l, c)
         enddo
      enddo
      chg = merge(1_c_int, 0_c_int, has_charge(1:nmp))
      frz = 0
      if (allocated(is_frozen)) frz = merge(1_c_int, 0_c_int, is_frozen(1:nmp))
      call gpu_check(sisgpu_set_topology(h, nmp, nchains, nmp_chain, moltypes, seq_mp, chg, frz, mass), 'topology')
      call gpu_check(sisgpu_set_replicas(h, nrep_proc), 'replicas')

      anglt = merge(1, 0, flg_angl_ReB)
      dihf = merge(1, 0, flg_dih_cos) + 2 * merge(1, 0, flg_dih_exp)
      call gpu_check(sisgpu_set_ff_bonded(h, bond_k, bond_r0, anglt, angl_k, angl_t0, dihf, dih_k, dih_w, dih_p0), 'bonded')
      call gpu_check(sisgpu_set_ff_wca(h, wca_sigma, wca_eps), 'wca')
      do t = 1, 3   ! GC, AU, GU; the CG/UA/UG sets are re-derived by the library
         bpp(:, t) = [bp_paras(t)%U0, bp_paras(t)%bond_k, bp_paras(t)%bond_r, &
                      bp_paras(t)%angl_k1, bp_paras(t)%angl_k2, bp_paras(t)%angl_k3, bp_paras(t)%angl_k4, &
                      bp_paras(t)%angl_theta1, bp_paras(t)%angl_theta2, bp_paras(t)%angl_theta3, bp_paras(t)%angl_theta4, &
                      bp_paras(t)%dihd_k1, bp_paras(t)%dihd_k2, bp_paras(t)%dihd_phi1, bp_paras(t)%dihd_phi2]
      enddo
      call gpu_check(sisgpu_set_ff_bp(h, bp_model, bp_min_loop, max_bp_per_nt, nstep_bp_MC, bp_cutoff_energy, bpp), 'bp')
      if (bp_model == 5) then
         call gpu_check(sisgpu_set_ff_nn(h, NN_dH(1:21), NN_dS(1:21), dH0, dS0, merge(1, 0, flg_NNend), &
                                         NNend_dH(1:6), NNend_dS(1:6), dHend0, dSend0), 'nn')
      endif
      call gpu_check(sisgpu_set_ele(h, merge(1, 0, flg_ele), ele_cutoff_type, ele_cutoff_inp, &
                                    merge(1, 0, ele_exclude_covalent_bond_pairs), merge(1, 0, ele_exclude_bond_angle_pairs), &
                                    length_per_charge, merge(1, 0, flg_ele_temp_independent), tempK_ele_ref), 'ele')
      call gpu_check(sisgpu_set_pbc(h, merge(1, 0, flg_pbc), pbc_box), 'pbc')
      call gpu_check(sisgpu_set_md(h, integrator, dt, viscosity_Pas, nl_margin, int(rng_seed, c_int64_t)), 'md')
      ! biases: sisgpu_set_stage / _tweezers_dcf / _tweezers_fr / _bias_rg / _restraints / _bias_ss
      ! are called here from var_potential / var_state in the same way (INTEGRATION.md table 2).
      do irep = 1, nrep_proc
         call gpu_set_replica_params(irep)
      enddo
      call gpu_check(sisgpu_finalize_setup(h), 'finalize_setup')

      if (flg_replica) then
         call gpu_check(sisgpu_remd_setup(h, nrep, replica_values, nstep_rep_exchange, myrank, nprocs), 'remd_setup')
#ifdef PAR_MPI
         if (nprocs > 1) then    ! the one hot-loop collective: NCCL all-gather instead of MPI_ALLREDUCE
            if (myrank == 0) call gpu_check(sisgpu_nccl_unique_id(nid), 'nccl id')
            call MPI_BCAST(nid, 128, MPI_BYTE, 0, MPI_COMM_WORLD, ierr)
            call gpu_check(sisgpu_comm_init(h, nprocs, myrank, nid), 'comm_init')
         endif
#endif
      endif
   end subroutine

   subroutine gpu_finalize()
      integer(c_int) :: rc
      rc = sisgpu_destroy(h)
      h = c_null_ptr
   end subroutine

   ! set_md_coef + set_ele for one local replica (src/job_md.F90:538-567)
   subroutine gpu_set_replica_params(irep)
      use var_state, only : tempK, ionic_strength
      use var_replica, only : flg_repvar, rep2val, irep2grep
      use const_idx, only : REPT
      integer, intent(in) :: irep
      real(c_double) :: T, ion, f, rg0
      T = tempK; ion = ionic_strength; f = 0.0_c_double; rg0 = 0.0_c_double
      if (flg_repvar(REPT%TEMP))   T   = rep2val(irep2grep(irep), REPT%TEMP)
      if (flg_repvar(REPT%ION))    ion = rep2val(irep2grep(irep), REPT%ION)
      if (flg_repvar(REPT%TWZDCF)) f   = rep2val(irep2grep(irep), REPT%TWZDCF)
      if (flg_repvar(REPT%RG))     rg0 = rep2val(irep2grep(irep), REPT%RG)
      call gpu_check(sisgpu_set_replica_params(h, int(irep - 1, c_int), T, ion, f, rg0), 'set_replica_params')
   end subroutine

   subroutine gpu_set_box(box)      ! pbc_resize, src/job_md.F90:576-594
      real(c_double), intent(in) :: box(3)
      call gpu_check(sisgpu_set_pbc(h, 1_c_int, box), 'set_pbc')
   end subroutine

   subroutine gpu_upload_state(irep)
      use var_state, only : xyz, velos, accels
      integer, intent(in) :: irep
      call gpu_check(sisgpu_set_state(h, int(irep - 1, c_int), c_loc(xyz(1, 1, irep)), c_loc(velos(1, 1, irep)), &
                                      c_loc(accels(1, 1, irep))), 'set_state')
   end subroutine

   subroutine gpu_download_state(irep)
      use var_state, only : xyz, velos, accels
      integer, intent(in) :: irep
      call gpu_check(sisgpu_get_state(h, int(irep - 1, c_int), c_loc(xyz(1, 1, irep)), c_loc(velos(1, 1, irep)), &
                                      c_loc(accels(1, 1, irep))), 'get_state')
   end subroutine

   ! PRNG / PRNGREP blocks of the .rst file (src/write_rst.F90:69-131, src/read_rst.F90:405-600): the
   ! counter-based device generators have no Mersenne-Twister state; their restart payload is
   ! (seed, istep, exchange draw counter, nrep_proc, Monte-Carlo draw counter per local replica),
   ! written under a new tag RSTBLK%PRNGCTR = 12 next to the existing ones (src/const_idx.F90:73-85)
   subroutine gpu_write_rst_prng(lunout)
      use var_replica, only : nrep_proc
      integer, intent(in) :: lunout
      integer(c_int64_t) :: st(4 + nrep_proc)
      integer(c_int) :: n
      call gpu_check(sisgpu_get_rng_state(h, st, int(size(st), c_int), n), 'get_rng_state')
      write(lunout) 12                                   ! RSTBLK%PRNGCTR
      write(lunout) int(8 * n, kind=4)                   ! nbytes
      write(lunout) st(1:n)
   end subroutine

   subroutine gpu_read_rst_prng(lunin, nbytes)
      integer, intent(in) :: lunin, nbytes
      integer(c_int64_t) :: st(nbytes / 8)
      read(lunin) st
      call gpu_check(sisgpu_set_rng_state(h, st, int(size(st), c_int)), 'set_rng_state')
   end subroutine

   ! REPLICA block (src/read_rst.F90:405-460): labels read from the .rst go back to the device tables
   subroutine gpu_remd_set_labels()
      use var_replica, only : rep2lab, nrep_all
      integer(c_int) :: r2l(nrep_all)
      r2l(:) = int(rep2lab(1:nrep_all), c_int)
      call gpu_check(sisgpu_remd_set_labels(h, r2l, int(0, c_int)), 'remd_set_labels')
   end subroutine

   ! exchange buffers of the ranks of one node mapped into each other (CUDA IPC): the MPI_ALLREDUCE of
   ! src/job_md.F90:531-533 becomes peer stores over NVLink inside gpu_md_steps
   subroutine gpu_peer_init(myrank, nprocs)
#ifdef PAR_MPI
      use mpi
#endif
      integer, intent(in) :: myrank, nprocs
      character(kind=c_char) :: mine(64), all_handles(64 * nprocs)
      integer :: ierr
      call gpu_check(sisgpu_peer_export(h, mine), 'peer_export')
#ifdef PAR_MPI
      call MPI_ALLGATHER(mine, 64, MPI_CHARACTER, all_handles, 64, MPI_CHARACTER, MPI_COMM_WORLD, ierr)
#else
      all_handles(1:64) = mine
#endif
      if (nprocs > 1) call gpu_check(sisgpu_peer_import(h, int(nprocs, c_int), all_handles), 'peer_import')
   end subroutine

   subroutine gpu_neighbor_list(irep)
      integer, intent(in) :: irep
      call gpu_check(sisgpu_neighbor_list(h, int(irep - 1, c_int)), 'neighbor_list')
   end subroutine

   subroutine gpu_full_lists()      ! list_bp / list_exv / list_ele, src/main.F90:150-152
      call gpu_check(sisgpu_full_lists(h), 'full_lists')
   end subroutine

   subroutine gpu_force(irep, forces)
      use var_top, only : nmp
      use var_state, only : flg_bp_MC
      integer, intent(in) :: irep
      real(c_double), intent(out) :: forces(3, nmp)
      call gpu_check(sisgpu_set_bp_mc(h, merge(1, 0, flg_bp_MC)), 'set_bp_mc')
      call gpu_check(sisgpu_force(h, int(irep - 1, c_int), forces), 'force')
   end subroutine

   subroutine gpu_energy_sumup(irep, tempK_in, energies)
      use var_state, only : flg_bp_energy, flg_bp_MC
      integer, intent(in) :: irep
      real(c_double), intent(in) :: tempK_in
      real(c_double), intent(out) :: energies(0:10)
      call gpu_check(sisgpu_set_bp_mc(h, merge(1, 0, flg_bp_MC)), 'set_bp_mc')
      call gpu_check(sisgpu_energy_sumup(h, int(irep - 1, c_int), tempK_in, merge(0, 1, flg_bp_energy), energies), &
                     'energy_sumup')
   end subroutine

   subroutine gpu_energy_kinetic(irep, Ekin)
      integer, intent(in) :: irep
      real(c_double), intent(out) :: Ekin
      call gpu_check(sisgpu_energy_kinetic(h, int(irep - 1, c_int), Ekin), 'energy_kinetic')
   end subroutine

   ! job_dcd (src/job_dcd.F90:65-79): energies of nframes DCD frames (real(4) coordinates as read by
   ! src/dcd.F90:584-592), nrep_proc frames per launch; energies(0:10, nframes)
   subroutine gpu_dcd_energies(nframes, xyz32, tempK_in, energies)
      integer, intent(in) :: nframes
      real(c_float), intent(in) :: xyz32(3, *)
      real(c_double), intent(in) :: tempK_in
      real(c_double), intent(out) :: energies(0:10, *)
      call gpu_check(sisgpu_dcd_energies(h, int(nframes, c_int), xyz32, tempK_in, energies), 'dcd_energies')
   end subroutine

   ! the loop body of src/job_md.F90:374-460 for every local replica, nsteps times, on the device
   subroutine gpu_md_steps(nsteps)
      integer(c_int64_t), intent(in) :: nsteps
      integer(c_int64_t) :: nreb
      call gpu_check(sisgpu_md_steps(h, nsteps, nreb), 'md_steps')
   end subroutine

   subroutine gpu_energy_replica(energies, replica_energies)
      real(c_double), intent(out) :: energies(0:10, *), replica_energies(2, *)
      call gpu_check(sisgpu_energy_replica(h, energies, replica_energies), 'energy_replica')
   end subroutine

   ! uniforms drawn by the caller from mts_rep (src/replica_exchange.F90:44-48) keep the
   ! reference's exchange stream; tables are copied back into var_replica afterwards
   subroutine gpu_replica_exchange(replica_energies, uniforms, n)
      use var_replica, only : rep2lab, lab2rep
      real(c_double), intent(in) :: replica_energies(2, *)
      integer, intent(in) :: n
      real(c_double), intent(in), target :: uniforms(n)
      integer(c_int) :: natt, nacc
      call gpu_check(sisgpu_replica_exchange(h, replica_energies, c_loc(uniforms), natt, nacc), 'replica_exchange')
      call gpu_check(sisgpu_remd_state(h, c_loc(rep2lab), c_loc(lab2rep), c_null_ptr), 'remd_state')
   end subroutine

   subroutine gpu_get_bp_state(irep, status, ene)    ! for write_bp, src/write_bp.F90:36-49
      integer, intent(in) :: irep
      integer(c_int), intent(out) :: status(*)
      real(c_double), intent(out) :: ene(*)
      call gpu_check(sisgpu_get_bp_state(h, int(irep - 1, c_int), status, ene), 'get_bp_state')
   end subroutine

end module sisgpu_mod
