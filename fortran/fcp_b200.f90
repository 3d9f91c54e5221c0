!
! fortran/fcp_b200.f90 -- ISO_C_BINDING layer between freeCappuccino's Fortran host and libfcp_b200.so (include/fcp.h).
!
! Two modules:
!   fcp_b200      the bind(C) interfaces, one per entry point of include/fcp.h
!   fcp_backend   drop-in replacements that keep the reference's module-level API for the hot path
!                 (csrsolve, grad_gauss, grad, grad_w_option, laplacian, gradp_and_sources, calcuvw, calcp_simple, calcp_piso, exchange, global_sum):
!                 same names, same dummy arguments, same module globals (geometry, sparse_matrix, variables), so that
!                 `use linear_solvers` / `use gradients` / `use pressure` in a caller is replaced by `use fcp_backend`
!                 (INTEGRATION.md shows the patch).
!
! This file cannot be compiled in the build image (no Fortran compiler there; SURVEY.md section 0, fact 9).  The same
! call sequence is exercised through host/fcp_host.hpp (C++) and freecappuccino-dev_b200/host.py (Python) by tests/.
! Reference signatures: src/linearSolvers/linear_solvers.f90:40-59 (csrsolve), src/finiteVolume/fvExplicit/gradients.f90
! :23-27,:1607 (grad, grad_gauss), src/finiteVolume/fvImplicit/laplacian.f90 (laplacian), Pressure/nablap.f90:19
! (gradp_and_sources), Pressure/pressure.f90:39-40 (calcp_simple), src-par/exchange.f90:3, src-par/global_sum_mpi.f90:4.
!
module fcp_b200
  use, intrinsic :: iso_c_binding
  implicit none

  ! constants of include/fcp.h
  integer(c_int), parameter :: FCP_OK = 0
  integer(c_int), parameter :: FCP_SOLVER_DPCG = 1, FCP_SOLVER_ICCG = 2, FCP_SOLVER_BICGSTAB = 3, FCP_SOLVER_GAUSS_SEIDEL = 4
  integer(c_int), parameter :: FCP_GRAD_GAUSS = 0, FCP_GRAD_LSQ = 1, FCP_GRAD_LSQ_DM = 2, FCP_GRAD_LSQ_QR = 3
  integer(c_int), parameter :: FCP_LIMITER_NONE = 0, FCP_LIMITER_BARTH_JESPERSEN = 1, FCP_LIMITER_VENKATAKRISHNAN = 2, &
                               FCP_LIMITER_R3 = 3, FCP_LIMITER_MULTIDIMENSIONAL = 4
  integer(c_int), parameter :: FCP_PSCHEME_LINEAR = 0, FCP_PSCHEME_CENTRAL = 1, FCP_PSCHEME_WEIGHTED = 2
  integer(c_int), parameter :: FCP_BC_WALL = 0, FCP_BC_INLET = 1, FCP_BC_OUTLET = 2, FCP_BC_SYMMETRY = 3, &
                               FCP_BC_PRESSURE = 4, FCP_BC_PERIODIC = 5, FCP_BC_EMPTY = 6, FCP_BC_PROCESS = 7
  enum, bind(c)   ! field ids, same order as the enum in include/fcp.h
    enumerator :: FCP_F_U = 0, FCP_F_V, FCP_F_W, FCP_F_P, FCP_F_PP, FCP_F_DEN, FCP_F_VIS, &
                  FCP_F_APU, FCP_F_APV, FCP_F_APW, FCP_F_SU, FCP_F_SV, FCP_F_SW, &
                  FCP_F_S0, FCP_F_S1, FCP_F_S2, FCP_F_S3, &
                  FCP_F_DUDXI, FCP_F_DVDXI, FCP_F_DWDXI, FCP_F_DPDXI, FCP_F_G0, FCP_F_G1, &
                  FCP_F_FLMASS, FCP_F_A, FCP_F_APR, FCP_F_H, FCP_F_RU, FCP_F_RV, FCP_F_RW, FCP_F_VISW, &
                  FCP_F_UO, FCP_F_VO, FCP_F_WO, FCP_F_UOO, FCP_F_VOO, FCP_F_WOO, FCP_F_UOOO, FCP_F_VOOO, FCP_F_WOOO, &
                  FCP_F_SPU, FCP_F_SPV, FCP_F_SP, &
                  FCP_F_TE, FCP_F_ED, FCP_F_PHIO, FCP_F_PHIOO, FCP_F_GEN, FCP_F_MAGSTRAIN, FCP_F_VORTICITY, &
                  FCP_F_DNW, FCP_F_TAU, FCP_F_YPL, FCP_F_SCTMP, &
                  FCP_F_FSST, FCP_F_WALLDIST, FCP_F_DTEDXI, FCP_F_DEDDXI
  end enum

  type, bind(c) :: fcp_mesh_desc
    integer(c_int32_t) :: numCells, numInnerFaces, numBoundaryFaces, numBoundaries
    type(c_ptr) :: owner, neighbour
    type(c_ptr) :: arx, ary, arz, xf, yf, zf, facint, Df, xc, yc, zc, vol
    type(c_ptr) :: bctype, nfaces, startFace
    type(c_ptr) :: startFaceTwin          ! per patch; c_null_ptr when the mesh has no periodic patch
    type(c_ptr) :: DfPeriodic             ! c_null_ptr in the serial tree (the library then reads Df(i) like calcp_simple.f90:199 does)
  end type

  type, bind(c) :: fcp_report
    real(c_double) :: res0, resl, factor, resor
    integer(c_int32_t) :: iters, solver
  end type

  type, bind(c) :: fcp_uvw_params
    integer(c_int32_t) :: solver, maxiter
    real(c_double) :: tol_abs, tol_rel
    real(c_double) :: urf(3)
    real(c_double) :: gds
    integer(c_int32_t) :: cscheme, grad_method, limiter, pscheme, tscheme, piso
    real(c_double) :: timestep
    integer(c_int32_t) :: const_mflux, pad
    real(c_double) :: gradPcmf, viscos
  end type

  integer(c_int), parameter :: FCP_SC_GENERIC = 0, FCP_SC_TKE_RLZB = 1, FCP_SC_EPS_RLZB = 2, FCP_SC_TKE_SST = 3, FCP_SC_OMEGA_SST = 4
  type, bind(c) :: fcp_scalar_params
    integer(c_int32_t) :: kind, solver, maxiter, cscheme, grad_method, limiter, tscheme, lowre
    real(c_double) :: tol_abs, tol_rel, urf, gds, timestep, prtr, viscos, densit
  end type

  type, bind(c) :: fcp_piso_params
    integer(c_int32_t) :: solver, maxiter
    real(c_double) :: tol_abs, tol_rel, urfp
    integer(c_int32_t) :: ncorr, npcor, pscheme, const_mflux
    real(c_double) :: flomas
  end type

  type, bind(c) :: fcp_simple_params
    integer(c_int32_t) :: solver, maxiter
    real(c_double) :: tol_abs, tol_rel, urfp
    integer(c_int32_t) :: npcor, pRefCell, pscheme, const_mflux
    real(c_double) :: flomas
    integer(c_int32_t) :: zero_pp
  end type

  interface
    function fcp_last_error() bind(c, name='fcp_last_error') result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function
    function fcp_ctx_create(mesh, device, ctx) bind(c, name='fcp_ctx_create') result(rc)
      import :: c_int, c_ptr, fcp_mesh_desc
      type(fcp_mesh_desc), intent(in) :: mesh
      integer(c_int), value :: device
      type(c_ptr), intent(out) :: ctx
      integer(c_int) :: rc
    end function
    function fcp_ctx_destroy(ctx) bind(c, name='fcp_ctx_destroy') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function
    function fcp_csr_pattern(ctx, ia, ja, diag, icell_jcell, jcell_icell) bind(c, name='fcp_csr_pattern') result(rc)
      import :: c_int, c_ptr, c_int32_t
      type(c_ptr), value :: ctx
      integer(c_int32_t), intent(out) :: ia(*), ja(*), diag(*), icell_jcell(*), jcell_icell(*)
      integer(c_int) :: rc
    end function
    function fcp_field_upload(ctx, field, host, count) bind(c, name='fcp_field_upload') result(rc)
      import :: c_int, c_ptr, c_double, c_int64_t
      type(c_ptr), value :: ctx
      integer(c_int), value :: field
      real(c_double), intent(in) :: host(*)
      integer(c_int64_t), value :: count
      integer(c_int) :: rc
    end function
    function fcp_field_download(ctx, field, host, count) bind(c, name='fcp_field_download') result(rc)
      import :: c_int, c_ptr, c_double, c_int64_t
      type(c_ptr), value :: ctx
      integer(c_int), value :: field
      real(c_double), intent(out) :: host(*)
      integer(c_int64_t), value :: count
      integer(c_int) :: rc
    end function
    function fcp_field_fill(ctx, field, val) bind(c, name='fcp_field_fill') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: field
      real(c_double), value :: val
      integer(c_int) :: rc
    end function
    function fcp_spmv(ctx, x_field, y_field) bind(c, name='fcp_spmv') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: x_field, y_field
      integer(c_int) :: rc
    end function
    function fcp_csrsolve(ctx, solver, fi_field, rhs_field, itr_max, tol_abs, tol_rel, rep) bind(c, name='fcp_csrsolve') result(rc)
      import :: c_int, c_ptr, c_double, c_int32_t, fcp_report
      type(c_ptr), value :: ctx
      integer(c_int), value :: solver, fi_field, rhs_field
      integer(c_int32_t), value :: itr_max
      real(c_double), value :: tol_abs, tol_rel
      type(fcp_report), intent(out) :: rep
      integer(c_int) :: rc
    end function
    function fcp_report_line(rep, chvar, buf, buflen) bind(c, name='fcp_report_line') result(rc)
      import :: c_int, c_char, fcp_report
      type(fcp_report), intent(in) :: rep
      character(kind=c_char), intent(in) :: chvar(*)
      character(kind=c_char), intent(out) :: buf(*)
      integer(c_int), value :: buflen
      integer(c_int) :: rc
    end function
    function fcp_create_lsq_grad_matrix(ctx, method) bind(c, name='fcp_create_lsq_grad_matrix') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: method
      integer(c_int) :: rc
    end function
    function fcp_grad(ctx, method, phi_field, grad_field, lsq_row2_reference) bind(c, name='fcp_grad') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: method, phi_field, grad_field, lsq_row2_reference
      integer(c_int) :: rc
    end function
    function fcp_laplacian(ctx, mu_field, phi_field) bind(c, name='fcp_laplacian') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: mu_field, phi_field
      integer(c_int) :: rc
    end function
    function fcp_gradp_and_sources(ctx, pscheme, p_field) bind(c, name='fcp_gradp_and_sources') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: pscheme, p_field
      integer(c_int) :: rc
    end function
    function fcp_calcp_simple(ctx, prm, rep) bind(c, name='fcp_calcp_simple') result(rc)
      import :: c_int, c_ptr, fcp_simple_params, fcp_report
      type(c_ptr), value :: ctx
      type(fcp_simple_params), intent(in) :: prm
      type(fcp_report), intent(out) :: rep(*)
      integer(c_int) :: rc
    end function
    function fcp_calcuvw(ctx, prm, rep) bind(c, name='fcp_calcuvw') result(rc)
      import :: c_int, c_ptr, fcp_uvw_params, fcp_report
      type(c_ptr), value :: ctx
      type(fcp_uvw_params), intent(in) :: prm
      type(fcp_report), intent(out) :: rep(3)
      integer(c_int) :: rc
    end function
    function fcp_calcp_piso(ctx, prm, rep) bind(c, name='fcp_calcp_piso') result(rc)
      import :: c_int, c_ptr, fcp_piso_params, fcp_report
      type(c_ptr), value :: ctx
      type(fcp_piso_params), intent(in) :: prm
      type(fcp_report), intent(out) :: rep(*)
      integer(c_int) :: rc
    end function
    function fcp_calcsc(ctx, prm, phi_field, rep, fimin, fimax) bind(c, name='fcp_calcsc') result(rc)
      import :: c_int, c_ptr, c_double, fcp_scalar_params, fcp_report
      type(c_ptr), value :: ctx
      type(fcp_scalar_params), intent(in) :: prm
      integer(c_int), value :: phi_field
      type(fcp_report), intent(out) :: rep
      real(c_double), intent(out) :: fimin, fimax
      integer(c_int) :: rc
    end function
    function fcp_calc_strain_and_vorticity(ctx) bind(c, name='fcp_calc_strain_and_vorticity') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function
    function fcp_modify_mu_eff_k_epsilon_rlzb(ctx, urfVis, viscos) bind(c, name='fcp_modify_mu_eff_k_epsilon_rlzb') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), value :: urfVis, viscos
      integer(c_int) :: rc
    end function
    function fcp_modify_mu_eff_k_omega_sst(ctx, urfVis, viscos, densit, lowre) bind(c, name='fcp_modify_mu_eff_k_omega_sst') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), value :: urfVis, viscos, densit
      integer(c_int), value :: lowre
      integer(c_int) :: rc
    end function
    function fcp_wall_distance(ctx, rep) bind(c, name='fcp_wall_distance') result(rc)
      import :: c_int, c_ptr, fcp_report
      type(c_ptr), value :: ctx
      type(fcp_report), intent(out) :: rep
      integer(c_int) :: rc
    end function
    function fcp_grad_gauss_fvx(ctx, phi_field, grad_field) bind(c, name='fcp_grad_gauss_fvx') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: phi_field, grad_field
      integer(c_int) :: rc
    end function
    function fcp_grad_gauss_iter(ctx, phi_field, grad_field, nigrad) bind(c, name='fcp_grad_gauss_iter') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: phi_field, grad_field, nigrad
      integer(c_int) :: rc
    end function
    function fcp_modify_viscosity_sgs(ctx, model, urfVis, viscos) bind(c, name='fcp_modify_viscosity_sgs') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: model
      real(c_double), value :: urfVis, viscos
      integer(c_int) :: rc
    end function
    function fcp_constant_mass_flow_forcing(ctx, magUbar, gradPcmf, magUbarStar) bind(c, name='fcp_constant_mass_flow_forcing') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), value :: magUbar
      real(c_double), intent(inout) :: gradPcmf
      real(c_double), intent(out) :: magUbarStar
      integer(c_int) :: rc
    end function
    function fcp_update_boundary(ctx, field) bind(c, name='fcp_update_boundary') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: field
      integer(c_int) :: rc
    end function
    function fcp_slope_limiter(ctx, limiter, phi_field, grad_field) bind(c, name='fcp_slope_limiter') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: limiter, phi_field, grad_field
      integer(c_int) :: rc
    end function
    function fcp_grad_opt(ctx, method, limiter, phi_field, grad_field) bind(c, name='fcp_grad_opt') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: method, limiter, phi_field, grad_field
      integer(c_int) :: rc
    end function
    function fcp_solver_create(n, nnz, ia, ja, diag, device, s) bind(c, name='fcp_solver_create') result(rc)
      import :: c_int, c_ptr, c_int32_t
      integer(c_int32_t), value :: n, nnz
      integer(c_int32_t), intent(in) :: ia(*), ja(*), diag(*)
      integer(c_int), value :: device
      type(c_ptr), intent(out) :: s
      integer(c_int) :: rc
    end function
    function fcp_solver_solve(s, solver, a, fi, rhs, itr_max, tol_abs, tol_rel, rep) bind(c, name='fcp_solver_solve') result(rc)
      import :: c_int, c_ptr, c_double, c_int32_t, fcp_report
      type(c_ptr), value :: s
      integer(c_int), value :: solver
      real(c_double), intent(in) :: a(*), rhs(*)
      real(c_double), intent(inout) :: fi(*)
      integer(c_int32_t), value :: itr_max
      real(c_double), value :: tol_abs, tol_rel
      type(fcp_report), intent(out) :: rep
      integer(c_int) :: rc
    end function
    function fcp_solver_destroy(s) bind(c, name='fcp_solver_destroy') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: s
      integer(c_int) :: rc
    end function
    function fcp_comm_unique_id(id128) bind(c, name='fcp_comm_unique_id') result(rc)
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id128(128)
      integer(c_int) :: rc
    end function
    function fcp_comm_init(ctx, rank, nranks, id128, peer_rank) bind(c, name='fcp_comm_init') result(rc)
      import :: c_int, c_ptr, c_char, c_int32_t
      type(c_ptr), value :: ctx
      integer(c_int), value :: rank, nranks
      character(kind=c_char), intent(in) :: id128(128)
      integer(c_int32_t), intent(in) :: peer_rank(*)
      integer(c_int) :: rc
    end function
    function fcp_set_flux_variant(ctx, variant, grad_method) bind(c, name='fcp_set_flux_variant') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: variant, grad_method
      integer(c_int) :: rc
    end function
    function fcp_set_process_facint(ctx, fpro, count) bind(c, name='fcp_set_process_facint') result(rc)
      import :: c_int, c_ptr, c_double, c_int32_t
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: fpro(*)
      integer(c_int32_t), value :: count
      integer(c_int) :: rc
    end function
    function fcp_set_process_orientation(ctx, flipped, count) bind(c, name='fcp_set_process_orientation') result(rc)
      import :: c_int, c_ptr, c_int32_t
      type(c_ptr), value :: ctx
      integer(c_int32_t), intent(in) :: flipped(*)
      integer(c_int32_t), value :: count
      integer(c_int) :: rc
    end function
    function fcp_exchange(ctx, field) bind(c, name='fcp_exchange') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: field
      integer(c_int) :: rc
    end function
    function fcp_global_sum(ctx, val) bind(c, name='fcp_global_sum') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: val
      integer(c_int) :: rc
    end function
    function fcp_global_max(ctx, val) bind(c, name='fcp_global_max') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: val
      integer(c_int) :: rc
    end function
    function fcp_global_min(ctx, val) bind(c, name='fcp_global_min') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: val
      integer(c_int) :: rc
    end function
    function fcp_global_isum(ctx, val) bind(c, name='fcp_global_isum') result(rc)
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: ctx
      integer(c_int64_t), intent(inout) :: val
      integer(c_int) :: rc
    end function
  end interface
end module fcp_b200


module fcp_backend
  use, intrinsic :: iso_c_binding
  use types, only: dp
  use parameters          ! numTotal, pRefCell, npcor, const_mflux, flomas ...
  use geometry            ! numCells, numInnerFaces, owner, neighbour, arx ... (src/mesh/geometry.f90:12-86)
  use sparse_matrix       ! nnz, ia, ja, diag, a, su, sv, sw, apu, apv, apw    (src/sparseMatrix/sparse_matrix.f90:20-40)
  use variables           ! u, v, w, p, pp, den, flmass, dPdxi ...
  use fcp_b200
  implicit none
  type(c_ptr), save :: ctx = c_null_ptr
  public

contains

  subroutine fcp_check(rc, what)
    ! the reference has no status codes: fatal conditions print and stop (linear_solvers.f90:1361-1376)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: what
    if (rc /= FCP_OK) then
      write(*,'(3a,i0)') ' libfcp_b200: ', what, ' failed with code ', rc
      stop
    end if
  end subroutine

  integer(c_int) function solver_id(solver)
    character(len=*), intent(in) :: solver
    select case (trim(solver))
    case ('dpcg');     solver_id = FCP_SOLVER_DPCG
    case ('iccg');     solver_id = FCP_SOLVER_ICCG
    case ('bicgstab'); solver_id = FCP_SOLVER_BICGSTAB
    case ('gauss-seidel'); solver_id = FCP_SOLVER_GAUSS_SEIDEL
    case default
      write(*,'(3a)') ' libfcp_b200: linear solver "', trim(solver), '" is not on the accelerated path'
      stop
    end select
  end function

  ! Call once after read_mesh + create_CSR_matrix (main.f90:112-114): uploads the mesh, builds the device pattern and
  ! checks that it is identical to the one create_CSR_matrix produced.
  subroutine fcp_init(device)
    integer, intent(in) :: device
    type(fcp_mesh_desc) :: md
    integer(c_int32_t), allocatable, target :: bct(:), nfa(:), sfa(:), stw(:)
    integer(c_int32_t), allocatable :: ia2(:), ja2(:), dg2(:), k1(:), k2(:)
    integer :: ib, iPer
    allocate(bct(numBoundaries), nfa(numBoundaries), sfa(numBoundaries), stw(numBoundaries))
    iPer = 0
    do ib = 1, numBoundaries
      nfa(ib) = nfaces(ib)
      sfa(ib) = startFace(ib)                     ! 0-based offset, as in the boundary file (geometry.f90:282-290)
      stw(ib) = -1
      if (trim(bctype(ib)) == 'periodic') then    ! the reference indexes startFaceTwin by the running count of periodic patches (geometry.f90:252-257)
        iPer = iPer + 1
        stw(ib) = startFaceTwin(iPer)
      end if
      select case (trim(bctype(ib)))
      case ('wall');     bct(ib) = FCP_BC_WALL
      case ('inlet');    bct(ib) = FCP_BC_INLET
      case ('outlet');   bct(ib) = FCP_BC_OUTLET
      case ('symmetry'); bct(ib) = FCP_BC_SYMMETRY
      case ('pressure'); bct(ib) = FCP_BC_PRESSURE
      case ('periodic'); bct(ib) = FCP_BC_PERIODIC
      case ('process');  bct(ib) = FCP_BC_PROCESS
      case default;      bct(ib) = FCP_BC_EMPTY
      end select
    end do
    md%numCells = numCells; md%numInnerFaces = numInnerFaces
    md%numBoundaryFaces = numBoundaryFaces; md%numBoundaries = numBoundaries
    md%owner = c_loc(owner); md%neighbour = c_loc(neighbour)
    md%arx = c_loc(arx); md%ary = c_loc(ary); md%arz = c_loc(arz)
    md%xf = c_loc(xf); md%yf = c_loc(yf); md%zf = c_loc(zf)
    md%facint = c_loc(facint); md%Df = c_loc(Df)
    md%xc = c_loc(xc); md%yc = c_loc(yc); md%zc = c_loc(zc); md%vol = c_loc(vol)
    md%bctype = c_loc(bct); md%nfaces = c_loc(nfa); md%startFace = c_loc(sfa)
    md%startFaceTwin = c_null_ptr
    md%DfPeriodic = c_null_ptr
    if (iPer > 0) md%startFaceTwin = c_loc(stw)
    call fcp_check(fcp_ctx_create(md, int(device, c_int), ctx), 'fcp_ctx_create')
    allocate(ia2(numCells+1), ja2(nnz), dg2(numCells), k1(numInnerFaces+numPeriodic), k2(numInnerFaces+numPeriodic))
    call fcp_check(fcp_csr_pattern(ctx, ia2, ja2, dg2, k1, k2), 'fcp_csr_pattern')
    if (any(ia2 /= ia) .or. any(ja2 /= ja) .or. any(dg2 /= diag) .or. &
        any(k1 /= icell_jcell_csr_index(1:numInnerFaces+numPeriodic)) .or. any(k2 /= jcell_icell_csr_index(1:numInnerFaces+numPeriodic))) then
      write(*,'(a)') ' libfcp_b200: device CSR pattern differs from create_CSR_matrix'
      stop
    end if
  end subroutine

  subroutine put(field, x, n)
    integer(c_int), intent(in) :: field
    real(dp), intent(in) :: x(*)
    integer, intent(in) :: n
    call fcp_check(fcp_field_upload(ctx, field, x, int(n, c_int64_t)), 'fcp_field_upload')
  end subroutine
  subroutine get(field, x, n)
    integer(c_int), intent(in) :: field
    real(dp), intent(out) :: x(*)
    integer, intent(in) :: n
    call fcp_check(fcp_field_download(ctx, field, x, int(n, c_int64_t)), 'fcp_field_download')
  end subroutine

  ! ---- csrsolve: same dummies as linear_solvers.f90:40-59 ---------------------------------------------------------------
  subroutine csrsolve(solver, fi, rhs, res0, itr_max, tol_abs, tol_rel, chvar)
    character(len=*), intent(in) :: solver
    real(dp), dimension(numTotal), intent(inout) :: fi
    real(dp), dimension(numCells), intent(in) :: rhs
    real(dp), intent(out) :: res0
    integer, intent(in) :: itr_max
    real(dp), intent(in) :: tol_abs, tol_rel
    character(len=*), intent(in) :: chvar
    type(fcp_report) :: rep
    character(kind=c_char) :: line(256)
    integer :: i
    call put(FCP_F_A, a, nnz)
    call put(FCP_F_S0, fi, numTotal)
    call put(FCP_F_S1, rhs, numCells)
    call fcp_check(fcp_csrsolve(ctx, solver_id(solver), FCP_F_S0, FCP_F_S1, int(itr_max, c_int32_t), tol_abs, tol_rel, rep), 'fcp_csrsolve')
    call get(FCP_F_S0, fi, numCells)
    res0 = rep%resor
    ! the report line of linear_solvers.f90:354-355 / 540-541 / 781-782 (parsed by examples/*/plotResiduals)
    call fcp_check(fcp_report_line(rep, trim(chvar)//c_null_char, line, 256_c_int), 'fcp_report_line')
    do i = 1, 256
      if (line(i) == c_null_char) exit
    end do
    write(*,'(256a)') line(1:i-1)
  end subroutine

  ! ---- gradients: gradients.f90:1607 (grad_gauss), :23-27 (grad) ------------------------------------------------------------
  subroutine grad_gauss(u_, dudxi)
    real(dp), dimension(numTotal), intent(in) :: u_
    real(dp), dimension(3,numTotal), intent(inout) :: dudxi
    call put(FCP_F_S0, u_, numTotal)
    call fcp_check(fcp_grad(ctx, FCP_GRAD_GAUSS, FCP_F_S0, FCP_F_G0, 1_c_int), 'fcp_grad')
    call get(FCP_F_G0, dudxi, 3*numTotal)
  end subroutine

  subroutine grad(phi, dPhidxi)         ! dispatch on the logicals lstsq / lstsq_dm like grad_scalar_field, gradients.f90:106-163
    use gradients, only: lstsq, lstsq_dm
    real(dp), dimension(numTotal), intent(in) :: phi
    real(dp), dimension(3,numTotal), intent(inout) :: dPhidxi
    integer(c_int) :: method
    method = FCP_GRAD_GAUSS
    if (lstsq) method = FCP_GRAD_LSQ
    if (lstsq_dm) method = FCP_GRAD_LSQ_DM
    call put(FCP_F_S0, phi, numTotal)
    call fcp_check(fcp_grad(ctx, method, FCP_F_S0, FCP_F_G0, 1_c_int), 'fcp_grad')
    call get(FCP_F_G0, dPhidxi, 3*numTotal)
  end subroutine

  ! grad(phi,dPhidxi,option,option_limiter)   grad_scalar_field_w_option, gradients.f90:217-278
  subroutine grad_w_option(phi, dPhidxi, option, option_limiter)
    real(dp), dimension(numTotal), intent(in) :: phi
    real(dp), dimension(3,numTotal), intent(inout) :: dPhidxi
    character(len=*), intent(in) :: option, option_limiter
    integer(c_int) :: method, limiter
    dPhidxi = 0.0_dp
    select case (option)
      case ('lsq');    method = FCP_GRAD_LSQ
      case ('lsq_qr'); method = FCP_GRAD_LSQ_QR
      case ('wlsq');   method = FCP_GRAD_LSQ_DM
      case ('gauss');  method = FCP_GRAD_GAUSS
      case default;    return
    end select
    select case (option_limiter)
      case ('Barth-Jespersen');  limiter = FCP_LIMITER_BARTH_JESPERSEN
      case ('Venkatakrishnan');  limiter = FCP_LIMITER_VENKATAKRISHNAN
      case ('R3');               limiter = FCP_LIMITER_R3
      case ('multidimensional'); limiter = FCP_LIMITER_MULTIDIMENSIONAL
      case default;              limiter = FCP_LIMITER_NONE
    end select
    call put(FCP_F_S0, phi, numTotal)
    if (method /= FCP_GRAD_GAUSS) call fcp_check(fcp_create_lsq_grad_matrix(ctx, method), 'fcp_create_lsq_grad_matrix')
    call fcp_check(fcp_grad_opt(ctx, method, limiter, FCP_F_S0, FCP_F_G0), 'fcp_grad_opt')
    call get(FCP_F_G0, dPhidxi, 3*numTotal)
  end subroutine

  subroutine create_lsq_grad_matrix()   ! gradients.f90:72-101
    use gradients, only: lstsq, lstsq_dm
    if (lstsq) call fcp_check(fcp_create_lsq_grad_matrix(ctx, FCP_GRAD_LSQ), 'fcp_create_lsq_grad_matrix')
    if (lstsq_dm) call fcp_check(fcp_create_lsq_grad_matrix(ctx, FCP_GRAD_LSQ_DM), 'fcp_create_lsq_grad_matrix')
  end subroutine

  ! ---- laplacian(mu,phi): fills the module's a(:) and su(:)   fvImplicit/laplacian.f90 ------------------------------------
  subroutine laplacian(mu, phi)
    real(dp), dimension(numCells), intent(in) :: mu
    real(dp), dimension(numTotal), intent(in) :: phi
    call put(FCP_F_S0, mu, numCells)
    call put(FCP_F_S1, phi, numTotal)
    call put(FCP_F_SU, su, numCells)
    call fcp_check(fcp_laplacian(ctx, FCP_F_S0, FCP_F_S1), 'fcp_laplacian')
    call get(FCP_F_A, a, nnz)
    call get(FCP_F_SU, su, numCells)
  end subroutine

  ! ---- gradp_and_sources(p): fills su, sv, sw, dPdxi   Pressure/nablap.f90:19 ------------------------------------------------
  subroutine gradp_and_sources(p_)
    use nablap, only: pscheme
    real(dp), dimension(numTotal), intent(inout) :: p_
    integer(c_int) :: ps
    ps = FCP_PSCHEME_LINEAR
    if (trim(pscheme) == 'central') ps = FCP_PSCHEME_CENTRAL
    if (trim(pscheme) == 'weighted') ps = FCP_PSCHEME_WEIGHTED
    call put(FCP_F_P, p_, numTotal)
    call put(FCP_F_APU, apu, numCells)
    call put(FCP_F_DPDXI, dPdxi, 3*numTotal)
    call fcp_check(fcp_gradp_and_sources(ctx, ps, FCP_F_P), 'fcp_gradp_and_sources')
    call get(FCP_F_P, p_, numTotal)
    call get(FCP_F_SU, su, numCells); call get(FCP_F_SV, sv, numCells); call get(FCP_F_SW, sw, numCells)
    call get(FCP_F_DPDXI, dPdxi, 3*numTotal)
  end subroutine

  ! ---- calcp_simple(): no arguments, everything through modules   Pressure/calcp_simple.f90 --------------------------------
  subroutine calcp_simple()
    use pressure, only: urfP, lSolverP, maxiterP, tolAbsP, tolRelP
    use nablap, only: pscheme
    type(fcp_simple_params) :: prm
    type(fcp_report) :: rep(8)
    character(kind=c_char) :: line(256)
    integer :: ipcorr, i
    call put(FCP_F_U, u, numTotal); call put(FCP_F_V, v, numTotal); call put(FCP_F_W, w, numTotal)
    call put(FCP_F_P, p, numTotal); call put(FCP_F_PP, pp, numTotal); call put(FCP_F_DEN, den, numTotal)
    call put(FCP_F_APU, apu, numCells); call put(FCP_F_APV, apv, numCells); call put(FCP_F_APW, apw, numCells)
    call put(FCP_F_DPDXI, dPdxi, 3*numTotal)
    call put(FCP_F_FLMASS, flmass, numFaces)        ! inlet fluxes are prescribed by the host
    prm%solver = solver_id(lSolverP); prm%maxiter = maxiterP
    prm%tol_abs = tolAbsP; prm%tol_rel = tolRelP; prm%urfp = urfP
    prm%npcor = npcor; prm%pRefCell = pRefCell
    prm%pscheme = FCP_PSCHEME_LINEAR
    if (trim(pscheme) == 'central') prm%pscheme = FCP_PSCHEME_CENTRAL
    if (trim(pscheme) == 'weighted') prm%pscheme = FCP_PSCHEME_WEIGHTED
    prm%const_mflux = merge(1, 0, const_mflux); prm%flomas = flomas
    prm%zero_pp = 0                                ! the serial tree warm-starts pp (quirk Q8)
    call fcp_check(fcp_calcp_simple(ctx, prm, rep), 'fcp_calcp_simple')
    do ipcorr = 1, npcor
      call fcp_check(fcp_report_line(rep(ipcorr), 'p'//c_null_char, line, 256_c_int), 'fcp_report_line')
      do i = 1, 256
        if (line(i) == c_null_char) exit
      end do
      write(*,'(256a)') line(1:i-1)
    end do
    call get(FCP_F_U, u, numTotal); call get(FCP_F_V, v, numTotal); call get(FCP_F_W, w, numTotal)
    call get(FCP_F_P, p, numTotal); call get(FCP_F_PP, pp, numTotal)
    call get(FCP_F_FLMASS, flmass, numFaces); call get(FCP_F_DPDXI, dPdxi, 3*numTotal)
    call get(FCP_F_SU, su, numCells); call get(FCP_F_SV, sv, numCells); call get(FCP_F_SW, sw, numCells)
    call get(FCP_F_A, a, nnz)
    call continuityErrors                            ! calcp_simple.f90:464 stays on the host (prints, sets resor(4))
  end subroutine

  ! ---- calcuvw(): no arguments   Velocity/velocity.f90:50-750 (tier "next" row f1) -----------------------------------------
  integer(c_int) function grad_method_id()          ! the logicals of gradients.f90:118-138
    use gradients, only: lstsq, lstsq_qr, lstsq_dm
    grad_method_id = FCP_GRAD_GAUSS
    if (lstsq) then
      grad_method_id = FCP_GRAD_LSQ
    else if (lstsq_qr) then
      grad_method_id = FCP_GRAD_LSQ_QR
    else if (lstsq_dm) then
      grad_method_id = FCP_GRAD_LSQ_DM
    end if
  end function
  integer(c_int) function limiter_id()              ! gradients.f90:140-160
    use gradients, only: limiter
    select case (limiter)
      case ('Barth-Jespersen');  limiter_id = FCP_LIMITER_BARTH_JESPERSEN
      case ('Venkatakrishnan');  limiter_id = FCP_LIMITER_VENKATAKRISHNAN
      case ('R3');               limiter_id = FCP_LIMITER_R3
      case ('multidimensional'); limiter_id = FCP_LIMITER_MULTIDIMENSIONAL
      case default;              limiter_id = FCP_LIMITER_NONE
    end select
  end function
  integer(c_int) function cscheme_id(scheme)      ! cSchemeU strings of interpolation.f90:28-113, :596-640 in source order
    character(len=*), intent(in) :: scheme
    character(len=24), parameter :: names(20) = [character(len=24) :: 'cds', 'central', 'linearUpwind', 'kappa', 'muscl', 'umist', &
      'koren', 'smart', 'avl-smart', 'charm', 'vanleer', 'ospre', 'minmod', 'boundedLinearUpwind', 'boundedLinearUpwind02', &
      'boundedCentral', 'fromm', 'cui', 'quick', 'spl13']
    integer :: k
    cscheme_id = -1
    do k = 1, 20
      if (trim(scheme) == trim(names(k))) cscheme_id = k - 1
    end do
    if (cscheme_id < 0) then
      write(*,'(a)') 'Fatal error: non-existing interpolation scheme!'    ! interpolation.f90:643-646
      stop
    end if
  end function

  subroutine calcuvw()
    use velocity, only: urfU, gdsU, cSchemeU, lSolverU, maxiterU, tolAbsU, tolRelU
    use gradients, only: lstsq, lstsq_qr, lstsq_dm, limiter
    use nablap, only: pscheme
    use mhd, only: calcEpot                          ! MHD/mhd.f90:21
    type(fcp_uvw_params) :: prm
    type(fcp_report) :: rep(3)
    character(kind=c_char) :: line(256)
    character(len=1), parameter :: chvar(3) = ['U', 'V', 'W']
    real(dp), allocatable :: viswf(:)
    integer :: ib, i, iWall, k
    if (CN .or. lbuoy .or. calcEpot) then
      write(*,'(a)') ' libfcp_b200: Crank-Nicolson, buoyancy and MHD terms of calcuvw are not on the accelerated path'
      stop
    end if
    call put(FCP_F_U, u, numTotal); call put(FCP_F_V, v, numTotal); call put(FCP_F_W, w, numTotal)
    call put(FCP_F_P, p, numTotal); call put(FCP_F_DEN, den, numTotal); call put(FCP_F_VIS, vis, numTotal)
    call put(FCP_F_APU, apu, numCells)               ! the 'weighted' pscheme reads the previous apu (nablap.f90:87)
    allocate(viswf(numTotal)); viswf = 0.0_dp          ! visw(iWall), wall faces in patch order (velocity.f90:441-443) -> boundary slots
    iWall = 0
    do ib = 1, numBoundaries
      if (bctype(ib) == 'wall') then
        do i = 1, nfaces(ib)
          iWall = iWall + 1
          viswf(iBndValueStart(ib) + i) = visw(iWall)
        end do
      end if
    end do
    call put(FCP_F_VISW, viswf, numTotal)
    call put(FCP_F_FLMASS, flmass, numFaces)
    call put(FCP_F_A, a, nnz)                        ! the stale diagonal enters the first row sum (velocity.f90:606)
    prm%tscheme = 0
    if (ltransient) then
      if (bdf) prm%tscheme = 1
      if (bdf2) prm%tscheme = 2
      if (bdf3) prm%tscheme = 3
      call put(FCP_F_UO, uo, numTotal); call put(FCP_F_VO, vo, numTotal); call put(FCP_F_WO, wo, numTotal)
      if (prm%tscheme >= 2) then
        call put(FCP_F_UOO, uoo, numTotal); call put(FCP_F_VOO, voo, numTotal); call put(FCP_F_WOO, woo, numTotal)
      end if
      if (prm%tscheme >= 3) then
        call put(FCP_F_UOOO, uooo, numTotal); call put(FCP_F_VOOO, vooo, numTotal); call put(FCP_F_WOOO, wooo, numTotal)
      end if
    end if
    prm%solver = solver_id(lSolverU); prm%maxiter = maxiterU; prm%tol_abs = tolAbsU; prm%tol_rel = tolRelU
    prm%urf = urfU(1:3); prm%gds = gdsU; prm%cscheme = cscheme_id(cSchemeU)
    prm%grad_method = FCP_GRAD_GAUSS
    if (lstsq) then
      prm%grad_method = FCP_GRAD_LSQ
    else if (lstsq_qr) then
      prm%grad_method = FCP_GRAD_LSQ_QR
    else if (lstsq_dm) then
      prm%grad_method = FCP_GRAD_LSQ_DM
    end if
    select case (limiter)
      case ('Barth-Jespersen');  prm%limiter = FCP_LIMITER_BARTH_JESPERSEN
      case ('Venkatakrishnan');  prm%limiter = FCP_LIMITER_VENKATAKRISHNAN
      case ('R3');               prm%limiter = FCP_LIMITER_R3
      case ('multidimensional'); prm%limiter = FCP_LIMITER_MULTIDIMENSIONAL
      case default;              prm%limiter = FCP_LIMITER_NONE
    end select
    prm%pscheme = FCP_PSCHEME_LINEAR
    if (trim(pscheme) == 'central') prm%pscheme = FCP_PSCHEME_CENTRAL
    if (trim(pscheme) == 'weighted') prm%pscheme = FCP_PSCHEME_WEIGHTED
    prm%piso = merge(1, 0, piso); prm%timestep = timestep
    prm%const_mflux = merge(1, 0, const_mflux); prm%pad = 0; prm%gradPcmf = gradPcmf; prm%viscos = viscos
    call fcp_check(fcp_calcuvw(ctx, prm, rep), 'fcp_calcuvw')
    do k = 1, 3
      call fcp_check(fcp_report_line(rep(k), chvar(k)//c_null_char, line, 256_c_int), 'fcp_report_line')
      do i = 1, 256
        if (line(i) == c_null_char) exit
      end do
      write(*,'(256a)') line(1:i-1)
      resor(k) = rep(k)%resor
    end do
    call get(FCP_F_U, u, numTotal); call get(FCP_F_V, v, numTotal); call get(FCP_F_W, w, numTotal); call get(FCP_F_P, p, numTotal)
    call get(FCP_F_APU, apu, numCells); call get(FCP_F_APV, apv, numCells); call get(FCP_F_APW, apw, numCells)
    call get(FCP_F_SU, su, numCells); call get(FCP_F_SV, sv, numCells); call get(FCP_F_SW, sw, numCells)
    call get(FCP_F_SPU, spu, numCells); call get(FCP_F_SPV, spv, numCells); call get(FCP_F_SP, sp, numCells)
    call get(FCP_F_DUDXI, dUdxi, 3*numTotal); call get(FCP_F_DVDXI, dVdxi, 3*numTotal); call get(FCP_F_DWDXI, dWdxi, 3*numTotal)
    call get(FCP_F_DPDXI, dPdxi, 3*numTotal)
    call get(FCP_F_A, a, nnz)
    if (piso) then
      call get(FCP_F_RU, rU, numCells); call get(FCP_F_RV, rV, numCells); call get(FCP_F_RW, rW, numCells)
    end if
    deallocate(viswf)
  end subroutine

  ! ---- calcp_piso(): no arguments   Pressure/calcp_piso.f90 ------------------------------------------------------------------
  subroutine calcp_piso()
    use pressure, only: urfP, lSolverP, maxiterP, tolAbsP, tolRelP
    use nablap, only: pscheme
    type(fcp_piso_params) :: prm
    type(fcp_report) :: rep(64)
    character(kind=c_char) :: line(256)
    integer :: k, i
    call put(FCP_F_U, u, numTotal); call put(FCP_F_V, v, numTotal); call put(FCP_F_W, w, numTotal)
    call put(FCP_F_P, p, numTotal); call put(FCP_F_PP, pp, numTotal); call put(FCP_F_DEN, den, numTotal)
    call put(FCP_F_APU, apu, numCells); call put(FCP_F_APV, apv, numCells); call put(FCP_F_APW, apw, numCells)
    call put(FCP_F_RU, rU, numCells); call put(FCP_F_RV, rV, numCells); call put(FCP_F_RW, rW, numCells)
    call put(FCP_F_A, a, nnz)                       ! momentum coefficients: the library does `h = a` (calcp_piso.f90:81)
    call put(FCP_F_DPDXI, dPdxi, 3*numTotal)
    call put(FCP_F_FLMASS, flmass, numFaces)
    prm%solver = solver_id(lSolverP); prm%maxiter = maxiterP
    prm%tol_abs = tolAbsP; prm%tol_rel = tolRelP; prm%urfp = urfP
    prm%ncorr = ncorr; prm%npcor = npcor
    prm%pscheme = FCP_PSCHEME_LINEAR
    if (trim(pscheme) == 'central') prm%pscheme = FCP_PSCHEME_CENTRAL
    if (trim(pscheme) == 'weighted') prm%pscheme = FCP_PSCHEME_WEIGHTED
    prm%const_mflux = merge(1, 0, const_mflux); prm%flomas = flomas
    call fcp_check(fcp_calcp_piso(ctx, prm, rep), 'fcp_calcp_piso')
    do k = 1, ncorr*npcor
      call fcp_check(fcp_report_line(rep(k), 'p'//c_null_char, line, 256_c_int), 'fcp_report_line')
      do i = 1, 256
        if (line(i) == c_null_char) exit
      end do
      write(*,'(256a)') line(1:i-1)
    end do
    call get(FCP_F_U, u, numTotal); call get(FCP_F_V, v, numTotal); call get(FCP_F_W, w, numTotal)
    call get(FCP_F_P, p, numTotal); call get(FCP_F_PP, pp, numTotal)
    call get(FCP_F_FLMASS, flmass, numFaces); call get(FCP_F_DPDXI, dPdxi, 3*numTotal)
    call get(FCP_F_SU, su, numCells); call get(FCP_F_SV, sv, numCells); call get(FCP_F_SW, sw, numCells)
    call get(FCP_F_A, a, nnz); call get(FCP_F_H, h, nnz)
    call continuityErrors                            ! calcp_piso.f90:390 stays on the host
    if (const_mflux) call constant_mass_flow_forcing ! :487
  end subroutine

  ! ---- wall_distance()   src/mesh/wall_distance.f90:55-133 (module geometry owns wallDistance(numCells)) ------------------------------------
  subroutine wall_distance()
    type(fcp_report) :: rep
    character(kind=c_char) :: line(256)
    integer :: i
    call fcp_check(fcp_wall_distance(ctx, rep), 'fcp_wall_distance')
    call fcp_check(fcp_report_line(rep, 'Wdis'//c_null_char, line, 256_c_int), 'fcp_report_line')
    do i = 1, 256
      if (line(i) == c_null_char) exit
    end do
    write(*,'(256a)') line(1:i-1)
    call get(FCP_F_WALLDIST, wallDistance, numCells)
  end subroutine

  ! ---- updateBoundary(phi)   src/finiteVolume/boundary/updateBoundary.f90 ------------------------------------------------------
  subroutine updateBoundary(phi)
    real(dp), dimension(numTotal), intent(inout) :: phi
    call put(FCP_F_S0, phi, numTotal)
    call fcp_check(fcp_update_boundary(ctx, FCP_F_S0), 'fcp_update_boundary')
    call get(FCP_F_S0, phi, numTotal)
  end subroutine

  ! ---- modify_viscosity_k_epsilon_rlzb()   TurbulenceModels/k_epsilon_rlzb.f90:38-50: calcsc_tke, calcsc_epsilon, modify_mu_eff ----
  ! (per-wall-face arrays visw, dnw, tau, ypl travel in the wall faces' boundary slots of numTotal-long fields; magStrain comes from
  !  calc_strain_and_vorticity on the device, the velocity gradients are the ones calcuvw left there)
  subroutine modify_viscosity_k_epsilon_rlzb()
    use TurbModelData, only: TurbModel
    type(fcp_scalar_params) :: prm
    type(fcp_report) :: rep
    real(c_double) :: fimin, fimax
    real(dp), allocatable :: wf(:)
    integer :: ib, i, iWall, isc
    call put(FCP_F_U, u, numTotal); call put(FCP_F_V, v, numTotal); call put(FCP_F_W, w, numTotal)
    call put(FCP_F_DEN, den, numTotal); call put(FCP_F_VIS, vis, numTotal); call put(FCP_F_FLMASS, flmass, numFaces)
    call put(FCP_F_TE, te, numTotal); call put(FCP_F_ED, ed, numTotal)
    ! modify_viscosity_turbulence.f90:28-33: velocity gradients of the corrected field, then strain and vorticity
    call fcp_check(fcp_grad(ctx, FCP_GRAD_GAUSS, FCP_F_U, FCP_F_DUDXI, 0_c_int), 'fcp_grad')
    call fcp_check(fcp_grad(ctx, FCP_GRAD_GAUSS, FCP_F_V, FCP_F_DVDXI, 0_c_int), 'fcp_grad')
    call fcp_check(fcp_grad(ctx, FCP_GRAD_GAUSS, FCP_F_W, FCP_F_DWDXI, 0_c_int), 'fcp_grad')
    allocate(wf(numTotal))
    wf = 0.0_dp; iWall = 0
    do ib = 1, numBoundaries
      if (bctype(ib) /= 'wall') cycle
      do i = 1, nfaces(ib)
        iWall = iWall + 1
        wf(iBndValueStart(ib) + i) = dnw(iWall)
      end do
    end do
    call put(FCP_F_DNW, wf, numTotal)
    wf = 0.0_dp; iWall = 0
    do ib = 1, numBoundaries
      if (bctype(ib) /= 'wall') cycle
      do i = 1, nfaces(ib)
        iWall = iWall + 1
        wf(iBndValueStart(ib) + i) = visw(iWall)
      end do
    end do
    call put(FCP_F_VISW, wf, numTotal)
    call fcp_check(fcp_calc_strain_and_vorticity(ctx), 'fcp_calc_strain_and_vorticity')
    do isc = 1, 2
      prm%kind = merge(FCP_SC_TKE_RLZB, FCP_SC_EPS_RLZB, isc == 1)
      prm%solver = solver_id(TurbModel%Scalar(isc)%lSolver); prm%maxiter = TurbModel%Scalar(isc)%maxiter
      prm%tol_abs = TurbModel%Scalar(isc)%tolAbs; prm%tol_rel = TurbModel%Scalar(isc)%tolRel
      prm%urf = TurbModel%Scalar(isc)%urf; prm%gds = TurbModel%Scalar(isc)%gds
      prm%cscheme = cscheme_id(TurbModel%Scalar(isc)%cScheme)
      prm%grad_method = grad_method_id(); prm%limiter = limiter_id()
      prm%tscheme = 0; prm%lowre = 0
      if (ltransient .and. (bdf .or. cn)) prm%tscheme = 1
      if (ltransient .and. bdf2) prm%tscheme = 2
      prm%timestep = timestep
      prm%prtr = merge(1.0_dp/1.0_dp, 1.0_dp/1.2_dp, isc == 1)      ! 1/sigma_k, 1/sigma_epsilon (k_epsilon_rlzb.f90:20-21)
      prm%viscos = viscos; prm%densit = densit
      if (isc == 1) then
        if (prm%tscheme >= 1) call put(FCP_F_PHIO, teo, numTotal)
        if (prm%tscheme >= 2) call put(FCP_F_PHIOO, teoo, numTotal)
        call fcp_check(fcp_calcsc(ctx, prm, FCP_F_TE, rep, fimin, fimax), 'fcp_calcsc')
        write(6,'(2x,es11.4,a,es11.4)') fimin, ' <= k <= ', fimax
      else
        if (prm%tscheme >= 1) call put(FCP_F_PHIO, edo, numTotal)
        if (prm%tscheme >= 2) call put(FCP_F_PHIOO, edoo, numTotal)
        call fcp_check(fcp_calcsc(ctx, prm, FCP_F_ED, rep, fimin, fimax), 'fcp_calcsc')
        write(6,'(2x,es11.4,a,es11.4)') fimin, ' <= epsilon <= ', fimax
      end if
    end do
    call fcp_check(fcp_modify_mu_eff_k_epsilon_rlzb(ctx, TurbModel%urfVis, viscos), 'fcp_modify_mu_eff_k_epsilon_rlzb')
    call get(FCP_F_TE, te, numTotal); call get(FCP_F_ED, ed, numTotal); call get(FCP_F_VIS, vis, numTotal)
    call get(FCP_F_GEN, gen, numCells)
    call get(FCP_F_DUDXI, dUdxi, 3*numTotal); call get(FCP_F_DVDXI, dVdxi, 3*numTotal); call get(FCP_F_DWDXI, dWdxi, 3*numTotal)
    call get(FCP_F_MAGSTRAIN, magStrain, numCells); call get(FCP_F_VORTICITY, vorticity, numCells)
    call get(FCP_F_VISW, wf, numTotal); iWall = 0
    do ib = 1, numBoundaries
      if (bctype(ib) /= 'wall') cycle
      do i = 1, nfaces(ib)
        iWall = iWall + 1
        visw(iWall) = wf(iBndValueStart(ib) + i)
      end do
    end do
    call get(FCP_F_YPL, wf, numTotal); iWall = 0
    do ib = 1, numBoundaries
      if (bctype(ib) /= 'wall') cycle
      do i = 1, nfaces(ib)
        iWall = iWall + 1
        ypl(iWall) = wf(iBndValueStart(ib) + i)
      end do
    end do
    call get(FCP_F_TAU, wf, numTotal); iWall = 0
    do ib = 1, numBoundaries
      if (bctype(ib) /= 'wall') cycle
      do i = 1, nfaces(ib)
        iWall = iWall + 1
        tau(iWall) = wf(iBndValueStart(ib) + i)
      end do
    end do
    deallocate(wf)
  end subroutine

  ! ---- modify_viscosity_k_omega_sst()   TurbulenceModels/k_omega_SST.f90:62-88: calcsc(TE,..,1), calcsc(ED,..,2), modify_mu_eff ------------
  ! (same data movement as modify_viscosity_k_epsilon_rlzb; additionally wallDistance per cell.  F1 (fsst) lives on the device between calls.)
  subroutine modify_viscosity_k_omega_sst()
    use TurbModelData, only: TurbModel
    use k_omega_SST, only: LowRe
    type(fcp_scalar_params) :: prm
    type(fcp_report) :: rep
    real(c_double) :: fimin, fimax
    real(dp), allocatable :: wf(:)
    integer :: ib, i, iWall, isc
    call put(FCP_F_U, u, numTotal); call put(FCP_F_V, v, numTotal); call put(FCP_F_W, w, numTotal)
    call put(FCP_F_DEN, den, numTotal); call put(FCP_F_VIS, vis, numTotal); call put(FCP_F_FLMASS, flmass, numFaces)
    call put(FCP_F_TE, te, numTotal); call put(FCP_F_ED, ed, numTotal); call put(FCP_F_WALLDIST, wallDistance, numCells)
    call fcp_check(fcp_grad(ctx, FCP_GRAD_GAUSS, FCP_F_U, FCP_F_DUDXI, 0_c_int), 'fcp_grad')      ! modify_viscosity_turbulence.f90:28-33
    call fcp_check(fcp_grad(ctx, FCP_GRAD_GAUSS, FCP_F_V, FCP_F_DVDXI, 0_c_int), 'fcp_grad')
    call fcp_check(fcp_grad(ctx, FCP_GRAD_GAUSS, FCP_F_W, FCP_F_DWDXI, 0_c_int), 'fcp_grad')
    call fcp_check(fcp_calc_strain_and_vorticity(ctx), 'fcp_calc_strain_and_vorticity')
    allocate(wf(numTotal))
    wf = 0.0_dp; iWall = 0
    do ib = 1, numBoundaries
      if (bctype(ib) /= 'wall') cycle
      do i = 1, nfaces(ib)
        iWall = iWall + 1
        wf(iBndValueStart(ib) + i) = dnw(iWall)
      end do
    end do
    call put(FCP_F_DNW, wf, numTotal)
    wf = 0.0_dp; iWall = 0
    do ib = 1, numBoundaries
      if (bctype(ib) /= 'wall') cycle
      do i = 1, nfaces(ib)
        iWall = iWall + 1
        wf(iBndValueStart(ib) + i) = visw(iWall)
      end do
    end do
    call put(FCP_F_VISW, wf, numTotal)
    do isc = 1, 2
      prm%kind = merge(FCP_SC_TKE_SST, FCP_SC_OMEGA_SST, isc == 1)
      prm%solver = solver_id(TurbModel%Scalar(isc)%lSolver); prm%maxiter = TurbModel%Scalar(isc)%maxiter
      prm%tol_abs = TurbModel%Scalar(isc)%tolAbs; prm%tol_rel = TurbModel%Scalar(isc)%tolRel
      prm%urf = TurbModel%Scalar(isc)%urf; prm%gds = TurbModel%Scalar(isc)%gds
      prm%cscheme = cscheme_id(TurbModel%Scalar(isc)%cScheme)
      prm%grad_method = grad_method_id(); prm%limiter = limiter_id()
      prm%tscheme = 0; prm%lowre = merge(1, 0, LowRe)
      if (ltransient .and. (bdf .or. cn)) prm%tscheme = 1
      if (ltransient .and. bdf2) prm%tscheme = 2
      prm%timestep = timestep; prm%prtr = 1.0_dp; prm%viscos = viscos; prm%densit = densit
      if (isc == 1) then
        if (prm%tscheme >= 1) call put(FCP_F_PHIO, teo, numTotal)
        if (prm%tscheme >= 2) call put(FCP_F_PHIOO, teoo, numTotal)
        call fcp_check(fcp_calcsc(ctx, prm, FCP_F_TE, rep, fimin, fimax), 'fcp_calcsc')
        write(*,'(2x,es11.4,a,es11.4)') fimin, ' <= k <= ', fimax
      else
        if (prm%tscheme >= 1) call put(FCP_F_PHIO, edo, numTotal)
        if (prm%tscheme >= 2) call put(FCP_F_PHIOO, edoo, numTotal)
        call fcp_check(fcp_calcsc(ctx, prm, FCP_F_ED, rep, fimin, fimax), 'fcp_calcsc')
        write(*,'(2x,es11.4,a,es11.4)') fimin, ' <= Omega <= ', fimax
      end if
    end do
    call fcp_check(fcp_modify_mu_eff_k_omega_sst(ctx, TurbModel%urfVis, viscos, densit, merge(1_c_int, 0_c_int, LowRe)), 'fcp_modify_mu_eff_k_omega_sst')
    call get(FCP_F_TE, te, numTotal); call get(FCP_F_ED, ed, numTotal); call get(FCP_F_VIS, vis, numTotal); call get(FCP_F_GEN, gen, numCells)
    call get(FCP_F_VISW, wf, numTotal); iWall = 0
    do ib = 1, numBoundaries
      if (bctype(ib) /= 'wall') cycle
      do i = 1, nfaces(ib)
        iWall = iWall + 1
        visw(iWall) = wf(iBndValueStart(ib) + i)
      end do
    end do
    deallocate(wf)
  end subroutine

  ! ---- modify_viscosity_wale_sgs() / modify_viscosity_vreman_sgs()   TurbulenceModels/wale_sgs.f90:33, vremanSGS.f90:33 -------------------
  subroutine modify_viscosity_sgs_device(model)
    use TurbModelData, only: TurbModel
    integer(c_int), intent(in) :: model              ! 0 = WALE, 1 = Vreman
    real(dp), allocatable :: wf(:)
    integer :: ib, i, iWall
    call put(FCP_F_U, u, numTotal); call put(FCP_F_V, v, numTotal); call put(FCP_F_W, w, numTotal)
    call put(FCP_F_DEN, den, numTotal); call put(FCP_F_VIS, vis, numTotal)
    call fcp_check(fcp_modify_viscosity_sgs(ctx, model, TurbModel%urfVis, viscos), 'fcp_modify_viscosity_sgs')
    call get(FCP_F_VIS, vis, numTotal)
    allocate(wf(numTotal))
    call get(FCP_F_VISW, wf, numTotal); iWall = 0
    do ib = 1, numBoundaries
      if (bctype(ib) /= 'wall') cycle
      do i = 1, nfaces(ib)
        iWall = iWall + 1
        visw(iWall) = wf(iBndValueStart(ib) + i)
      end do
    end do
    deallocate(wf)
    write(*,'(2x,es11.4,a,es11.4)') minval(vis/viscos), ' <= Viscosity ratio <= ', maxval(vis/viscos)
  end subroutine
  subroutine modify_viscosity_wale_sgs()
    call modify_viscosity_sgs_device(0_c_int)
  end subroutine
  subroutine modify_viscosity_vreman_sgs()
    call modify_viscosity_sgs_device(1_c_int)
  end subroutine

  ! ---- constant_mass_flow_forcing()   src/cappuccino/constant_mass_flow_forcing.f90 ------------------------------------------------
  subroutine constant_mass_flow_forcing_device()
    use parameters, only: magUbar, gradPcmf
    real(c_double) :: ustar
    call put(FCP_F_U, u, numTotal); call put(FCP_F_APU, apu, numCells)
    call fcp_check(fcp_constant_mass_flow_forcing(ctx, magUbar, gradPcmf, ustar), 'fcp_constant_mass_flow_forcing')
    call get(FCP_F_U, u, numTotal)
    write(6,'(2(a,es13.6))') "  Uncorrected Ubar = ", ustar, " pressure gradient = ", gradPcmf
  end subroutine

  ! ---- src-par: exchange(phi), global_sum(phi)   src-par/exchange.f90:3, src-par/global_sum_mpi.f90:4 ---------------------------
  subroutine exchange(phi)
    real(dp), dimension(numTotal), intent(inout) :: phi
    call put(FCP_F_S0, phi, numTotal)
    call fcp_check(fcp_exchange(ctx, FCP_F_S0), 'fcp_exchange')
    call get(FCP_F_S0, phi, numTotal)
  end subroutine

  subroutine global_sum(phi)
    real(dp), intent(inout) :: phi
    call fcp_check(fcp_global_sum(ctx, phi), 'fcp_global_sum')
  end subroutine

  subroutine global_max(phi)                 ! src-par/global_max_mpi.f90
    real(dp), intent(inout) :: phi
    call fcp_check(fcp_global_max(ctx, phi), 'fcp_global_max')
  end subroutine

  subroutine global_min(phi)                 ! src-par/global_min_mpi.f90
    real(dp), intent(inout) :: phi
    call fcp_check(fcp_global_min(ctx, phi), 'fcp_global_min')
  end subroutine

  subroutine global_isum(i)                  ! src-par/global_isum_mpi.f90
    integer, intent(inout) :: i
    integer(c_int64_t) :: v
    v = int(i, c_int64_t)
    call fcp_check(fcp_global_isum(ctx, v), 'fcp_global_isum')
    i = int(v)
  end subroutine

end module fcp_backend
