/*
 * include/fcp.h -- C-ABI of libfcp_b200.so, the B200-native pressure-velocity coupling hot path
 * of freeCappuccino (assembly into CSR, Gauss / least-squares gradients, DPCG / ICCG / BiCGStab).
 *
 * The reference (nikola-m/freeCappuccino-dev, pure Fortran) has no FFI layer; its boundary is the
 * Fortran module-procedure level.  Each entry point below names the reference
 * procedure (file:line, relative to the reference root) it replaces.  The Fortran binding
 * (interface ... bind(C)) is in fortran/fcp_b200.f90, the C++ host mirror in host/fcp_host.hpp.
 *
 * Conventions (those of the Fortran host):
 *   - every index array crossing this boundary is 1-based int32 (default INTEGER),
 *   - reals are IEEE binary64 (real(dp)),
 *   - gradients are (3,numTotal) column-major: x,y,z interleaved per cell,
 *   - fields have length numTotal = numCells + numBoundaryFaces; the value of boundary face
 *     `iface` lives at numCells + (iface - numInnerFaces)       (src/mesh/geometry.f90:282-290),
 *   - all pointers are HOST pointers; the library owns the device copies,
 *   - every function returns 0 on success, a negative FCP_E* code otherwise; the reference has no
 *     status codes (fatal conditions `stop`), so the Fortran shim turns non-zero into `stop`.
 *   - there is NO CPU fallback: if no sm_100 device is usable fcp_ctx_create fails with
 *     FCP_ENODEVICE.
 * Threading: like the reference (module-level workspaces), a context is not re-entrant.
 */
#ifndef FCP_H
#define FCP_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCP_OK         0
#define FCP_EINVAL    -1   /* bad argument */
#define FCP_ECUDA     -2   /* CUDA runtime error (fcp_last_error() has the text) */
#define FCP_ENODEVICE -3   /* no usable CUDA device */
#define FCP_ENCCL     -4   /* NCCL error or NCCL library not loadable */
#define FCP_ESTATE    -5   /* call sequence error (e.g. LSQ matrix not created) */

/* patch types: character(len=30) bctype(ib) of src/mesh/geometry.f90:62-70 (+ `process`, src-par/geometry.f90:218-240) */
enum { FCP_BC_WALL = 0, FCP_BC_INLET = 1, FCP_BC_OUTLET = 2, FCP_BC_SYMMETRY = 3,
       FCP_BC_PRESSURE = 4, FCP_BC_PERIODIC = 5, FCP_BC_EMPTY = 6, FCP_BC_PROCESS = 7 };

/* linear solvers: the strings of csrsolve, src/linearSolvers/linear_solvers.f90:40-91 */
enum { FCP_SOLVER_DPCG = 1, FCP_SOLVER_ICCG = 2, FCP_SOLVER_BICGSTAB = 3,
       FCP_SOLVER_GAUSS_SEIDEL = 4 /* 'gauss-seidel', linear_solvers.f90:96-201; single GPU only (src-par has no such solver) */ };

/* gradient methods: the logicals lstsq / lstsq_dm / (default) gauss of gradients.f90:118-138 */
enum { FCP_GRAD_GAUSS = 0, FCP_GRAD_LSQ = 1, FCP_GRAD_LSQ_DM = 2, FCP_GRAD_LSQ_QR = 3 };

/* option_limiter of grad_scalar_field_w_option, gradients.f90:258-276: 'Barth-Jespersen' :288, 'Venkatakrishnan' :378,
 * 'R3' :464 (the formula the reference has active under that name is R4), 'multidimensional' :556 */
enum { FCP_LIMITER_NONE = 0, FCP_LIMITER_BARTH_JESPERSEN = 1, FCP_LIMITER_VENKATAKRISHNAN = 2, FCP_LIMITER_R3 = 3,
       FCP_LIMITER_MULTIDIMENSIONAL = 4 };

/* pscheme of Pressure/nablap.f90:51-112 */
enum { FCP_PSCHEME_LINEAR = 0, FCP_PSCHEME_CENTRAL = 1, FCP_PSCHEME_WEIGHTED = 2 };

/* device-resident fields = the reference's module arrays (variables.f90, sparse_matrix.f90:20-40).
 * Scalars have numTotal entries, gradients 3*numTotal, FCP_F_A nnz, FCP_F_FLMASS numFaces,
 * FCP_F_APR npro (src-par/sparse_matrix.f90:25). */
enum {
  FCP_F_U = 0, FCP_F_V, FCP_F_W, FCP_F_P, FCP_F_PP, FCP_F_DEN, FCP_F_VIS,
  FCP_F_APU, FCP_F_APV, FCP_F_APW, FCP_F_SU, FCP_F_SV, FCP_F_SW,
  FCP_F_S0, FCP_F_S1, FCP_F_S2, FCP_F_S3,          /* user scalars (phi, mu, rhs ... of laplacian / grad / csrsolve) */
  FCP_F_DUDXI, FCP_F_DVDXI, FCP_F_DWDXI, FCP_F_DPDXI, FCP_F_G0, FCP_F_G1,   /* (3,numTotal) */
  FCP_F_FLMASS, FCP_F_A, FCP_F_APR,
  FCP_F_H,                                         /* h(nnz): momentum coefficients saved by calcp_piso (calcp_piso.f90:81) */
  FCP_F_RU, FCP_F_RV, FCP_F_RW,                    /* rU,rV,rW: momentum right-hand sides (velocity.f90:567) */
  FCP_F_VISW,                                      /* visw(iWall): effective wall viscosity, stored in the wall faces' boundary slots */
  FCP_F_UO, FCP_F_VO, FCP_F_WO, FCP_F_UOO, FCP_F_VOO, FCP_F_WOO, FCP_F_UOOO, FCP_F_VOOO, FCP_F_WOOO,   /* past time levels */
  FCP_F_SPU, FCP_F_SPV, FCP_F_SP,                  /* spu, spv, sp: implicit source terms of the three momentum equations */
  FCP_F_TE, FCP_F_ED,                              /* te, ed: turbulence kinetic energy and its dissipation rate */
  FCP_F_PHIO, FCP_F_PHIOO,                         /* past time levels of the scalar fcp_calcsc is solving for (teo/teoo, edo/edoo ...) */
  FCP_F_GEN, FCP_F_MAGSTRAIN, FCP_F_VORTICITY,     /* gen, magStrain, vorticity (variables module) */
  FCP_F_DNW, FCP_F_TAU, FCP_F_YPL,                 /* dnw, tau, ypl(iWall): per wall face, stored in the wall faces' boundary slots like visw above */
  FCP_F_SCTMP,                                     /* scratch of fcp_calcsc */
  FCP_F_FSST, FCP_F_WALLDIST,                      /* k-omega SST: blending function F1 (fsst), wallDistance per cell */
  FCP_F_DTEDXI, FCP_F_DEDDXI,                      /* (3,numTotal): gradients of k and omega kept by the SST pair (cross diffusion) */
  FCP_F_COUNT
};

typedef struct fcp_ctx fcp_ctx;        /* one mesh (partition) on one GPU */
typedef struct fcp_solver fcp_solver;  /* one CSR pattern on one GPU, no mesh (explicit-CSR signature) */

/* mirror of the arrays of the `geometry` module, src/mesh/geometry.f90:12-86 */
typedef struct {
  int32_t numCells, numInnerFaces, numBoundaryFaces, numBoundaries;
  const int32_t *owner;       /* [numInnerFaces+numBoundaryFaces] */
  const int32_t *neighbour;   /* [numInnerFaces] */
  const double *arx, *ary, *arz, *xf, *yf, *zf;   /* [numFaces] */
  const double *facint, *Df;                      /* [numInnerFaces]  geometry.f90:581-606, 648-664 */
  const double *xc, *yc, *zc, *vol;               /* [numCells] */
  const int32_t *bctype;      /* [numBoundaries] FCP_BC_* */
  const int32_t *nfaces;      /* [numBoundaries] */
  const int32_t *startFace;   /* [numBoundaries] 0-based offset of the patch's first face (boundary file, Appendix D) */
  const int32_t *startFaceTwin; /* [numBoundaries] or NULL: for FCP_BC_PERIODIC patches the startFace of the twin patch (listed as
                               * FCP_BC_EMPTY, same face count, faces already paired by face_mapping, geometry.f90:82,251-257,1848-1997);
                               * ignored for the other patch types.  NULL = the mesh has no periodic patch. */
  const double *DfPeriodic;   /* [numBoundaryFaces] or NULL.  The reference reads `Df(i)` for periodic face i of a patch, i.e. the Df of INNER face
                               * number i (quirk Q21, calcp_simple.f90:199).  NULL: the library does the same with this mesh's own inner faces.  A
                               * partitioned run that must reproduce the unpartitioned reference passes here, per boundary face, the value the
                               * GLOBAL mesh would have used (the partitioner knows it); entries of non-periodic faces are ignored. */
} fcp_mesh_desc;

/* the numbers the reference prints in its solver report line, linear_solvers.f90:354-355,540-541,781-782 */
typedef struct {
  double res0;      /* initial L1 residual sum|b - A x|                        (:264) */
  double resl;      /* final L1 residual                                        (:327) */
  double factor;    /* sum|a_ii x_i| + small after the first update             (:336) */
  double resor;     /* res0/factor: the value returned in the resor/res0 dummy  (:340) */
  int32_t iters;    /* itr_used */
  int32_t solver;   /* FCP_SOLVER_* */
} fcp_report;

/* ---- library ------------------------------------------------------------------------------ */
int fcp_version(void);
const char *fcp_last_error(void);
/* number of kernels this library has launched so far in this process (bench.py's gpu_launches) */
int64_t fcp_launch_count(void);

/* ---- context: mesh upload + create_CSR_matrix (src/sparseMatrix/sparse_matrix.f90:86-296) ---- */
int fcp_ctx_create(const fcp_mesh_desc *mesh, int device, fcp_ctx **out);
int fcp_ctx_destroy(fcp_ctx *ctx);
int fcp_ctx_sizes(const fcp_ctx *ctx, int32_t *numCells, int32_t *numTotal, int32_t *numFaces, int32_t *nnz, int32_t *npro);
/* ia(numCells+1), ja(nnz), diag(numCells), icell_jcell_csr_index(numInnerFaces+numPeriodic), jcell_icell_csr_index(same);
 * any pointer may be NULL (sparse_matrix.f90:183-202, 251-293).  nnz = numCells + 2 (numInnerFaces + numPeriodic): every
 * periodic face adds the twin entries (owner(face), owner(twin face)) and its transpose (:141-171); a periodic pair that joins
 * two cells which already share an inner face is refused (the reference would silently alias the two coefficients). */
int fcp_csr_pattern(const fcp_ctx *ctx, int32_t *ia, int32_t *ja, int32_t *diag, int32_t *icell_jcell, int32_t *jcell_icell);
int fcp_sync(fcp_ctx *ctx);

/* ---- fields ------------------------------------------------------------------------------- */
int fcp_field_upload(fcp_ctx *ctx, int field, const double *host, int64_t count);     /* first `count` entries */
int fcp_field_download(fcp_ctx *ctx, int field, double *host, int64_t count);
int fcp_field_fill(fcp_ctx *ctx, int field, double value);
int fcp_field_copy(fcp_ctx *ctx, int dst_field, int src_field);
/* dst = alpha*x + beta*y over fields of equal extent: the arithmetic of the fvEquation operators (+), (-), (==),
 * fvImplicit/fvEquation.f90:158-404 (coef(nnz) and source(numCells) added / subtracted element by element) */
int fcp_field_axpby(fcp_ctx *ctx, int dst_field, double alpha, int x_field, double beta, int y_field);
/* raw device address of a field (for zero-copy interop with other CUDA code in the same process) */
int fcp_field_devptr(fcp_ctx *ctx, int field, void **devptr, int64_t *count);

/* ---- operators ---------------------------------------------------------------------------- */
/* y = A x with the context matrix FCP_F_A  (the SpMV loops linear_solvers.f90:256-261, 308-313) */
int fcp_spmv(fcp_ctx *ctx, int x_field, int y_field);
/* csrsolve(solver, fi, rhs, res0, itr_max, tol_abs, tol_rel, chvar)   linear_solvers.f90:40-91;
 * dpcg :206-359, iccg :364-545, bicgstab :548-786.  fi and rhs are device fields. */
int fcp_csrsolve(fcp_ctx *ctx, int solver, int fi_field, int rhs_field, int32_t itr_max,
                 double tol_abs, double tol_rel, fcp_report *rep);
/* the reference's report line for `rep` (same text, parsed by examples/ * /plotResiduals) */
int fcp_report_line(const fcp_report *rep, const char *chvar, char *buf, int buflen);

/* grad(phi,dPhidxi): grad_gauss gradients.f90:1607-1693, grad_lsq :782-893, grad_lsq_dm :1334-1486, grad_lsq_qr :1057-1152.
 * lsq_row2_reference != 0 reproduces the reference's back-substitution row 2.
 * FCP_GRAD_LSQ_QR (create_matrix_lsq_qr :900-1052, thin QR by modified Gram-Schmidt misc/matrix.f90:366-419) holds at
 * most m = 6 faces per cell (:924) and fails with FCP_EINVAL on any other mesh; see DESIGN.md quirk Q20. */
int fcp_create_lsq_grad_matrix(fcp_ctx *ctx, int method);     /* gradients.f90:72-101, 660-779, 1157-1326 */
int fcp_grad(fcp_ctx *ctx, int method, int phi_field, int grad_field, int lsq_row2_reference);
/* the limiter alone, applied in place to grad_field; BJ/Venkatakrishnan/R3 use the GLOBAL extrema of phi(1:numCells)
 * (gradients.f90:317-318; all ranks in a multi-GPU run) */
int fcp_slope_limiter(fcp_ctx *ctx, int limiter, int phi_field, int grad_field);
/* grad(phi,dPhidxi,option,option_limiter)   gradients.f90:217-278: dPhidxi = 0, gradient by `method`, then the limiter */
int fcp_grad_opt(fcp_ctx *ctx, int method, int limiter, int phi_field, int grad_field);
/* laplacian(mu,phi): fills FCP_F_A, accumulates into FCP_F_SU   src/finiteVolume/fvImplicit/laplacian.f90 */
int fcp_laplacian(fcp_ctx *ctx, int mu_field, int phi_field);
/* gradp_and_sources(p): fills FCP_F_SU/SV/SW and FCP_F_DPDXI, extrapolates p to boundaries   Pressure/nablap.f90:19-208 */
int fcp_gradp_and_sources(fcp_ctx *ctx, int pscheme, int p_field);
/* inner-face + patch assembly of the pressure-correction equation   Pressure/calcp_simple.f90:69-234,
 * fluxes/faceflux_mass.f90:175-249 (facefluxmass2), :765-831, :833-916 */
int fcp_assemble_pcorr_simple(fcp_ctx *ctx, int const_mflux, double flomas);
/* Which mass-flux routine the INNER faces of that assembly use (the two trees are switchable where they differ, quirk Q10).
 * variant 0 (default): facefluxmass2 of the serial tree (calcp_simple.f90:90, faceflux_mass.f90:175-249); variant 1: facefluxmass of the MPI tree
 * (src-par/calcp_simple.f90:40-79, src-par/faceflux_mass.f90:28-180: face_value_central velocities from grad(U), grad(V), grad(W) computed with
 * `grad_method` = FCP_GRAD_* into FCP_F_DUDXI/DVDXI/DWDXI, per-component (Vol/Ap)_f from apu, apv, apw, the P'/E' pressure correction).  Process faces
 * keep facefluxmass2 in both (src-par/calcp_simple.f90:96-118).  Affects fcp_assemble_pcorr_simple and fcp_calcp_simple; calcp_piso has no MPI twin. */
int fcp_set_flux_variant(fcp_ctx *ctx, int variant, int grad_method);
/* flux / velocity / pressure correction after the solve   calcp_simple.f90:331-429 */
int fcp_correct_simple(fcp_ctx *ctx, int pscheme, double urfp, int32_t pRefCell);
/* non-orthogonal corrector   calcp_simple.f90:433-455 + fluxmc2 faceflux_mass.f90:650-696 */
int fcp_nonorth_corrector(fcp_ctx *ctx);

typedef struct {
  int32_t solver;       /* lSolverP  */
  int32_t maxiter;      /* maxiterP  */
  double tol_abs, tol_rel;   /* tolAbsP, tolRelP   pressure.f90:28-33 */
  double urfp;          /* urfP */
  int32_t npcor;        /* number of pressure-correction passes (non-orthogonal correctors = npcor-1) */
  int32_t pRefCell;     /* 1-based */
  int32_t pscheme;      /* FCP_PSCHEME_* */
  int32_t const_mflux;  /* skip adjustMassFlow */
  double flomas;        /* inlet mass flow used by adjustMassFlow */
  int32_t zero_pp;      /* 0: warm start like the serial tree; 1: pp=0 like src-par/calcp_simple.f90:170 (quirk Q8) */
} fcp_simple_params;
/* one whole calcp_simple   Pressure/calcp_simple.f90:1-470 ; rep[ipcorr] for ipcorr < npcor */
int fcp_calcp_simple(fcp_ctx *ctx, const fcp_simple_params *prm, fcp_report *rep);

typedef struct {
  int32_t solver;       /* lSolverP */
  int32_t maxiter;      /* maxiterP */
  double tol_abs, tol_rel;
  double urfp;          /* urfP (calcp_piso.f90:333) */
  int32_t ncorr;        /* PISO correctors (:84) */
  int32_t npcor;        /* non-orthogonal passes per corrector (:308) */
  int32_t pscheme;      /* FCP_PSCHEME_* of the closing gradp_and_sources (:425) */
  int32_t const_mflux;  /* skip adjustMassFlow (:186) */
  double flomas;
} fcp_piso_params;
/* one whole calcp_piso   Pressure/calcp_piso.f90:81-489 (+ facefluxmass_piso faceflux_mass.f90:389-459, fluxmc :564-647).
 * On entry FCP_F_A holds the momentum coefficients (copied to FCP_F_H like `h = a`), FCP_F_RU/RV/RW the momentum
 * right-hand sides, FCP_F_APU/APV/APW the reciprocal diagonals.  rep[(icorr-1)*npcor + ipcorr-1].
 * Periodic patches: :248-293 (facefluxmass2_periodic) and :435-460; H(U) runs over the inner faces only, as in the reference. */
int fcp_calcp_piso(fcp_ctx *ctx, const fcp_piso_params *prm, fcp_report *rep);

/* constant_mass_flow_forcing   src/cappuccino/constant_mass_flow_forcing.f90 (the caller runs it after calcp_simple / calcp_piso when
 * const_mflux, calcp_simple.f90:468, calcp_piso.f90:492): magUbarStar = volume-weighted mean of U, rUAw = that of APU,
 * gragPplus = (magUbar - magUbarStar)/rUAw; U(1:numCells) += APU*gragPplus; *gradPcmf += gragPplus.  The two sums follow the
 * library's fixed reduction tree (the reference adds cell by cell), so the result agrees with the reference to rounding. */
int fcp_constant_mass_flow_forcing(fcp_ctx *ctx, double magUbar, double *gradPcmf, double *magUbarStar /* may be NULL */);
/* updateBoundary(phi)   src/finiteVolume/boundary/updateBoundary.f90: outlet / symmetry / pressure / empty patches copy the owner
 * value into the boundary slot, a periodic face and its twin take the mean of the two cells across the pair */
int fcp_update_boundary(fcp_ctx *ctx, int field);

/* cSchemeU of face_value, interpolation.f90:28-113 and :596-640 (flux limiters in source order) */
enum { FCP_CS_CDS = 0, FCP_CS_CENTRAL, FCP_CS_LINEAR_UPWIND, FCP_CS_KAPPA, FCP_CS_MUSCL, FCP_CS_UMIST, FCP_CS_KOREN, FCP_CS_SMART,
       FCP_CS_AVL_SMART, FCP_CS_CHARM, FCP_CS_VANLEER, FCP_CS_OSPRE, FCP_CS_MINMOD, FCP_CS_BOUNDED_LINEAR_UPWIND,
       FCP_CS_BOUNDED_LINEAR_UPWIND02, FCP_CS_BOUNDED_CENTRAL, FCP_CS_FROMM, FCP_CS_CUI, FCP_CS_QUICK, FCP_CS_SPL13, FCP_CS_COUNT };
typedef struct {
  int32_t solver, maxiter;      /* lSolverU, maxiterU   velocity.f90:30-31 */
  double tol_abs, tol_rel;      /* tolAbsU, tolRelU */
  double urf[3];                /* urfU(1:3) */
  double gds;                   /* gdsU */
  int32_t cscheme;              /* FCP_CS_* */
  int32_t grad_method;          /* FCP_GRAD_* used by grad(U,dUdxi) (:171-173) */
  int32_t limiter;              /* FCP_LIMITER_* applied by grad (gradients.f90:140-160) */
  int32_t pscheme;              /* FCP_PSCHEME_* of gradp_and_sources(p) (:177) */
  int32_t tscheme;              /* 0 steady, 1 bdf, 2 bdf2, 3 bdf3 (:217-281); Crank-Nicolson is not built */
  int32_t piso;                 /* keep rU,rV,rW for calcp_piso (:564-568) */
  double timestep;
  int32_t const_mflux, pad;
  double gradPcmf;              /* constant-mass-flow forcing (:189-191) */
  double viscos;                /* molecular viscosity: wall faces use max(viscos, visw) (:443) */
} fcp_uvw_params;
/* calcuvw: the momentum predictor   Velocity/velocity.f90:50-750 (tier "next" row f1): updateVelocityAtBoundary,
 * grad(U/V/W), gradp_and_sources(p), volume + face + boundary terms into FCP_F_A / SU,SV,SW / SPU,SPV,SP, then the
 * three under-relaxed solves; leaves apu,apv,apw (and rU,rV,rW when piso) for calcp_simple / calcp_piso.
 * Periodic patches: :393-432 + facefluxuvw_periodic :1038-1180.  Not built: Crank-Nicolson, buoyancy, MHD. rep[0..2] = U, V, W. */
int fcp_calcuvw(fcp_ctx *ctx, const fcp_uvw_params *prm, fcp_report *rep);

/* ---- row f4: the scalar transport template (calcsc) of the turbulence models --------------------------------------------------
 * fluxes/scalar_fluxes.f90:32-343 (facefluxsc, facefluxsc_periodic, facefluxsc_boundary) inside the assembly of
 * TurbulenceModels/k_epsilon_rlzb.f90: grad(phi) -> volume sources + bdf/bdf2 term -> inner faces -> inlet/outlet/pressure, periodic and
 * wall patches -> a(diag) = sp - sum(off-diagonals), under-relaxation -> csrsolve -> updateBoundary -> min/max, clip to `small` when negative.
 *   FCP_SC_GENERIC   volume sources su, sp read from FCP_F_S2, FCP_F_S3; wall faces add nothing (zero flux); no clipping
 *   FCP_SC_TKE_RLZB  calcsc_tke :52-445 (phi_field must be FCP_F_TE): gen = |vis - viscos| S^2 (FCP_F_GEN out), wall cells use the production
 *                    from the wall shear stress (FCP_F_TAU out; reads FCP_F_VISW, FCP_F_DNW, U, V, W)
 *   FCP_SC_EPS_RLZB  calcsc_epsilon :447-790 (phi_field must be FCP_F_ED): realizable c1; wall cells: row cleared, ed = cmu75 k^1.5/(cappa dnw)
 * Inputs: FCP_F_DEN, FCP_F_VIS (effective viscosity, boundary slots included), FCP_F_FLMASS, FCP_F_MAGSTRAIN, FCP_F_TE, FCP_F_ED, FCP_F_PHIO/PHIOO.
 * Outputs: the scalar, FCP_F_A (its matrix), FCP_F_SU, FCP_F_SP, FCP_F_G0 (its gradient).  k^1.5 uses the device pow(): that value agrees
 * with the reference's libm to rounding, everything else follows the reference's operation order.  Not built: Crank-Nicolson, buoyancy.
 * Partitioned meshes: supported except for the SST pair (FCP_ESTATE with a communicator). */
/*   FCP_SC_TKE_SST / FCP_SC_OMEGA_SST   the k and omega equations of TurbulenceModels/k_omega_SST.f90:91-788 (omega lives in FCP_F_ED as in the
 *                    reference): production limiter, F1 = tanh(ksi^4) (FCP_F_FSST, written by the omega call, read by both -- the k call uses the
 *                    previous one, zero on the first call), cross diffusion from FCP_F_DTEDXI . FCP_F_DEDDXI, sigma from the OWNER cell of a face,
 *                    omega imposed in wall cells, extrema and clip over the whole array; needs FCP_F_WALLDIST; `lowre` = the module's LowRe switch */
enum { FCP_SC_GENERIC = 0, FCP_SC_TKE_RLZB = 1, FCP_SC_EPS_RLZB = 2, FCP_SC_TKE_SST = 3, FCP_SC_OMEGA_SST = 4 };
typedef struct {
  int32_t kind;                 /* FCP_SC_* */
  int32_t solver, maxiter;      /* TurbModel%Scalar(i)%lSolver, %maxiter */
  int32_t cscheme;              /* FCP_CS_* (TurbModel%Scalar(i)%cScheme) */
  int32_t grad_method, limiter; /* the configuration of grad(phi, dPhidxi) */
  int32_t tscheme;              /* 0 steady, 1 bdf, 2 bdf2 */
  int32_t lowre;                /* k-omega SST only: LowRe (k_omega_SST.f90:19) */
  double tol_abs, tol_rel, urf, gds, timestep;
  double prtr;                  /* 1/sigma of the scalar */
  double viscos, densit;
} fcp_scalar_params;
int fcp_calcsc(fcp_ctx *ctx, const fcp_scalar_params *prm, int phi_field, fcp_report *rep, double *fimin, double *fimax);
/* wall_distance (src/mesh/wall_distance.f90:75-133): the Poisson-equation wall distance -- laplacian(1, phi) (every patch Dirichlet, as the reference's
 * laplacian does), q = -vol, csrsolve('iccg', 500, 1e-12, 1e-10), owner values into the boundary slots of every patch but 'wall', grad_gauss,
 * d = -|grad phi| + sqrt(|grad phi|^2 + 2 phi) -> FCP_F_WALLDIST.  Scratch: FCP_F_S0 (phi), FCP_F_S1, FCP_F_SU, FCP_F_G0, FCP_F_A.  rep may be NULL. */
int fcp_wall_distance(fcp_ctx *ctx, fcp_report *rep);
/* calc_strain_and_vorticity (fvExplicit/calc_strain_and_vorticity.f90): FCP_F_DUDXI/DVDXI/DWDXI -> FCP_F_MAGSTRAIN, FCP_F_VORTICITY */
int fcp_calc_strain_and_vorticity(fcp_ctx *ctx);
/* modify_mu_eff of the realizable k-epsilon model (k_epsilon_rlzb.f90:792-975): effective viscosity from te, ed and the velocity gradients,
 * under-relaxed by urfVis; updateBoundary(vis); wall functions -> FCP_F_VISW, FCP_F_YPL, FCP_F_TAU and vis at the wall faces.  acos, cos and
 * log are the device's: agreement with the reference's libm is to rounding. */
int fcp_modify_mu_eff_k_epsilon_rlzb(fcp_ctx *ctx, double urfVis, double viscos);
/* modify_mu_eff of the k-omega SST model (k_omega_SST.f90:790-958): F2 = tanh(etha^2), vis = viscos + den a1 k / max(a1 omega, S F2) (LowRe variant),
 * updateBoundary(vis), automatic wall treatment (u+ blended from the viscous and the log law) -> FCP_F_VISW, FCP_F_YPL, FCP_F_TAU.  tanh and log are
 * the device's. */
int fcp_modify_mu_eff_k_omega_sst(fcp_ctx *ctx, double urfVis, double viscos, double densit, int lowre);

/* fvxGradient's Gauss gradient (fvExplicit/fvxGradient.f90:1549-1662, gradco :1761-1817): two passes, the second interpolates the face value
 * with the first pass's gradient (skewness correction).  It is what Grad(U) of the tensor-field layer returns with its default flags; note that
 * it is NOT the grad_gauss of gradients.f90 (FCP_GRAD_GAUSS).  Scratch: the G1 gradient field (G0 when grad_field is G1). */
int fcp_grad_gauss_fvx(fcp_ctx *ctx, int phi_field, int grad_field);
/* grad_gauss of the MPI tree (src-par/gradients.f90:1547-1664, gradco :1775-1837): `nigrad` passes, pass k interpolates the face value with the
 * gradient of pass k-1 (zero in the first pass); the serial tree's one-pass grad_gauss is FCP_GRAD_GAUSS, whose face value P + (N - P) lambda
 * rounds differently from gradco's P fxp + N fxn even for nigrad = 1.  SURVEY 0.1: the two trees are switchable where they differ.  Same scratch
 * field as fcp_grad_gauss_fvx, which is the nigrad = 2 case. */
int fcp_grad_gauss_iter(fcp_ctx *ctx, int phi_field, int grad_field, int nigrad);
/* modify_viscosity_wale_sgs (TurbulenceModels/wale_sgs.f90:33-185) / modify_viscosity_vreman_sgs (vremanSGS.f90:33-179): D = Grad(U) with the
 * gradient above (left in FCP_F_DUDXI/DVDXI/DWDXI), the model's tensor algebra (tensorFields.f90, quirk Q24 of its inner product reproduced),
 * vis = urfVis (mu_sgs + viscos) + (1 - urfVis) vis, then the boundary values: wall faces FCP_F_VISW = vis = max(viscos, 0), periodic pairs the
 * mean of the two cells, every other patch the owner value.  pow() is the device's: agreement with the reference's libm is to rounding. */
enum { FCP_SGS_WALE = 0, FCP_SGS_VREMAN = 1 };
int fcp_modify_viscosity_sgs(fcp_ctx *ctx, int model, double urfVis, double viscos);

/* ---- explicit-CSR solver signature: dpcg|iccg|bicgstab(n,nnz,ia,ja,a,diag,fi,rhs,...) -------- */
/* linear_solvers.f90:206, :364, :548 ; pattern analysed once, values per solve */
int fcp_solver_create(int32_t n, int32_t nnz, const int32_t *ia, const int32_t *ja, const int32_t *diag,
                      int device, fcp_solver **out);
int fcp_solver_destroy(fcp_solver *s);
int fcp_solver_solve(fcp_solver *s, int solver, const double *a, double *fi, const double *rhs,
                     int32_t itr_max, double tol_abs, double tol_rel, fcp_report *rep);

/* ---- multi-GPU: src-par/exchange.f90:3-129, src-par/global_sum_mpi.f90:4-37 ----------------- */
/* NCCL bootstrap: rank 0 calls fcp_comm_unique_id, the host distributes the 128 bytes, all ranks call fcp_comm_init.
 * peer_rank[ib] = rank on the other side of `process` patch ib (-1 otherwise); both sides list the
 * shared faces in the same order (src-par/geometry.f90:218-240). */
int fcp_comm_unique_id(void *id128);
int fcp_comm_init(fcp_ctx *ctx, int rank, int nranks, const void *id128, const int32_t *peer_rank);
/* HOST-ONLY (no GPU needed): the halo layout fcp_comm_init derives from the patch table -- per process patch its peer,
 * offset and face count in the halo buffers (patch order); per process face its owner cell and ghost slot (0-based field
 * indices); the faces grouped by the 2048-row chunk of their owner cell (chunk_ptr / chunk_face), the launch order of
 * the chunks (bit 31 = owns process faces) and the boundary-face -> face-ordinal map.  Any pointer may be NULL.  Used by
 * the CPU tests to validate the layout for ranks with several neighbours (src-par/my_mpi_module.f90:11-17 analogue). */
int fcp_comm_plan(const fcp_mesh_desc *mesh, const int32_t *peer_rank, int rank, int nranks, int32_t *npatch, int32_t *patch_peer,
                  int32_t *patch_off, int32_t *patch_cnt, int32_t *cell, int32_t *slot, int32_t *chunk_ptr, int32_t *chunk_face,
                  int32_t *chunk_order, int32_t *ghost_ord);
/* 1: halo values and reduction partials travel as peer-memory stores over NVLink (CUDA IPC windows, fused into the
 * Krylov kernels); 0: NCCL send/recv + all-gather (FCP_COMM=nccl or peer mapping unavailable); -1: no communicator */
int fcp_comm_mode(const fcp_ctx *ctx);
/* Interpolation factors of the `process` faces (src-par/geometry.f90:822-868 `fpro`, patch order, the weight of the GHOST cell seen from this rank).
 * fcp_comm_init computes them from the ghost cell centres with the serial tree's formula (geometry.f90:581-606); a host that follows the MPI tree's
 * line-plane variant (quirk Q9) or reads them from a file overrides them here, after fcp_comm_init.  count must equal the number of process faces. */
int fcp_set_process_facint(fcp_ctx *ctx, const double *fpro, int32_t count);
/* Orientation of the `process` faces in the UNPARTITIONED mesh (patch order): flipped[i] != 0 when the cell of this rank is the face's NEIGHBOUR there
 * (its owner lives on the peer).  A rank always sees itself as the owner of its process faces (src-par layout), and nearly every face formula of
 * the path is symmetric in that choice up to rounding -- the exception is the k-omega SST pair, which takes the 1/sigma of a face from the blending
 * function of the face's OWNER cell (k_omega_SST.f90:440-449): fcp_calcsc(FCP_SC_TKE_SST / FCP_SC_OMEGA_SST) on a partition needs this call (after
 * fcp_comm_init; the partitioner knows the orientation) and fails with FCP_ESTATE without it.  count must equal the number of process faces. */
int fcp_set_process_orientation(fcp_ctx *ctx, const int32_t *flipped, int32_t count);
int fcp_exchange(fcp_ctx *ctx, int field);                    /* ghost slots of `process` patches <- owner values on the peer */
int fcp_global_sum(fcp_ctx *ctx, double *value);              /* in place, all ranks */
int fcp_global_max(fcp_ctx *ctx, double *value);
int fcp_global_min(fcp_ctx *ctx, double *value);
int fcp_global_isum(fcp_ctx *ctx, int64_t *value);           /* src-par/global_isum_mpi.f90 (nnz, cell counts at set-up): exact, in place, all ranks */

/* ---- per-kernel-class device timing (CUDA events on the context stream, around every launch of the class) ---- */
enum { FCP_K_SPMV_DOT = 0, FCP_K_CG_PK, FCP_K_CG_UPDATE, FCP_K_CG_INIT, FCP_K_PRECOND, FCP_K_DOT, FCP_K_BICG_ELEM,
       FCP_K_ASSEMBLE, FCP_K_GRADP, FCP_K_CORRECT_FLUX, FCP_K_GRAD, FCP_K_LAPLACIAN, FCP_K_SPMV, FCP_K_HALO, FCP_K_LIMITER,
       FCP_K_PISO_H, FCP_K_UVW, FCP_K_SCALAR,
       FCP_K_KRYLOV_PERSIST /* the whole solve as one persistent kernel; its phases are also booked under SPMV_DOT / CG_PK / CG_UPDATE / PRECOND / DOT */,
       FCP_K_COUNT };
int fcp_profile_enable(fcp_ctx *ctx, int on);
int fcp_profile_reset(fcp_ctx *ctx);
int fcp_profile_read(fcp_ctx *ctx, int kclass, double *total_ms, int64_t *launches);   /* synchronises; totals since reset */

/* ---- timing helper for benches: runs fn-independent CUDA event timing on the context stream --- */
int fcp_timer_start(fcp_ctx *ctx);
int fcp_timer_stop(fcp_ctx *ctx, float *milliseconds);        /* synchronises */
/* writes a buffer larger than L2 (256 MiB) so the next timed launch starts cold */
int fcp_flush_l2(fcp_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* FCP_H */
