/*
 * oracle/orc.h -- CPU restatement ("the oracle") of freeCappuccino's pressure-velocity
 * coupling hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (freecappuccino-dev_b200/) never links, imports or calls
 * it; the product fails loudly when its CUDA library is missing.
 *
 * PARITY PIN: the reference is Fortran and there is no Fortran compiler in this image, so the
 * reference itself cannot be run here (oracle/_ref cannot be built).  The oracle is pinned
 * against the known answers the reference's own tests print beside their output:
 *   test/test_linear_solvers_spsolve.f90:13-53   (two 5x5 systems)
 *   test/testFieldOperations/testFieldOperations.f90:137-160 (grad(x+y+z) = (1,1,1))
 *   applications/Poisson/poisson.f90:63-104      (sin(2 pi x) sin(2 pi y), O(h^2))
 *   src/mesh/wall_distance.f90:96-133            (laplacian + iccg + grad_gauss)
 *   Pressure/calcp_simple.f90:314                (sum(su) = 0 on a closed domain)
 * Beyond those the restatement itself is the pin ("parity unpinned by the reference").
 *
 * Conventions follow the Fortran: every index array is 1-based int32, reals are IEEE binary64,
 * gradients are (3,numTotal) column-major (xyz interleaved per cell).  Build with
 * -O2 -ffp-contract=off so that no FMA contraction changes the rounding the reference
 * (gfortran -O3, x86-64, no -march => no FMA) would produce.
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* patch type codes (reference: character bctype(ib), src/mesh/geometry.f90:62-70) */
enum { ORC_BC_WALL = 0, ORC_BC_INLET = 1, ORC_BC_OUTLET = 2, ORC_BC_SYMMETRY = 3,
       ORC_BC_PRESSURE = 4, ORC_BC_PERIODIC = 5, ORC_BC_EMPTY = 6, ORC_BC_PROCESS = 7 };

/* mirror of the reference's `geometry` module (src/mesh/geometry.f90:12-86) */
typedef struct {
  int32_t numCells, numInnerFaces, numBoundaryFaces, numFaces, numTotal, numBoundaries;
  const int32_t *owner;      /* [numFaces]      1-based */
  const int32_t *neighbour;  /* [numInnerFaces] 1-based */
  const double *arx, *ary, *arz, *xf, *yf, *zf; /* [numFaces] */
  const double *facint, *Df;                    /* [numInnerFaces] */
  const double *xc, *yc, *zc, *vol;             /* [numCells] (numTotal in the src-par layout) */
  const int32_t *bctype, *nfaces, *startFace, *iBndValueStart; /* [numBoundaries]; iface = startFace+i, ijb = iBndValueStart+i, i=1..nfaces */
  const int32_t *startFaceTwin; /* [numBoundaries] or NULL: for ORC_BC_PERIODIC patches the startFace of the twin patch (geometry.f90:82,251-257;
                                 * indexed per PATCH here, the reference indexes it by the running count of periodic patches iPer) */
} orc_mesh;

/* solver report, mirrors the values the reference prints (linear_solvers.f90:354-355) */
typedef struct {
  double res0;     /* initial L1 residual */
  double resl;     /* final   L1 residual */
  double factor;   /* sum|a_ii fi_i| + small after the first update */
  double resor;    /* res0/factor (the value returned through `resor`/`res0` argument) */
  int32_t iters;   /* itr_used */
} orc_report;

/* summation order used by the solver reductions:
 * ORC_SUM_SEQ  : left-to-right, what gfortran's inline SUM() does without -ffast-math
 * ORC_SUM_TREE : the fixed reduction tree the CUDA kernels use (see orc_sum_tree) */
enum { ORC_SUM_SEQ = 0, ORC_SUM_TREE = 1 };

double orc_small(void);   /* parameters.f90:6  small = 1e-20 (single literal) */
double orc_sum_tree(const double *v, int64_t n);

/* geometry.f90:416-530, 581-606 (variant 2), 648-664 */
void orc_geometry(int32_t numNodes, int32_t numCells, int32_t numInnerFaces, int32_t numFaces,
                  const double *x, const double *y, const double *z,
                  const int32_t *nnodes, const int32_t *node /* [nomax*numFaces] col-major node(j,iface), 1-based */,
                  int32_t nomax, const int32_t *owner, const int32_t *neighbour,
                  double *arx, double *ary, double *arz, double *xf, double *yf, double *zf,
                  double *vol, double *xc, double *yc, double *zc, double *facint, double *Df);
/* geometry.f90:698-754 : dnw/srdw for wall faces, dns/srds for symmetry faces, in patch order */
void orc_wall_geometry(const orc_mesh *m, double *dnw, double *srdw, double *dns, double *srds);

/* sparse_matrix.f90:86-296, periodic twins included (:141-171, :262-293): icell_jcell / jcell_icell have
 * numInnerFaces + numPeriodic entries, the periodic ones in patch order after the inner faces */
int32_t orc_num_periodic(const orc_mesh *m);
int32_t orc_csr_nnz(const orc_mesh *m);
void orc_csr_create(const orc_mesh *m, int32_t *ia, int32_t *ja, int32_t *diag,
                    int32_t *icell_jcell, int32_t *jcell_icell);

/* fvImplicit/laplacian.f90 : a is cleared, su is accumulated into */
void orc_laplacian(const orc_mesh *m, const int32_t *diag, const int32_t *icell_jcell,
                   const int32_t *jcell_icell, const double *mu, const double *phi,
                   double *a, int32_t nnz, double *su);

/* gradients.f90:1607-1693 */
void orc_grad_gauss(const orc_mesh *m, const double *u, double *dudxi);
/* gradients.f90:660-779 (weighted=0) and :1157-1326 (weighted=1) */
void orc_create_matrix_lsq(const orc_mesh *m, int weighted, double *Dmat);
/* gradients.f90:782-893 and :1334-1486 ; row2_correct=0 reproduces quirk Q1 */
void orc_grad_lsq(const orc_mesh *m, int weighted, int row2_correct, const double *Dmat,
                  const double *phi, double *dPhidxi);

/* Pressure/bpres.f90 */
void orc_bpres(const orc_mesh *m, double *p, const double *dPdxi, int istage);
/* Pressure/nablap.f90:19-208 ; pscheme 0 linear, 1 central, 2 weighted */
void orc_gradp_and_sources(const orc_mesh *m, int pscheme, double *p, const double *apu,
                           double *su, double *sv, double *sw, double *dPdxi);

/* calcp_simple.f90:69-234 incompressible, + faceflux_mass.f90:175-249,765-831,833-916 */
void orc_assemble_pcorr(const orc_mesh *m, const int32_t *diag, const int32_t *icell_jcell,
                        const int32_t *jcell_icell, int32_t nnz,
                        const double *den, double *u, double *v, double *w, const double *p,
                        double *pp, const double *dPdxi, const double *apu,
                        const double *apv, const double *apw /* read on periodic faces only (facefluxmass2_periodic :313-384); may be NULL otherwise */,
                        int const_mflux, double flomas,
                        double *a, double *su, double *flmass);
/* the same with the MPI tree's inner-face flux `facefluxmass` (quirk Q10; src-par/calcp_simple.f90:40-79, src-par/faceflux_mass.f90:28-180):
 * gU, gV, gW = the (3,numTotal) gradients of the tentative velocities; apv, apw are read on every inner face */
void orc_assemble_pcorr_mpi(const orc_mesh *m, const int32_t *diag, const int32_t *icell_jcell,
                            const int32_t *jcell_icell, int32_t nnz,
                            const double *den, double *u, double *v, double *w, const double *p,
                            double *pp, const double *dPdxi, const double *apu, const double *apv, const double *apw,
                            int const_mflux, double flomas, double *a, double *su, double *flmass,
                            const double *gU, const double *gV, const double *gW);
/* calcp_simple.f90:331-429 (one ipcorr pass after the solve) */
void orc_correct_simple(const orc_mesh *m, const int32_t *icell_jcell, int pscheme,
                        const double *a, const double *den, double *u, double *v, double *w, double *p, double *pp,
                        const double *apu, const double *apv, const double *apw,
                        double urfp, int32_t pRefCell,
                        double *su, double *sv, double *sw, double *dPdxi, double *flmass);
/* calcp_simple.f90:433-455 + faceflux_mass.f90:650-696 */
void orc_nonorth_corrector(const orc_mesh *m, const double *den, const double *apu,
                           const double *dPdxi, double *su, double *flmass);
/* velocity.f90:1184-1277 */
void orc_update_velocity_at_boundary(const orc_mesh *m, double *u, double *v, double *w);
/* src/cappuccino/constant_mass_flow_forcing.f90 + volumeWeightedAverage (fieldManipulation.f90:41-70); sum_mode as for the solvers.
 * Returns gragPplus; u(1:numCells) is corrected in place; *magUbarStar = the uncorrected bulk velocity. */
double orc_constant_mass_flow_forcing(const orc_mesh *m, double magUbar, const double *apu, double *u, int sum_mode, double *magUbarStar);
/* boundary/updateBoundary.f90: outlet/symmetry/pressure/empty copy the owner value, periodic faces and their twins take the mean of the two cells */
void orc_update_boundary(const orc_mesh *m, double *phi);


/* ---- slope limiters, gradients.f90:288-656 (applied in place to dPhidxi(3,numTotal)) ----------
 * kind: ORC_LIM_BJ :288-373, ORC_LIM_VENKAT :378-461, ORC_LIM_R3 :464-552 (the option string is 'R3'
 * but the active formula is R4), ORC_LIM_MDL :556-656.  BJ/Venkat/R3 use the GLOBAL extrema
 * fimax/fimin (quirk Q3); `1.e-6` is a default-real literal. */
enum { ORC_LIM_NONE = 0, ORC_LIM_BJ = 1, ORC_LIM_VENKAT = 2, ORC_LIM_R3 = 3, ORC_LIM_MDL = 4 };
void orc_slope_limiter(const orc_mesh *m, const int32_t *ia, const int32_t *ja, const int32_t *diag,
                       int kind, const double *phi, double *dPhidxi);

/* ---- QR least-squares gradient, gradients.f90:900-1152 + misc/matrix.f90:137-167 (inv), :366-419
 * (mgs_qr).  D is (3,6,numCells) column-major as in the reference.  Returns -1 when a cell has
 * more than m=6 faces (the reference would write out of bounds).  QUIRK Q20: the reference passes
 * R(6,3)/Q(6,6) actuals to mgs_qr's r(3,3)/q(6,3) dummies, so its R(1:3,1:3) picks r(1,1),r(2,1),
 * r(3,1),r(1,3),r(2,3),r(3,3) and three never-written stack words: its output is undefined.  The
 * oracle implements the algorithm the routine documents (thin QR by modified Gram-Schmidt). */
int orc_create_matrix_lsq_qr(const orc_mesh *m, double *D);
int orc_grad_lsq_qr(const orc_mesh *m, const double *D, const double *phi, double *dPhidxi);

/* ---- calcp_piso, Pressure/calcp_piso.f90:81-489 + faceflux_mass.f90:389-459 (facefluxmass_piso),
 * :564-647 (fluxmc), :699-762, :765-831, :833-916.  `a` holds the momentum coefficients on entry
 * (h = a, :81) and the pressure matrix of the last corrector on exit.  Periodic patches :248-295, :435-460 (H(U) runs over
 * the inner faces only, :110-124, as in the reference).
 * rep[(icorr-1)*npcor + (ipcorr-1)]. */
void orc_calcp_piso(const orc_mesh *m, const int32_t *ia, const int32_t *ja, const int32_t *diag,
                    const int32_t *icell_jcell, const int32_t *jcell_icell, int32_t nnz,
                    int solver, int32_t maxiter, double tol_abs, double tol_rel, int sum_mode,
                    int ncorr, int npcor, int pscheme, double urfp, int const_mflux, double flomas,
                    const double *rU, const double *rV, const double *rW,
                    const double *den, const double *apu, const double *apv, const double *apw,
                    double *a, double *h, double *u, double *v, double *w, double *p, double *pp,
                    double *su, double *sv, double *sw, double *dPdxi, double *flmass, orc_report *rep);

/* ---- calcuvw: the momentum predictor, Velocity/velocity.f90:50-750 (+ facefluxuvw :754-878, facefluxuvw_bnd :882-1034,
 * sngrad gradients.f90:1720-1779, the face_value family interpolation.f90:28-113, 116-650).  Tier "next" row f1.
 * Periodic patches :393-432 + facefluxuvw_periodic :1038-1180.  Not restated: Crank-Nicolson (:569-600), buoyancy (:186-200), MHD (:205-212).
 * cscheme: 0 cds, 1 central, 2 linearUpwind, 3 kappa, then the flux limiters of interpolation.f90:596-640 in source order:
 * 4 muscl, 5 umist, 6 koren, 7 smart, 8 avl-smart, 9 charm, 10 vanleer, 11 ospre, 12 minmod, 13 boundedLinearUpwind,
 * 14 boundedLinearUpwind02, 15 boundedCentral, 16 fromm, 17 cui, 18 quick, 19 spl13. */
typedef struct {
  int32_t solver, maxiter;        /* lSolverU, maxiterU */
  double tol_abs, tol_rel;        /* tolAbsU, tolRelU */
  double urf[3];                  /* urfU(1:3) */
  double gds;                     /* gdsU: deferred-correction blending */
  int32_t cscheme;                /* cSchemeU */
  int32_t grad_method;            /* 0 gauss, 1 lsq, 2 wlsq, 3 lsq_qr (the logicals of gradients.f90:118-138) */
  int32_t limiter;                /* ORC_LIM_* (gradients.f90:140-160) */
  int32_t pscheme;
  int32_t tscheme;                /* 0 steady, 1 bdf (or cn=.false.), 2 bdf2, 3 bdf3 */
  double timestep;
  int32_t piso;                   /* rU,rV,rW = su,sv,sw (:564-568) */
  int32_t const_mflux;
  double gradPcmf;
  double viscos;                  /* molecular viscosity: wall faces use max(viscos, visw) (:443) */
  int32_t sum_mode;               /* reductions inside the linear solver */
  int32_t pad;
} orc_uvw_params;
double orc_face_value(const orc_mesh *m, int cscheme, int32_t ijp, int32_t ijn /* 1-based */, double xf, double yf, double zf,
                      double lambda, const double *u, const double *dUdxi);
/* visw: [numBoundaryFaces] effective wall viscosity in the boundary slot order (only wall faces are read).
 * uo/uoo/uooo etc. may be NULL when tscheme does not need them.  a: stale values on entry (its diagonal enters the
 * first row sum, :606), the W-equation matrix on exit. */
void orc_calcuvw(const orc_mesh *m, const int32_t *ia, const int32_t *ja, const int32_t *diag,
                 const int32_t *icell_jcell, const int32_t *jcell_icell, int32_t nnz, const orc_uvw_params *prm,
                 double *u, double *v, double *w, double *p, const double *den, const double *vis, const double *visw,
                 const double *flmass, const double *uo, const double *vo, const double *wo,
                 const double *uoo, const double *voo, const double *woo, const double *uooo, const double *vooo, const double *wooo,
                 double *a, double *su, double *sv, double *sw, double *spu, double *spv, double *sp,
                 double *apu, double *apv, double *apw, double *dUdxi, double *dVdxi, double *dWdxi, double *dPdxi,
                 double *rU, double *rV, double *rW, orc_report *rep /* [3] */);

/* ---- row f4: scalar transport (fluxes/scalar_fluxes.f90:32-343) in the calcsc template of TurbulenceModels/k_epsilon_rlzb.f90
 * kind 0: GENERIC  -- su_vol/sp_vol hold the caller's volume sources, wall faces add nothing (zero flux)
 * kind 1: calcsc_tke     k_epsilon_rlzb.f90:52-445   (gen = |vis-viscos| S^2, wall cells: production from the wall shear stress, tau written)
 * kind 2: calcsc_epsilon k_epsilon_rlzb.f90:447-790  (realizable c1, wall cells: row zeroed, ed = cmu75 k^1.5/(cappa dnw) imposed)
 * Steps: grad(phi) with the configured method/limiter; volume sources + bdf/bdf2 term; facefluxsc on inner faces, facefluxsc_boundary on
 * inlet/outlet/pressure patches, facefluxsc_periodic on periodic pairs, the wall treatment; a(diag) = sp - sum(off-diagonals) in CSR order,
 * under-relaxation; csrsolve; updateBoundary; min/max report and the clip to `small` when the minimum is negative.  Crank-Nicolson and
 * buoyancy are not restated.  Per-wall-face arrays (visw, dnw, tau) are indexed by BOUNDARY FACE ordinal here (the reference counts
 * wall faces with iWall). */
typedef struct {
  int32_t kind, solver, maxiter, cscheme, grad_method, limiter, tscheme, sum_mode;
  double tol_abs, tol_rel, urf, gds, timestep, prtr, viscos, densit;
} orc_scalar_params;
void orc_calcsc(const orc_mesh *m, const int32_t *ia, const int32_t *ja, const int32_t *diag, const int32_t *icell_jcell,
                const int32_t *jcell_icell, int32_t nnz, const orc_scalar_params *prm,
                double *phi, const double *phio, const double *phioo, double *te, double *ed /* the k and epsilon fields; phi aliases one of them for kind 1/2 */,
                const double *den, const double *vis, const double *visw, const double *dnw, const double *flmass,
                const double *u, const double *v, const double *w, const double *magStrain, double *gen, double *tau,
                const double *su_vol, const double *sp_vol,
                double *a, double *su, double *sp, double *dPhidxi, orc_report *rep, double *fimin, double *fimax,
                double *fsst /* kinds 3, 4: the SST blending function F1 (written by kind 4, read by both) */, const double *walldist,
                const double *dTEdxi /* kind 4: the gradient of k the k call left behind */, int lowre);
/* kinds 3 and 4 of orc_calcsc: the k and omega equations of TurbulenceModels/k_omega_SST.f90:91-788 (same template; production limiter,
 * F1 = tanh(ksi^4), cross diffusion, blended constants, sigma taken from the OWNER cell of a face, omega imposed in wall cells, min/max and
 * the clip taken over the whole array).  modify_mu_eff :790-958. */
void orc_modify_mu_eff_sst(const orc_mesh *m, double urf, double viscos, double densit, int lowre, const double *magStrain, const double *walldist,
                           const double *te, const double *ed, const double *den, const double *u, const double *v, const double *w,
                           const double *dnw, double *vis, double *visw, double *ypl, double *tau);
/* fvExplicit/calc_strain_and_vorticity.f90 */
void orc_calc_strain_and_vorticity(const orc_mesh *m, const double *dUdxi, const double *dVdxi, const double *dWdxi, double *magStrain, double *vorticity);
/* modify_mu_eff of the realizable k-epsilon model, k_epsilon_rlzb.f90:792-975 (cell loop, updateBoundary(vis), wall functions) */
void orc_modify_mu_eff_rlzb(const orc_mesh *m, double urf, double viscos, const double *dUdxi, const double *dVdxi, const double *dWdxi,
                            const double *te, const double *ed, const double *den, const double *u, const double *v, const double *w,
                            const double *dnw, double *vis, double *visw, double *ypl, double *tau);

/* ---- SGS viscosity of the LES models (row f4): TurbulenceModels/wale_sgs.f90:33-185 (model 0) and vremanSGS.f90:33-179 (model 1), written in the
 * reference with the operator-overloaded tensorFields module (finiteVolume/tensorFields/tensorFields.f90) on top of fvxGradient's Grad(U).
 * QUIRK Q24: inner_product_rank2_tensors computes its zx component as T1zx*T2xx + T1zx*T2yx + T1zz*T2zx (tensorFields.f90:508: `zx` twice);
 * reproduced.  Boundary: wall faces get visw = vis = max(viscos, 0), a periodic face and its twin the mean of the two cells, every other patch
 * the owner value.  visw is indexed by BOUNDARY FACE ordinal. */
void orc_grad_gauss_iter(const orc_mesh *m, const double *u, int npass, double *dudx, double *dudy, double *dudz);   /* src-par/gradients.f90:1547-1664 (nigrad passes of gradco) */
void orc_grad_gauss_fvx(const orc_mesh *m, const double *u, double *dudx, double *dudy, double *dudz);   /* fvxGradient.f90:1549-1662 (two passes, gradco) */
void orc_modify_viscosity_sgs(const orc_mesh *m, int model, double urf, double viscos, const double *u, const double *v, const double *w,
                              const double *den, double *vis, double *visw);

/* linear_solvers.f90:206-359, 364-545, 548-786 */
void orc_spmv(int32_t n, const int32_t *ia, const int32_t *ja, const double *a, const double *x, double *y);
void orc_dpcg(int32_t n, int32_t nnz, const int32_t *ia, const int32_t *ja, const double *a,
              const int32_t *diag, double *fi, const double *rhs, int32_t itr_max,
              double tol_abs, double tol_rel, int sum_mode, orc_report *rep);
void orc_iccg(int32_t n, int32_t nnz, const int32_t *ia, const int32_t *ja, const double *a,
              const int32_t *diag, double *fi, const double *rhs, int32_t itr_max,
              double tol_abs, double tol_rel, int sum_mode, orc_report *rep);
void orc_bicgstab(int32_t n, int32_t nnz, const int32_t *ia, const int32_t *ja, const double *a,
                  const int32_t *diag, double *fi, const double *rhs, int32_t itr_max,
                  double tol_abs, double tol_rel, int sum_mode, orc_report *rep);
/* the report line of linear_solvers.f90:354-355 / 540-541 / 781-782, written into buf */
int orc_report_line(int solver /*1 dpcg 2 iccg 3 bicgstab*/, const char *chvar, const orc_report *rep,
                    char *buf, int buflen);

/* ---- src-par layout: P partitions driven in lock-step inside one process ("virtual ranks").
 * Each rank has its own orc_mesh with PROCESS patches, a local CSR, and apr[npro] (one
 * off-rank coefficient per process face, patch order).  Follows src-par/dpcg.f90:60-190,
 * src-par/exchange.f90:48-127 and global_sum_mpi.f90, except that the stop test uses the
 * serial tree's criterion so that counts are comparable with orc_dpcg.
 * peer_rank[r][ipatch]/peer_patch: which rank / which of its process patches faces patch ipatch. */
typedef struct {
  const orc_mesh *mesh;
  const int32_t *ia, *ja, *diag;
  const double *a;
  const double *apr;          /* [npro] */
  int32_t npro;
  const int32_t *peer_rank;   /* [numBoundaries] (-1 for non-process patches) */
  const int32_t *peer_patch;  /* [numBoundaries] index of the matching patch on the peer */
  double *fi;                 /* [numTotal] */
  const double *rhs;          /* [numCells] */
} orc_rank;
void orc_exchange(int32_t nranks, const orc_rank *ranks, double **phi /* [nranks][numTotal] */);
void orc_dpcg_par(int32_t nranks, orc_rank *ranks, int32_t itr_max, double tol_abs, double tol_rel,
                  int sum_mode, orc_report *rep);

#ifdef __cplusplus
}
#endif
#endif
