// host/fcp_host.hpp -- C++ mirror of the reference's Fortran module-level API for the hot path, on top of the C-ABI
// (include/fcp.h).  The reference keeps its state in module globals (geometry, sparse_matrix, variables) and its
// procedures take few or no arguments; this header keeps that shape so that a driver written
// against it reads like the Fortran one (compare host/poisson_app.cpp with applications/Poisson/poisson.f90:50-104).
// Arrays are std::vector with the Fortran's 1-based CONTENTS (index arrays hold 1-based values); element i of the
// Fortran array is [i-1].
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../include/fcp.h"

namespace fcp {

typedef double dp;   // types.f90:6

// ---- module geometry (src/mesh/geometry.f90:12-86) ----------------------------------------------------------------------
namespace geometry {
inline int numCells = 0, numInnerFaces = 0, numBoundaryFaces = 0, numFaces = 0, numTotal = 0, numBoundaries = 0;
inline std::vector<int32_t> owner, neighbour, nfaces, startFace, iBndValueStart, bctype;
inline std::vector<int32_t> startFaceTwin;   // per patch (-1 unless periodic); empty = no periodic patch (geometry.f90:82 indexes it by iPer)
inline int numPeriodic = 0;
inline std::vector<dp> arx, ary, arz, xf, yf, zf, facint, Df, xc, yc, zc, vol;
}
// ---- module sparse_matrix (src/sparseMatrix/sparse_matrix.f90:20-40) -----------------------------------------------------
namespace sparse_matrix {
inline int nnz = 0;
inline std::vector<int32_t> ia, ja, diag, icell_jcell_csr_index, jcell_icell_csr_index;
inline std::vector<dp> a, su, sv, sw, apu, apv, apw;
inline std::vector<dp> h, rU, rV, rW;   // calcp_piso.f90:81 `h = a`; velocity.f90:567 rU = su
}
// ---- module variables ---------------------------------------------------------------------------------------------------------
namespace variables {
inline std::vector<dp> u, v, w, p, pp, den, flmass, dPdxi;
inline std::vector<dp> vis, visw, dUdxi, dVdxi, dWdxi;             // visw(iWall): wall faces in patch order (velocity.f90:441-443)
inline std::vector<dp> uo, vo, wo, uoo, voo, woo, uooo, vooo, wooo;  // past time levels
}
// ---- module parameters / pressure (parameters.f90, Pressure/pressure.f90:28-33) ------------------------------------------------
namespace parameters {
inline int pRefCell = 1, npcor = 1, ncorr = 2;
inline bool ltransient = false, bdf = false, bdf2 = false, bdf3 = false, piso = false;
inline dp timestep = 0.0, gradPcmf = 0.0, viscos = 0.0;
inline bool const_mflux = false;
inline dp flomas = 0.0;
}
namespace velocity {   // Velocity/velocity.f90:22-42
inline dp urfU[3] = {0.8, 0.8, 0.8}, gdsU = 1.0, tolAbsU = 1e-13, tolRelU = 0.025;
inline int maxiterU = 5;
inline std::string cSchemeU = "cds", lSolverU = "bicgstab";
}
namespace gradients {  // gradients.f90:44-52
inline bool lstsq = false, lstsq_qr = false, lstsq_dm = false;
inline std::string limiter = "none";
}
namespace pressure {
inline dp urfP = 0.2, tolAbsP = 1e-13, tolRelP = 0.025;
inline int maxiterP = 30;
inline std::string lSolverP = "iccg", pscheme = "linear";
}

inline fcp_ctx *ctx = nullptr;

inline void check(int rc, const char *what) {   // the reference prints and stops (linear_solvers.f90:1361-1376)
  if (rc != FCP_OK) {
    std::fprintf(stderr, " libfcp_b200: %s failed with code %d: %s\n", what, rc, fcp_last_error());
    std::exit(1);
  }
}
inline int solver_id(const std::string &s) {
  if (s == "dpcg") return FCP_SOLVER_DPCG;
  if (s == "iccg") return FCP_SOLVER_ICCG;
  if (s == "bicgstab") return FCP_SOLVER_BICGSTAB;
  if (s == "gauss-seidel") return FCP_SOLVER_GAUSS_SEIDEL;
  std::fprintf(stderr, " libfcp_b200: linear solver \"%s\" is not on the accelerated path\n", s.c_str());
  std::exit(1);
}
inline int pscheme_id(const std::string &s) { return s == "central" ? FCP_PSCHEME_CENTRAL : s == "weighted" ? FCP_PSCHEME_WEIGHTED : FCP_PSCHEME_LINEAR; }
inline void put(int field, const std::vector<dp> &x, int64_t n) { check(fcp_field_upload(ctx, field, x.data(), n), "fcp_field_upload"); }
inline void get(int field, std::vector<dp> &x, int64_t n) { check(fcp_field_download(ctx, field, x.data(), n), "fcp_field_download"); }

// create_CSR_matrix (sparse_matrix.f90:86): uploads the mesh of module geometry, fills module sparse_matrix
inline void create_CSR_matrix(int device = 0) {
  using namespace geometry;
  using namespace sparse_matrix;
  fcp_mesh_desc md{};
  md.numCells = numCells; md.numInnerFaces = numInnerFaces; md.numBoundaryFaces = numBoundaryFaces; md.numBoundaries = numBoundaries;
  md.owner = owner.data(); md.neighbour = neighbour.data();
  md.arx = arx.data(); md.ary = ary.data(); md.arz = arz.data(); md.xf = xf.data(); md.yf = yf.data(); md.zf = zf.data();
  md.facint = facint.data(); md.Df = Df.data(); md.xc = xc.data(); md.yc = yc.data(); md.zc = zc.data(); md.vol = vol.data();
  md.bctype = bctype.data(); md.nfaces = nfaces.data(); md.startFace = startFace.data();
  md.startFaceTwin = startFaceTwin.empty() ? nullptr : startFaceTwin.data();
  check(fcp_ctx_create(&md, device, &ctx), "fcp_ctx_create");
  int32_t n, nt, nf, nz, npro;
  check(fcp_ctx_sizes(ctx, &n, &nt, &nf, &nz, &npro), "fcp_ctx_sizes");
  nnz = nz;
  ia.resize(n + 1); ja.resize(nnz); diag.resize(n);
  icell_jcell_csr_index.resize(numInnerFaces + numPeriodic); jcell_icell_csr_index.resize(numInnerFaces + numPeriodic);   // sparse_matrix.f90:246-247
  check(fcp_csr_pattern(ctx, ia.data(), ja.data(), diag.data(), icell_jcell_csr_index.data(), jcell_icell_csr_index.data()), "fcp_csr_pattern");
  a.assign(nnz, 0.0);
  for (auto *vec : {&su, &sv, &sw, &apu, &apv, &apw}) vec->assign(numCells, 0.0);   // sparse_matrix.f90:215-246
}

// csrsolve(solver, fi, rhs, res0, itr_max, tol_abs, tol_rel, chvar)   linear_solvers.f90:40-59
inline void csrsolve(const std::string &solver, std::vector<dp> &fi, const std::vector<dp> &rhs, dp &res0, int itr_max, dp tol_abs,
                     dp tol_rel, const std::string &chvar) {
  put(FCP_F_A, sparse_matrix::a, sparse_matrix::nnz);
  put(FCP_F_S0, fi, geometry::numTotal);
  put(FCP_F_S1, rhs, geometry::numCells);
  fcp_report rep;
  check(fcp_csrsolve(ctx, solver_id(solver), FCP_F_S0, FCP_F_S1, itr_max, tol_abs, tol_rel, &rep), "fcp_csrsolve");
  get(FCP_F_S0, fi, geometry::numCells);
  res0 = rep.resor;
  char line[256];
  check(fcp_report_line(&rep, chvar.c_str(), line, sizeof(line)), "fcp_report_line");
  std::puts(line);   // the reference's report line (linear_solvers.f90:354-355)
}

// grad_gauss(u, dudxi)  gradients.f90:1607 ; dudxi is (3,numTotal) column-major
inline void grad_gauss(const std::vector<dp> &u_, std::vector<dp> &dudxi) {
  put(FCP_F_S0, u_, geometry::numTotal);
  check(fcp_grad(ctx, FCP_GRAD_GAUSS, FCP_F_S0, FCP_F_G0, 1), "fcp_grad");
  get(FCP_F_G0, dudxi, 3 * (int64_t)geometry::numTotal);
}
// laplacian(mu, phi): fills sparse_matrix::a, accumulates into sparse_matrix::su   fvImplicit/laplacian.f90
inline void laplacian(const std::vector<dp> &mu, const std::vector<dp> &phi) {
  put(FCP_F_S0, mu, geometry::numCells);
  put(FCP_F_S1, phi, geometry::numTotal);
  put(FCP_F_SU, sparse_matrix::su, geometry::numCells);
  check(fcp_laplacian(ctx, FCP_F_S0, FCP_F_S1), "fcp_laplacian");
  get(FCP_F_A, sparse_matrix::a, sparse_matrix::nnz);
  get(FCP_F_SU, sparse_matrix::su, geometry::numCells);
}
// gradp_and_sources(p)   Pressure/nablap.f90:19
inline void gradp_and_sources(std::vector<dp> &p_) {
  using namespace sparse_matrix;
  put(FCP_F_P, p_, geometry::numTotal);
  put(FCP_F_APU, apu, geometry::numCells);
  check(fcp_gradp_and_sources(ctx, pscheme_id(pressure::pscheme), FCP_F_P), "fcp_gradp_and_sources");
  get(FCP_F_P, p_, geometry::numTotal);
  get(FCP_F_SU, su, geometry::numCells); get(FCP_F_SV, sv, geometry::numCells); get(FCP_F_SW, sw, geometry::numCells);
  get(FCP_F_DPDXI, variables::dPdxi, 3 * (int64_t)geometry::numTotal);
}
// calcp_simple(): no arguments, everything through the modules   Pressure/calcp_simple.f90
inline void calcp_simple() {
  using namespace variables;
  using namespace sparse_matrix;
  const int nT = geometry::numTotal, n = geometry::numCells;
  put(FCP_F_U, u, nT); put(FCP_F_V, v, nT); put(FCP_F_W, w, nT); put(FCP_F_P, p, nT); put(FCP_F_PP, pp, nT); put(FCP_F_DEN, den, nT);
  put(FCP_F_APU, apu, n); put(FCP_F_APV, apv, n); put(FCP_F_APW, apw, n);
  put(FCP_F_DPDXI, dPdxi, 3 * (int64_t)nT); put(FCP_F_FLMASS, flmass, geometry::numFaces);
  fcp_simple_params prm{};
  prm.solver = solver_id(pressure::lSolverP); prm.maxiter = pressure::maxiterP; prm.tol_abs = pressure::tolAbsP; prm.tol_rel = pressure::tolRelP;
  prm.urfp = pressure::urfP; prm.npcor = parameters::npcor; prm.pRefCell = parameters::pRefCell; prm.pscheme = pscheme_id(pressure::pscheme);
  prm.const_mflux = parameters::const_mflux; prm.flomas = parameters::flomas; prm.zero_pp = 0;
  std::vector<fcp_report> rep(parameters::npcor);
  check(fcp_calcp_simple(ctx, &prm, rep.data()), "fcp_calcp_simple");
  for (auto &r : rep) {
    char line[256];
    fcp_report_line(&r, "p", line, sizeof(line));
    std::puts(line);
  }
  get(FCP_F_U, u, nT); get(FCP_F_V, v, nT); get(FCP_F_W, w, nT); get(FCP_F_P, p, nT); get(FCP_F_PP, pp, nT);
  get(FCP_F_FLMASS, flmass, geometry::numFaces); get(FCP_F_DPDXI, dPdxi, 3 * (int64_t)nT);
  get(FCP_F_SU, su, n); get(FCP_F_SV, sv, n); get(FCP_F_SW, sw, n); get(FCP_F_A, a, nnz);
}
// grad(phi, dPhidxi, option, option_limiter)   gradients.f90:217-278
inline void grad(const std::vector<dp> &phi, std::vector<dp> &dPhidxi, const std::string &option, const std::string &option_limiter) {
  const int method = option == "lsq" ? FCP_GRAD_LSQ : option == "lsq_qr" ? FCP_GRAD_LSQ_QR : option == "wlsq" ? FCP_GRAD_LSQ_DM
                   : option == "gauss" ? FCP_GRAD_GAUSS : -1;
  const int limiter = option_limiter == "Barth-Jespersen" ? FCP_LIMITER_BARTH_JESPERSEN : option_limiter == "Venkatakrishnan" ? FCP_LIMITER_VENKATAKRISHNAN
                    : option_limiter == "R3" ? FCP_LIMITER_R3 : option_limiter == "multidimensional" ? FCP_LIMITER_MULTIDIMENSIONAL : FCP_LIMITER_NONE;
  if (method < 0) { dPhidxi.assign(dPhidxi.size(), 0.0); return; }   // no branch taken: dPhidxi stays 0 (:238)
  put(FCP_F_S0, phi, geometry::numTotal);
  if (method != FCP_GRAD_GAUSS) check(fcp_create_lsq_grad_matrix(ctx, method), "fcp_create_lsq_grad_matrix");
  check(fcp_grad_opt(ctx, method, limiter, FCP_F_S0, FCP_F_G0), "fcp_grad_opt");
  get(FCP_F_G0, dPhidxi, 3 * (int64_t)geometry::numTotal);
}
// calcuvw(): no arguments   Velocity/velocity.f90:50-750
inline void calcuvw() {
  using namespace variables;
  using namespace sparse_matrix;
  using namespace geometry;
  static const char *names[] = {"cds", "central", "linearUpwind", "kappa", "muscl", "umist", "koren", "smart", "avl-smart", "charm", "vanleer", "ospre",
                                "minmod", "boundedLinearUpwind", "boundedLinearUpwind02", "boundedCentral", "fromm", "cui", "quick", "spl13"};
  const int nT = numTotal, n = numCells;
  fcp_uvw_params prm{};
  prm.cscheme = -1;
  for (int k = 0; k < FCP_CS_COUNT; ++k) if (velocity::cSchemeU == names[k]) prm.cscheme = k;
  if (prm.cscheme < 0) { std::fprintf(stderr, "Fatal error: non-existing interpolation scheme!\n"); std::exit(1); }   // interpolation.f90:643-646
  put(FCP_F_U, u, nT); put(FCP_F_V, v, nT); put(FCP_F_W, w, nT); put(FCP_F_P, p, nT); put(FCP_F_DEN, den, nT); put(FCP_F_VIS, vis, nT);
  put(FCP_F_APU, apu, n);
  std::vector<dp> viswf(nT, 0.0);
  size_t iw = 0;
  for (int ib = 0; ib < numBoundaries; ++ib)
    if (bctype[ib] == FCP_BC_WALL)
      for (int i = 0; i < nfaces[ib]; ++i) viswf[iBndValueStart[ib] + i] = iw < visw.size() ? visw[iw++] : 0.0;
  put(FCP_F_VISW, viswf, nT); put(FCP_F_FLMASS, flmass, numFaces); put(FCP_F_A, a, nnz);
  prm.tscheme = !parameters::ltransient ? 0 : parameters::bdf3 ? 3 : parameters::bdf2 ? 2 : 1;
  if (prm.tscheme >= 1) { put(FCP_F_UO, uo, nT); put(FCP_F_VO, vo, nT); put(FCP_F_WO, wo, nT); }
  if (prm.tscheme >= 2) { put(FCP_F_UOO, uoo, nT); put(FCP_F_VOO, voo, nT); put(FCP_F_WOO, woo, nT); }
  if (prm.tscheme >= 3) { put(FCP_F_UOOO, uooo, nT); put(FCP_F_VOOO, vooo, nT); put(FCP_F_WOOO, wooo, nT); }
  prm.solver = solver_id(velocity::lSolverU); prm.maxiter = velocity::maxiterU; prm.tol_abs = velocity::tolAbsU; prm.tol_rel = velocity::tolRelU;
  for (int q = 0; q < 3; ++q) prm.urf[q] = velocity::urfU[q];
  prm.gds = velocity::gdsU;
  prm.grad_method = gradients::lstsq ? FCP_GRAD_LSQ : gradients::lstsq_qr ? FCP_GRAD_LSQ_QR : gradients::lstsq_dm ? FCP_GRAD_LSQ_DM : FCP_GRAD_GAUSS;
  const std::string &lim = gradients::limiter;
  prm.limiter = lim == "Barth-Jespersen" ? FCP_LIMITER_BARTH_JESPERSEN : lim == "Venkatakrishnan" ? FCP_LIMITER_VENKATAKRISHNAN
              : lim == "R3" ? FCP_LIMITER_R3 : lim == "multidimensional" ? FCP_LIMITER_MULTIDIMENSIONAL : FCP_LIMITER_NONE;
  prm.pscheme = pscheme_id(pressure::pscheme); prm.piso = parameters::piso; prm.timestep = parameters::timestep;
  prm.const_mflux = parameters::const_mflux; prm.gradPcmf = parameters::gradPcmf; prm.viscos = parameters::viscos;
  fcp_report rep[3];
  check(fcp_calcuvw(ctx, &prm, rep), "fcp_calcuvw");
  const char *chvar[3] = {"U", "V", "W"};
  for (int q = 0; q < 3; ++q) {
    char line[256];
    fcp_report_line(&rep[q], chvar[q], line, sizeof(line));
    std::puts(line);
  }
  get(FCP_F_U, u, nT); get(FCP_F_V, v, nT); get(FCP_F_W, w, nT); get(FCP_F_P, p, nT);
  get(FCP_F_APU, apu, n); get(FCP_F_APV, apv, n); get(FCP_F_APW, apw, n);
  get(FCP_F_SU, su, n); get(FCP_F_SV, sv, n); get(FCP_F_SW, sw, n);
  for (auto *g : {&dUdxi, &dVdxi, &dWdxi, &dPdxi}) g->resize(3 * (size_t)nT);
  get(FCP_F_DUDXI, dUdxi, 3 * (int64_t)nT); get(FCP_F_DVDXI, dVdxi, 3 * (int64_t)nT); get(FCP_F_DWDXI, dWdxi, 3 * (int64_t)nT);
  get(FCP_F_DPDXI, dPdxi, 3 * (int64_t)nT); get(FCP_F_A, a, nnz);
  if (parameters::piso) { rU.resize(n); rV.resize(n); rW.resize(n); get(FCP_F_RU, rU, n); get(FCP_F_RV, rV, n); get(FCP_F_RW, rW, n); }
}
// updateBoundary(phi)   src/finiteVolume/boundary/updateBoundary.f90
inline void updateBoundary(std::vector<dp> &phi) {
  put(FCP_F_S0, phi, (int64_t)phi.size());
  check(fcp_update_boundary(ctx, FCP_F_S0), "fcp_update_boundary");
  get(FCP_F_S0, phi, (int64_t)phi.size());
}
// calcsc: the scalar transport template of the turbulence models (k_epsilon_rlzb.f90:52-790 + scalar_fluxes.f90); the fields it reads
// (den, vis, flmass, te, ed, magStrain, dnw, visw, u, v, w, phio/phioo) are uploaded by the caller with put()
inline fcp_report calcsc(int kind, int phi_field, int solver, int maxiter, dp tolAbs, dp tolRel, dp urf, dp gds, int cscheme, dp prtr, dp viscos, dp densit,
                         dp *fimin = nullptr, dp *fimax = nullptr, int tscheme = 0, dp timestep = 0.0, int grad_method = FCP_GRAD_GAUSS,
                         int limiter = FCP_LIMITER_NONE) {
  fcp_scalar_params prm{};
  prm.kind = kind; prm.solver = solver; prm.maxiter = maxiter; prm.cscheme = cscheme; prm.grad_method = grad_method; prm.limiter = limiter;
  prm.tscheme = tscheme; prm.tol_abs = tolAbs; prm.tol_rel = tolRel; prm.urf = urf; prm.gds = gds; prm.timestep = timestep; prm.prtr = prtr;
  prm.viscos = viscos; prm.densit = densit;
  fcp_report rep{};
  check(fcp_calcsc(ctx, &prm, phi_field, &rep, fimin, fimax), "fcp_calcsc");
  return rep;
}
// modify_viscosity_wale_sgs / modify_viscosity_vreman_sgs (u, v, w, den, vis already on the device)
inline void modify_viscosity_sgs(int model, dp urfVis, dp viscos) { check(fcp_modify_viscosity_sgs(ctx, model, urfVis, viscos), "fcp_modify_viscosity_sgs"); }
inline void modify_mu_eff_k_omega_sst(dp urfVis, dp viscos, dp densit, bool lowRe = false) {
  check(fcp_modify_mu_eff_k_omega_sst(ctx, urfVis, viscos, densit, lowRe ? 1 : 0), "fcp_modify_mu_eff_k_omega_sst");
}
// wall_distance (src/mesh/wall_distance.f90): result in field FCP_F_WALLDIST
inline fcp_report wall_distance() { fcp_report rep{}; check(fcp_wall_distance(ctx, &rep), "fcp_wall_distance"); return rep; }
inline void calc_strain_and_vorticity() { check(fcp_calc_strain_and_vorticity(ctx), "fcp_calc_strain_and_vorticity"); }
inline void modify_mu_eff_k_epsilon_rlzb(dp urfVis, dp viscos) { check(fcp_modify_mu_eff_k_epsilon_rlzb(ctx, urfVis, viscos), "fcp_modify_mu_eff_k_epsilon_rlzb"); }
// constant_mass_flow_forcing   src/cappuccino/constant_mass_flow_forcing.f90 (U and APU already on the device)
inline dp constant_mass_flow_forcing(dp magUbar, dp &gradPcmf) {
  dp ustar = 0.0;
  check(fcp_constant_mass_flow_forcing(ctx, magUbar, &gradPcmf, &ustar), "fcp_constant_mass_flow_forcing");
  return ustar;
}

// calcp_piso(): no arguments   Pressure/calcp_piso.f90
inline void calcp_piso() {
  using namespace variables;
  using namespace sparse_matrix;
  const int nT = geometry::numTotal, n = geometry::numCells;
  put(FCP_F_U, u, nT); put(FCP_F_V, v, nT); put(FCP_F_W, w, nT); put(FCP_F_P, p, nT); put(FCP_F_PP, pp, nT); put(FCP_F_DEN, den, nT);
  put(FCP_F_APU, apu, n); put(FCP_F_APV, apv, n); put(FCP_F_APW, apw, n);
  put(FCP_F_RU, rU, n); put(FCP_F_RV, rV, n); put(FCP_F_RW, rW, n); put(FCP_F_A, a, nnz);
  put(FCP_F_DPDXI, dPdxi, 3 * (int64_t)nT); put(FCP_F_FLMASS, flmass, geometry::numFaces);
  fcp_piso_params prm{};
  prm.solver = solver_id(pressure::lSolverP); prm.maxiter = pressure::maxiterP; prm.tol_abs = pressure::tolAbsP; prm.tol_rel = pressure::tolRelP;
  prm.urfp = pressure::urfP; prm.ncorr = parameters::ncorr; prm.npcor = parameters::npcor; prm.pscheme = pscheme_id(pressure::pscheme);
  prm.const_mflux = parameters::const_mflux; prm.flomas = parameters::flomas;
  std::vector<fcp_report> rep((size_t)parameters::ncorr * parameters::npcor);
  check(fcp_calcp_piso(ctx, &prm, rep.data()), "fcp_calcp_piso");
  for (auto &r : rep) {
    char line[256];
    fcp_report_line(&r, "p", line, sizeof(line));
    std::puts(line);
  }
  get(FCP_F_U, u, nT); get(FCP_F_V, v, nT); get(FCP_F_W, w, nT); get(FCP_F_P, p, nT); get(FCP_F_PP, pp, nT);
  get(FCP_F_FLMASS, flmass, geometry::numFaces); get(FCP_F_DPDXI, dPdxi, 3 * (int64_t)nT);
  get(FCP_F_SU, su, n); get(FCP_F_SV, sv, n); get(FCP_F_SW, sw, n); get(FCP_F_A, a, nnz);
  h.resize(nnz); get(FCP_F_H, h, nnz);
}
inline void finalize() {
  if (ctx) fcp_ctx_destroy(ctx);
  ctx = nullptr;
}

// uniform box mesh nx x ny x nz on [0,lx]x[0,ly]x[0,lz], cells i-fastest, six wall patches: fills module geometry
inline void box_mesh(int nx, int ny, int nz, dp lx, dp ly, dp lz) {
  using namespace geometry;
  const dp dx = lx / nx, dy = ly / ny, dz = lz / nz;
  numCells = nx * ny * nz;
  owner.clear(); neighbour.clear();
  for (auto *vec : {&arx, &ary, &arz, &xf, &yf, &zf, &facint, &Df}) vec->clear();
  auto cid = [&](int i, int j, int k) { return i + nx * (j + ny * k) + 1; };
  auto face = [&](int o, dp sx, dp sy, dp sz, dp x, dp y, dp z) { owner.push_back(o); arx.push_back(sx); ary.push_back(sy); arz.push_back(sz); xf.push_back(x); yf.push_back(y); zf.push_back(z); };
  for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
    const dp x = (i + 0.5) * dx, y = (j + 0.5) * dy, z = (k + 0.5) * dz;
    if (i < nx - 1) { face(cid(i, j, k), dy * dz, 0, 0, x + 0.5 * dx, y, z); neighbour.push_back(cid(i + 1, j, k)); facint.push_back(0.5); Df.push_back(dy * dz / dx); }
    if (j < ny - 1) { face(cid(i, j, k), 0, dx * dz, 0, x, y + 0.5 * dy, z); neighbour.push_back(cid(i, j + 1, k)); facint.push_back(0.5); Df.push_back(dx * dz / dy); }
    if (k < nz - 1) { face(cid(i, j, k), 0, 0, dx * dy, x, y, z + 0.5 * dz); neighbour.push_back(cid(i, j, k + 1)); facint.push_back(0.5); Df.push_back(dx * dy / dz); }
  }
  numInnerFaces = (int)neighbour.size();
  nfaces.clear(); startFace.clear(); bctype.clear();
  auto patch = [&](int count) { startFace.push_back((int)owner.size() - count); nfaces.push_back(count); bctype.push_back(FCP_BC_WALL); };
  for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) face(cid(i, ny - 1, k), 0, dx * dz, 0, (i + 0.5) * dx, ly, (k + 0.5) * dz);
  patch(nx * nz);
  for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) face(cid(i, 0, k), 0, -dx * dz, 0, (i + 0.5) * dx, 0, (k + 0.5) * dz);
  patch(nx * nz);
  for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) face(cid(0, j, k), -dy * dz, 0, 0, 0, (j + 0.5) * dy, (k + 0.5) * dz);
  patch(ny * nz);
  for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) face(cid(nx - 1, j, k), dy * dz, 0, 0, lx, (j + 0.5) * dy, (k + 0.5) * dz);
  patch(ny * nz);
  for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) face(cid(i, j, 0), 0, 0, -dx * dy, (i + 0.5) * dx, (j + 0.5) * dy, 0);
  patch(nx * ny);
  for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) face(cid(i, j, nz - 1), 0, 0, dx * dy, (i + 0.5) * dx, (j + 0.5) * dy, lz);
  patch(nx * ny);
  numFaces = (int)owner.size();
  numBoundaryFaces = numFaces - numInnerFaces;
  numBoundaries = 6;
  numTotal = numCells + numBoundaryFaces;
  xc.resize(numCells); yc.resize(numCells); zc.resize(numCells); vol.assign(numCells, dx * dy * dz);
  for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
    const int c = cid(i, j, k) - 1;
    xc[c] = (i + 0.5) * dx; yc[c] = (j + 0.5) * dy; zc[c] = (k + 0.5) * dz;
  }
  iBndValueStart.resize(6);
  for (int ib = 0; ib < 6; ++ib) iBndValueStart[ib] = numCells + startFace[ib] - numInnerFaces;
}

// ---- src-par tree: exchange(phi), global_sum / global_isum / global_max / global_min   (src-par/exchange.f90:3, global_*_mpi.f90) ----
// (after fcp_comm_init(ctx, rank, nranks, id, peer_rank); on a single rank they are the identity, like the MPI routines on one process)
inline void exchange(std::vector<dp> &phi) {
  put(FCP_F_S0, phi, geometry::numTotal);
  check(fcp_exchange(ctx, FCP_F_S0), "fcp_exchange");
  get(FCP_F_S0, phi, geometry::numTotal);
}
inline void global_sum(dp &x) { check(fcp_global_sum(ctx, &x), "fcp_global_sum"); }
inline void global_max(dp &x) { check(fcp_global_max(ctx, &x), "fcp_global_max"); }
inline void global_min(dp &x) { check(fcp_global_min(ctx, &x), "fcp_global_min"); }
inline void global_isum(int &i) {
  int64_t v = i;
  check(fcp_global_isum(ctx, &v), "fcp_global_isum");
  i = (int)v;
}

}  // namespace fcp
