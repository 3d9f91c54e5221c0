// fvm_scalar.cu -- tier "next" row f4: the scalar transport template (calcsc) the turbulence models are written in,
// fluxes/scalar_fluxes.f90:32-343 (facefluxsc, facefluxsc_periodic, facefluxsc_boundary) inside the assembly of
// TurbulenceModels/k_epsilon_rlzb.f90 (calcsc_tke :52-445, calcsc_epsilon :447-790, modify_mu_eff :792-975), plus
// fvExplicit/calc_strain_and_vorticity.f90.  Same design as fvm.cu / fvm_uvw.cu: one thread owns one cell and walks its faces in
// ascending face index (the order in which the reference's sequential loops touch that cell), every face quantity is evaluated in the
// face's own orientation on both sides, no atomics, no FMA contraction.
//   k_strain          calc_strain_and_vorticity
//   k_sc_assemble     volume sources (kind), bdf/bdf2 term, facefluxsc on inner faces, patches, wall treatment (kind)
//   k_sc_diag         a(diag) = sp - sum(off-diagonals) in CSR order, under-relaxation (:397-415)
//   k_clip            phi = max(phi, small) (:430)
//   k_mu_eff_cell / k_mu_eff_wall   modify_mu_eff (acos, cos, log: these agree with the reference's libm to rounding, not to the bit)
// Not built: Crank-Nicolson, buoyancy.  Partitioned meshes: a process face is a two-sided face with its ghost cell (phi, its gradient and vis are
// exchanged by the callers); the SST pair is single-GPU only (sigma comes from the face's owner, and both ranks see themselves as the owner).
#include "fcp_internal.h"
#include "fvm_common.cuh"
#include "interp.cuh"

#define FCP_CAPPA 0.41                          // parameters.f90:15
#define FCP_ELOG 8.432                          // :17
#define FCP_CTRANS ((double)11.63f)             // :18  `11.63` is a default-real literal
#define FCP_CMU 0.09                            // k_epsilon_rlzb.f90:16
#define FCP_C2RLZ 1.90                          // :18
#define FCP_A0RLZ ((double)4.04f)               // :22  `4.04`: default-real literal
__device__ __forceinline__ double cmu25_dev() { return sqrt(sqrt(FCP_CMU)); }                       // :25
__device__ __forceinline__ double cmu75_dev() { const double c = cmu25_dev(); return c * c * c; }    // :26

struct ScArgs {
  int kind, cscheme, tscheme, lowre;
  double gds, prtr, viscos, densit, timestep;
  const double *phi, *phio, *phioo, *te, *ed, *den, *vis, *visw, *dnw, *flmass, *u, *v, *w, *magStrain, *su_vol, *sp_vol, *g;
  const double *fsst, *walldist, *gte;    // k-omega SST: blending function F1, wall distance, gradient of k
  double *gen, *tau, *a, *su, *sp, *phi_new;
};
// k-omega SST constants, k_omega_SST.f90:21-40
#define SST_BETTAST 0.09
#define SST_SIGMK1 0.85
#define SST_SIGMK2 1.0
#define SST_SIGMOM1 0.5
#define SST_SIGMOM2 0.856
#define SST_BETAI1 0.075
#define SST_BETAI2 0.0828
#define SST_A1 0.31
#define SST_ALPHA1 (5.0 / 9.0)
#define SST_ALPHA2 0.44
__device__ __forceinline__ double p4_dev(double x) { return (x * x) * (x * x); }   // x**4
// 1/sigma of the SST equations from the blending function of ONE cell (the reference takes the owner's, k_omega_SST.f90:440-449)
template <int KIND> __device__ __forceinline__ double sst_prtr(double fs) {
  return KIND == 3 ? fs * SST_SIGMK1 + (1.0 - fs) * SST_SIGMK2 : fs * SST_SIGMOM1 + (1.0 - fs) * SST_SIGMOM2;
}

__global__ void __launch_bounds__(FCP_TPB) k_strain(int32_t n, const double *__restrict__ gU, const double *__restrict__ gV, const double *__restrict__ gW,
                                                     double *__restrict__ magStrain, double *__restrict__ vorticity) {
  FCP_CELL_LOOP(c, n) {
    const int64_t b = 3 * (int64_t)c;
    const double dudx = gU[b], dudy = gU[b + 1], dudz = gU[b + 2];
    const double dvdx = gV[b], dvdy = gV[b + 1], dvdz = gV[b + 2];
    const double dwdx = gW[b], dwdy = gW[b + 1], dwdz = gW[b + 2];
    const double s11 = dudx, s12 = 0.5 * (dudy + dvdx), s13 = 0.5 * (dudz + dwdx), s22 = dvdy, s23 = 0.5 * (dvdz + dwdy), s33 = dwdz;
    const double w12 = (dudy - dvdx), w13 = (dudz - dwdx), w23 = (dvdz - dwdy);
    magStrain[c] = sqrt(2 * (s11 * s11 + s22 * s22 + s33 * s33 + 2 * (s12 * s12 + s13 * s13 + s23 * s23)));
    vorticity[c] = sqrt(w12 * w12 + w23 * w23 + w13 * w13);
  }
}

// calcsc_epsilon overwrites ed(ijp) in wall cells WHILE it walks the patches (:726), so a periodic patch listed after a wall patch reads
// the imposed value, not the old one.  This returns ed(cell) as the reference's boundary loop holds it when it reaches face fp: the value
// imposed by the cell's last wall face that precedes fp, else the old value.
// the value the wall branch imposes in a wall cell: epsilon = cmu75 k^1.5/(cappa dnw) (k_epsilon_rlzb.f90:726), omega = sqrt(wvis^2 + wlog^2) (k_omega_SST.f90:673-677)
template <int KIND> __device__ __forceinline__ double wall_imposed_value(const ScArgs &g, int32_t cell, double dn) {
  if (KIND == 2) return cmu75_dev() * pow(g.te[cell], 1.5) / (FCP_CAPPA * dn);
  const double wlog = sqrt(g.te[cell]) / (cmu25_dev() * FCP_CAPPA * dn);
  const double wvis = 6.0 * (g.viscos / g.den[cell]) / (SST_BETAI1 * (dn * dn));
  return sqrt(wvis * wvis + wlog * wlog);
}
template <int KIND>
__device__ __forceinline__ double eps_value_at_face(const MeshView &m, const ScArgs &g, int32_t cell, int32_t fp, double old_value) {
  const int64_t base = m.slptr[cell >> 5] + (cell & 31);
  const int32_t len = m.len[cell];
  double v = old_value;
  for (int32_t q = 0; q < len; ++q) {
    const int32_t e = m.ent[base + (int64_t)q * 32], sl = m.slot[base + (int64_t)q * 32];
    const int32_t f = (e > 0 ? e : -e) - 1;
    if (sl == -1 - FCP_BC_WALL && f < fp) v = wall_imposed_value<KIND>(g, cell, g.dnw[m.n + (f - m.F)]);
  }
  return v;
}

// F1 of the SST model, k_omega_SST.f90:242-268 (omega call, before its sources)
__global__ void __launch_bounds__(FCP_TPB) k_sst_blend(int32_t n, double viscos, const double *__restrict__ walldist, const double *__restrict__ gte,
                                                        const double *__restrict__ gom, const double *__restrict__ den, const double *__restrict__ te,
                                                        const double *__restrict__ ed, double *__restrict__ fsst) {
  FCP_CELL_LOOP(c, n) {
    const int64_t b = 3 * (int64_t)c;
    const double wldist = walldist[c];
    const double dot = gte[b] * gom[b] + gte[b + 1] * gom[b + 1] + gte[b + 2] * gom[b + 2];
    const double domegapl = fmax(2 * SST_SIGMOM2 * den[c] / (ed[c]) * dot, FCP_SMALL);
    const double ksi = fmin(fmax(sqrt(te[c]) / (SST_BETTAST * wldist * ed[c] + FCP_SMALL), 500.0 * viscos / den[c] / (wldist * wldist * ed[c] + FCP_SMALL)),
                            4.0 * den[c] * te[c] * SST_SIGMOM2 / (domegapl * (wldist * wldist)));
    fsst[c] = tanh(p4_dev(ksi));
  }
}

// operands of one face, gathered for W faces at a time before the first use (k_sc_assemble)
struct ScCell { double xc, yc, zc, vol, phic, visc, denc, gc[3]; };
struct ScOps { double arx, ary, arz, xo, yo, zo, phio, viso, go[3], lambda, Df, fm, xf, yf, zf, fso; };
// e = 0: no face; two = the face is two-sided AND is evaluated in this pass (its 16 neighbour / face operands are needed)
template <int KIND>
__device__ __forceinline__ void sc_gather(const MeshView &m, const ScArgs &g, int32_t c, int32_t e, int32_t o, bool two, ScOps &q) {
  const int32_t f = (e > 0 ? e : -e) - 1;
  const bool on = e != 0;
  q.arx = on ? __ldg(m.arx + f) : 0.0; q.ary = on ? __ldg(m.ary + f) : 0.0; q.arz = on ? __ldg(m.arz + f) : 0.0;
  q.xo = two ? __ldg(m.xc + o) : 0.0; q.yo = two ? __ldg(m.yc + o) : 0.0; q.zo = two ? __ldg(m.zc + o) : 0.0;
  q.phio = two ? g.phi[o] : 0.0; q.viso = two ? g.vis[o] : 0.0;
  q.go[0] = two ? g.g[3 * (int64_t)o] : 0.0; q.go[1] = two ? g.g[3 * (int64_t)o + 1] : 0.0; q.go[2] = two ? g.g[3 * (int64_t)o + 2] : 0.0;
  q.lambda = two ? __ldg(m.facint + f) : 0.0; q.Df = two ? __ldg(m.Df + f) : 0.0; q.fm = two ? g.flmass[f] : 0.0;
  q.xf = two ? __ldg(m.xf + f) : 0.0; q.yf = two ? __ldg(m.yf + f) : 0.0; q.zf = two ? __ldg(m.zf + f) : 0.0;
  if (KIND == 3 || KIND == 4) {
    // 1/sigma comes from the face's OWNER cell; on a process face that is the peer's cell when the face is flipped against the unpartitioned mesh
    bool own = e > 0;
    if (two && own && m.proc_flip && f >= m.F) own = __ldg(m.proc_flip + (f - m.F)) == 0;
    q.fso = two ? g.fsst[own ? c : o] : 0.0;
  } else {
    q.fso = 0.0;
  }
}
// one face of cell c: e = signed entry, o = index across the face, sl = matrix slot (>= 0) or -1 - bctype; the reference's face body
template <int KIND>
__device__ __forceinline__ void sc_face(const MeshView &m, const ScArgs &g, int32_t c, const ScCell &cc, int32_t e, int32_t o, int32_t sl, const ScOps &q,
                                        double &s, double &p, double &genc, double &phin) {
  const double xc = cc.xc, yc = cc.yc, zc = cc.zc, vol = cc.vol, phic = cc.phic, visc = cc.visc, denc = cc.denc;
  const double gc[3] = {cc.gc[0], cc.gc[1], cc.gc[2]};
  (void)vol; (void)visc; (void)denc; (void)zc; (void)yc; (void)xc; (void)phic;
  {
  const int32_t f = (e > 0 ? e : -e) - 1;
  const double arx = q.arx, ary = q.ary, arz = q.arz;
  if (sl >= 0) {
    // ---- facefluxsc, scalar_fluxes.f90:32-141, in the face's orientation (P = owner, N = neighbour)
    const bool own = e > 0;
    const double xo = q.xo, yo = q.yo, zo = q.zo, phio_ = q.phio, viso = q.viso;
    const double go[3] = {q.go[0], q.go[1], q.go[2]};
    const double lambda = q.lambda, Df = q.Df, fm = q.fm;
    const double fxn = lambda, fxp = 1.0 - lambda;
    const double xP = own ? xc : xo, yP = own ? yc : yo, zP = own ? zc : zo, xN = own ? xo : xc, yN = own ? yo : yc, zN = own ? zo : zc;
    const double phiP = own ? phic : phio_, phiN = own ? phio_ : phic;
    const double visP = own ? visc : viso, visN = own ? viso : visc;
    const double gP[3] = {own ? gc[0] : go[0], own ? gc[1] : go[1], own ? gc[2] : go[2]};
    const double gN[3] = {own ? go[0] : gc[0], own ? go[1] : gc[1], own ? go[2] : gc[2]};
    const double viste = (visP + (visN - visP) * lambda) - g.viscos;
    const double prf = (KIND == 3 || KIND == 4) ? sst_prtr<KIND>(q.fso) : g.prtr;   // SST: sigma of the face's OWNER cell (gathered as fsst[own ? c : o])
    const double dcoef = g.viscos + viste * prf;
    const double xpn = xN - xP, ypn = yN - yP, zpn = zN - zP;
    const double de = dcoef * Df;
    const double ce = fmin(fm, 0.0), cp = fmax(fm, 0.0);
    const double can = -de + ce, cap = -de - cp;
    double dfixi = gP[0] * fxp + gN[0] * fxn, dfiyi = gP[1] * fxp + gN[1] * fxn, dfizi = gP[2] * fxp + gN[2] * fxn;
    dfixi = dfixi * (arx - Df * xpn); dfiyi = dfiyi * (ary - Df * ypn); dfizi = dfizi * (arz - Df * zpn);
    const double fdfie = dcoef * (dfixi + dfiyi + dfizi);
    const double xf = q.xf, yf = q.yf, zf = q.zf;
    double fii;
    if (fm >= 0.0) fii = face_value_dev(g.cscheme, phiP, phiN, gP, gN, xP, yP, zP, xN, yN, zN, xf, yf, zf, fxp);
    else fii = face_value_dev(g.cscheme, phiN, phiP, gN, gP, xN, yN, zN, xP, yP, zP, xf, yf, zf, fxn);
    double fcfie = fm * fii;
    const double fcfii = ce * phiN + cp * phiP;
    fcfie = g.gds * (fcfie - fcfii);
    const double suadd = -fcfie + fdfie;
    if (own) { g.a[sl] = can; s = s + suadd; }          // a(icell,jcell) = can ; su(ijp) += suadd
    else     { g.a[sl] = cap; s = s - suadd; }          // a(jcell,icell) = cap ; su(ijn) -= suadd
  } else {
    const int type = -1 - sl;
    if (type == FCP_BC_INLET || type == FCP_BC_OUTLET || type == FCP_BC_PRESSURE) {
      // ---- facefluxsc_boundary :236-300
      const double prf = (KIND == 3 || KIND == 4) ? sst_prtr<KIND>(g.fsst[c]) : g.prtr;
      const double viste = g.vis[o] - g.viscos, dcoef = g.viscos + viste * prf;
      const double xpn = m.xf[f] - xc, ypn = m.yf[f] - yc, zpn = m.zf[f] - zc;
      const double Dfi = (arx * arx + ary * ary + arz * arz) / (xpn * arx + ypn * ary + zpn * arz);
      const double de = dcoef * Dfi;
      const double ce = fmin(g.flmass[f], 0.0);
      const double can = -de + ce;
      double dfixi = gc[0], dfiyi = gc[1], dfizi = gc[2];
      dfixi = dfixi * (arx - Dfi * xpn); dfiyi = dfiyi * (ary - Dfi * ypn); dfizi = dfizi * (arz - Dfi * zpn);
      const double suadd = dcoef * (dfixi + dfiyi + dfizi);
      p = p - can;
      s = s - can * g.phi[o] + suadd;
    } else if (m.per_cell && (type == FCP_BC_PERIODIC || type == FCP_BC_EMPTY) && m.per_cell[f - m.F] >= 0) {
      // ---- facefluxsc_periodic :145-232 in the orientation of the PERIODIC face fp (quirk Q21: Df(i), i = ordinal in the patch)
      const int32_t b = f - m.F, q = m.per_cell[b];
      const bool own = type == FCP_BC_PERIODIC;
      const int32_t fp = own ? f : m.per_face[b];
      const double ax = m.arx[fp], ay = m.ary[fp], az = m.arz[fp];
      const double xo = m.xc[q], yo = m.yc[q], zo = m.zc[q], phio_ = g.phi[q], viso = g.vis[q];
      const double go[3] = {g.g[3 * (int64_t)q], g.g[3 * (int64_t)q + 1], g.g[3 * (int64_t)q + 2]};
      const double xP = own ? xc : xo, yP = own ? yc : yo, zP = own ? zc : zo;
      double phiP = own ? phic : phio_, phiN = own ? phio_ : phic;
      if (KIND == 2 || KIND == 4) {                      // wall cells already hold their imposed epsilon / omega when a later patch reads them
        phiP = eps_value_at_face<KIND>(m, g, own ? c : q, fp, phiP);
        phiN = eps_value_at_face<KIND>(m, g, own ? q : c, fp, phiN);
      }
      const double visP = own ? visc : viso, visN = own ? viso : visc;
      const double gP[3] = {own ? gc[0] : go[0], own ? gc[1] : go[1], own ? gc[2] : go[2]};
      const double gN[3] = {own ? go[0] : gc[0], own ? go[1] : gc[1], own ? go[2] : gc[2]};
      const double fxn = 0.5, fxp = fxn;
      const double prf = (KIND == 3 || KIND == 4) ? 0.5 * (sst_prtr<KIND>(g.fsst[own ? c : q]) + sst_prtr<KIND>(g.fsst[own ? q : c])) : g.prtr;
      const double viste = 0.5 * (visP + visN) - g.viscos, dcoef = g.viscos + viste * prf;
      const double xpn = 2 * (m.xf[fp] - xP), ypn = 2 * (m.yf[fp] - yP), zpn = 2 * (m.zf[fp] - zP);
      const double Dfq = m.per_df[b], fm = g.flmass[fp];
      const double de = dcoef * Dfq;
      const double ce = fmin(fm, 0.0), cp = fmax(fm, 0.0);
      const double can = -de + ce, cap = -de - cp;
      double dfixi = gP[0] * fxp + gN[0] * fxn, dfiyi = gP[1] * fxp + gN[1] * fxn, dfizi = gP[2] * fxp + gN[2] * fxn;
      dfixi = dfixi * (ax - Dfq * xpn); dfiyi = dfiyi * (ay - Dfq * ypn); dfizi = dfizi * (az - Dfq * zpn);
      const double fdfie = dcoef * (dfixi + dfiyi + dfizi);
      double fii;
      if (fm >= 0.0) fii = phiP + (phiN - phiP) * fxp; else fii = phiN + (phiP - phiN) * fxn;
      double fcfie = fm * fii;
      const double fcfii = ce * phiN + cp * phiP;
      fcfie = g.gds * (fcfie - fcfii);
      const double suadd = -fcfie + fdfie;
      if (own) { g.a[m.per_slot[b]] = can; s = s + suadd; }
      else     { g.a[m.per_slot[b]] = cap; s = s - suadd; }
    } else if (type == FCP_BC_WALL && (KIND == 1 || KIND == 3)) {
      // ---- wall function for k, k_epsilon_rlzb.f90:331-368: production from the wall shear stress replaces the standard one
      const double viss = fmax(g.viscos, g.visw[o]);
      const double are = sqrt(arx * arx + ary * ary + arz * arz);
      const double nxf = arx / are, nyf = ary / are, nzf = arz / are;
      const double uc = g.u[c], vc = g.v[c], wc = g.w[c];
      const double Vnp = uc * nxf + vc * nyf + wc * nzf;
      double xtp = uc - Vnp * nxf, ytp = vc - Vnp * nyf, ztp = wc - Vnp * nzf;
      const double Vtp = sqrt(xtp * xtp + ytp * ytp + ztp * ztp);
      xtp = xtp / Vtp; ytp = ytp / Vtp; ztp = ztp / Vtp;
      const double Ut2 = fabs((g.u[o] - uc) * xtp + (g.v[o] - vc) * ytp + (g.w[o] - wc) * ztp);
      const double dn = g.dnw[o];
      const double tau = viss * Ut2 / dn;
      g.tau[o] = tau;
      s = s - genc * vol;
      genc = fabs(tau) * cmu25_dev() * sqrt(phic) / (dn * FCP_CAPPA);
      s = s + genc * vol;
    } else if (type == FCP_BC_WALL && KIND == 2) {
      // ---- wall cells of the epsilon equation :712-728: the row is cleared, sp = 1, su = ed = cmu75 k^1.5/(cappa dnw)
      const int64_t base = m.a_slptr[c >> 5] + (c & 31);
      const int32_t len = m.a_rinfo[c] & 0xffff;
      for (int32_t k = 0; k < len; ++k) g.a[base + (int64_t)k * 32] = 0.0;
      p = 1.0;
      phin = wall_imposed_value<2>(g, c, g.dnw[o]);
      s = phin;
    } else if (type == FCP_BC_WALL && KIND == 4) {
      // ---- wall cells of the omega equation, k_omega_SST.f90:668-680
      const int64_t base = m.a_slptr[c >> 5] + (c & 31);
      const int32_t len = m.a_rinfo[c] & 0xffff;
      phin = wall_imposed_value<4>(g, c, g.dnw[o]);
      s = phin;
      for (int32_t k = 0; k < len; ++k) g.a[base + (int64_t)k * 32] = 0.0;
      p = 1.0;
    }
  }
  }
}

// Staged face lists (fvm_common.cuh) and gather rounds of W = 2 faces: the 19 operands of two faces are in flight together and the list of the
// next cell travels meanwhile, so a hexahedron costs three memory round trips instead of twelve.  Faces are evaluated one by one in list order
// (the wall branches overwrite what earlier faces wrote).  Cells with more faces than the stage holds walk the lists in global memory.
template <int KIND, int WS>
__global__ void __launch_bounds__(FCP_TPB, 2) k_sc_assemble(MeshView m, ScArgs g) {
  FCP_STAGE_DYN_N(WS, 2, stage);
  FCP_STAGED_LOOP_BEGIN(stage, m, m.n, c, st)
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c], vol = m.vol[c];
    const double phic = g.phi[c], visc = g.vis[c], denc = g.den[c];
    const double gc[3] = {g.g[3 * (int64_t)c], g.g[3 * (int64_t)c + 1], g.g[3 * (int64_t)c + 2]};
    double s, p, genc = 0.0, phin = phic;
    // ---- volume sources
    if (KIND == 0) { s = g.su_vol[c]; p = g.sp_vol[c]; }
    else if (KIND == 1) {                                   // calcsc_tke :103-120
      const double tec = phic, edc = g.ed[c];
      genc = fabs(visc - g.viscos) * g.magStrain[c] * g.magStrain[c];
      const double genp = fmax(genc, 0.0), genn = fmin(genc, 0.0);
      s = genp * vol;
      p = edc * denc * vol / (tec + FCP_SMALL);
      p = p - genn * vol / (tec + FCP_SMALL);
    } else if (KIND == 3) {                                 // k of the SST model, k_omega_SST.f90:143-170
      const double tec = phic, edc = g.ed[c];
      genc = fabs(visc - g.viscos) * g.magStrain[c] * g.magStrain[c];
      genc = fmin(genc, 0.9 * denc * tec * edc);
      if (g.lowre) {
        const double x4 = p4_dev(denc * tec / (8.0 * g.viscos * edc));
        const double tmp = 10 * SST_BETTAST * (4.0 / 15.0 + x4) / (1.0 + x4);
        genc = fmin(genc, tmp * denc * tec * edc);
      }
      const double genp = fmax(genc, 0.0), genn = fmin(genc, 0.0);
      s = genp * vol;
      p = SST_BETTAST * edc * denc * vol;
      if (g.lowre) {
        const double x4 = p4_dev(denc * tec / (8 * g.viscos * edc));
        const double tmp = SST_BETTAST * (4.0 / 15.0 + x4) / (1.0 + x4);
        p = tmp * edc * denc * vol;
      }
      p = p - genn * vol / (tec + FCP_SMALL);
    } else if (KIND == 4) {                                 // omega of the SST model :272-318 (gen: what the k call left behind)
      const double tec = g.te[c], edc = phic, fs = g.fsst[c];
      const double gn = g.gen[c];
      const double genp = fmax(gn, 0.0), genn = fmin(gn, 0.0);
      const double vist = (visc - g.viscos) / g.densit;
      double alphasst = fs * SST_ALPHA1 + (1.0 - fs) * SST_ALPHA2;
      if (g.lowre) {
        const double alphast = (0.024 + (g.densit * tec) / (6.0 * g.viscos * edc)) / (1.0 + (g.densit * tec) / (6.0 * g.viscos * edc));
        const double tmp = SST_ALPHA1 / alphast * (1.0 / 9.0 + (g.densit * tec) / (2.95 * g.viscos * edc)) / (1.0 + (g.densit * tec) / (2.95 * g.viscos * edc));
        alphasst = fs * tmp + (1.0 - fs) * SST_ALPHA2;
      }
      s = alphasst * genp * vol / (vist + FCP_SMALL);
      const int64_t b = 3 * (int64_t)c;
      const double dot = g.gte[b] * gc[0] + g.gte[b + 1] * gc[1] + g.gte[b + 2] * gc[2];
      double domega = 2 * (1.0 - fs) * denc * SST_SIGMOM2 / (edc + FCP_SMALL) * dot;
      domega = fmax(domega, 0.0);
      s = s + domega * vol;
      const double bettasst = fs * SST_BETAI1 + (1.0 - fs) * SST_BETAI2;
      p = bettasst * denc * edc * vol;
      p = p - alphasst * genn * vol / (vist * edc + FCP_SMALL);
    } else {                                                // calcsc_epsilon :500-513
      const double tec = g.te[c], edc = phic, ms = g.magStrain[c];
      const double genp = fmax(ms, 0.0), genn = fmin(ms, 0.0);
      const double etarlzb = ms * tec / (edc + FCP_SMALL);
      const double c1 = fmax((double)0.43f, etarlzb / (etarlzb + 5.0));
      s = c1 * genp * edc * vol;
      p = FCP_C2RLZ * denc * edc * vol / (tec + sqrt(g.viscos / g.densit * edc) + FCP_SMALL);
      p = p - c1 * genn * edc * vol;
    }
    if (g.tscheme) {                                        // :160-171
      const double apotime = denc * vol / g.timestep;
      if (g.tscheme == 1) { s = s + apotime * g.phio[c]; p = p + apotime; }
      else { s = s + apotime * (2 * g.phio[c] - 0.5 * g.phioo[c]); p = p + 1.5 * apotime; }
    }
    // On a partition the process faces sit BEHIND the physical patches in the cell's list, but they are inner faces of the unpartitioned mesh:
    // the reference's order is "inner faces, then the patches", and the wall branches of the epsilon / omega equations overwrite what the inner
    // faces wrote (row, su, sp).  Two passes restore that order: two-sided faces first, boundary faces second.
    const int npass = m.a_llen ? 2 : 1;
    const ScCell cc{xc, yc, zc, vol, phic, visc, denc, {gc[0], gc[1], gc[2]}};
    constexpr int W = 2;
    const int32_t flen = stage.len(st);
    if (flen <= WS) {
      for (int pass = 0; pass < npass; ++pass)
        for (int32_t q0 = 0; q0 < flen; q0 += W) {
          int32_t e_[W], o_[W], sl_[W];
          ScOps q_[W];
#pragma unroll
          for (int k = 0; k < W; ++k) {
            const bool in = q0 + k < flen;
            const int idx = in ? q0 + k : 0;
            sl_[k] = in ? stage.slot(st, idx, false) : -1;
            o_[k] = in ? stage.oth(st, idx) : 0;
            const bool act = in && !(npass == 2 && ((sl_[k] >= 0) != (pass == 0)));
            e_[k] = act ? stage.ent(st, idx) : 0;
            sc_gather<KIND>(m, g, c, e_[k], o_[k], act && sl_[k] >= 0, q_[k]);
          }
#pragma unroll
          for (int k = 0; k < W; ++k)
            if (e_[k] != 0) sc_face<KIND>(m, g, c, cc, e_[k], o_[k], sl_[k], q_[k], s, p, genc, phin);
        }
    } else {
      const int64_t fbase__ = m.slptr[c >> 5] + (c & 31);
      const int32_t flen__ = flen;
      for (int pass = 0; pass < npass; ++pass)
        for (int32_t q__ = 0; q__ < flen__; ++q__) {
          FCP_FACE_FETCH(m);
          if (npass == 2 && ((sl >= 0) != (pass == 0))) continue;
          ScOps q;
          sc_gather<KIND>(m, g, c, e, o, sl >= 0, q);
          sc_face<KIND>(m, g, c, cc, e, o, sl, q, s, p, genc, phin);
        }
    }
    g.su[c] = s;
    g.sp[c] = p;
    g.phi_new[c] = phin;
    if (KIND == 1 || KIND == 3) g.gen[c] = genc;
  FCP_STAGED_LOOP_END
}

// a(diag) = sp; a(diag) -= a(k) for the off-diagonals in CSR order; under-relaxation; phi <- the value the reference holds at this point
__global__ void __launch_bounds__(FCP_TPB) k_sc_diag(MeshView m, double *a, const double *__restrict__ sp, double *__restrict__ su,
                                                      const double *__restrict__ phi_new, double *__restrict__ phi, double urf) {
  const double urfrs = 1.0 / urf, urfms = 1.0 - urf;
  FCP_CELL_LOOP(c, m.n) {
    const int64_t base = m.a_slptr[c >> 5] + (c & 31);
    const int32_t ri = m.a_rinfo[c];
    const int32_t dpos = (ri >> 16) & 0xffff, len = ri & 0xffff;
    double ad = sp[c];
    for (int32_t k = 0; k < len; ++k) {
      if (k == dpos) continue;
      ad = ad - a[base + (int64_t)k * 32];
    }
    ad = ad * urfrs;
    a[base + (int64_t)dpos * 32] = ad;
    const double ph = phi_new[c];
    su[c] = su[c] + urfms * ad * ph;
    phi[c] = ph;
  }
}

__global__ void __launch_bounds__(FCP_TPB) k_clip(int32_t n, double *__restrict__ phi) {   // n = numCells (k-epsilon) or numTotal (SST, which clips the whole array)
  FCP_CELL_LOOP(c, n) { phi[c] = fmax(phi[c], FCP_SMALL); }
}

// modify_mu_eff, cell loop :810-880
__global__ void __launch_bounds__(FCP_TPB) k_mu_eff_cell(int32_t n, double urf, double viscos, const double *__restrict__ gU, const double *__restrict__ gV,
                                                          const double *__restrict__ gW, const double *__restrict__ te, const double *__restrict__ ed,
                                                          const double *__restrict__ den, double *__restrict__ vis) {
  FCP_CELL_LOOP(c, n) {
    const int64_t b = 3 * (int64_t)c;
    const double visold = vis[c];
    const double dudx = gU[b], dudy = gU[b + 1], dudz = gU[b + 2];
    const double dvdx = gV[b], dvdy = gV[b + 1], dvdz = gV[b + 2];
    const double dwdx = gW[b], dwdy = gW[b + 1], dwdz = gW[b + 2];
    const double s11 = dudx, s12 = 0.5 * (dudy + dvdx), s13 = 0.5 * (dudz + dwdx), s22 = dvdy, s23 = 0.5 * (dvdz + dwdy), s33 = dwdz;
    const double s21 = s12, s31 = s13, s32 = s23;
    const double w12 = 0.5 * (dudy - dvdx), w13 = 0.5 * (dudz - dwdx), w23 = 0.5 * (dvdz - dwdy);
    const double stild = sqrt(s11 * s11 + s22 * s22 + s33 * s33 + 2 * (s12 * s12 + s13 * s13 + s23 * s23));
    const double wrlzb = (s11 * s11 * s11 + s11 * s12 * s21 + s11 * s13 * s31 + s12 * s21 * s11 + s12 * s22 * s21 + s12 * s23 * s31 + s13 * s31 * s11 +
                          s13 * s32 * s21 + s13 * s33 * s31 + s21 * s11 * s12 + s21 * s12 * s22 + s21 * s13 * s32 + s22 * s21 * s12 + s22 * s22 * s22 +
                          s22 * s23 * s32 + s23 * s31 * s12 + s23 * s32 * s22 + s23 * s33 * s32 + s31 * s11 * s13 + s31 * s12 * s23 + s31 * s13 * s33 +
                          s32 * s21 * s13 + s32 * s22 * s23 + s32 * s23 * s33 + s33 * s31 * s13 + s33 * s32 * s23 + s33 * s33 * s33) /
                         (stild * stild * stild);
    const double ffi = S13 * acos(fmax(-1.0, fmin(sqrt(6.0) * wrlzb, 1.0)));
    const double ass = sqrt(6.0) * cos(ffi);
    const double ust = sqrt(s11 * s11 + s22 * s22 + s33 * s33 + 2 * (s12 * s12 + s13 * s13 + s23 * s23 + w12 * w12 + w13 * w13 + w23 * w23));
    const double cmur = 1.0 / (FCP_A0RLZ + ass * ust * te[c] / (ed[c] + FCP_SMALL));
    const double vist = den[c] * cmur * (te[c] * te[c]) / (ed[c] + FCP_SMALL);
    double v = viscos + vist;
    v = urf * v + (1.0 - urf) * visold;
    vis[c] = v;
  }
}
// modify_mu_eff, wall faces :888-947 (boundary-face parallel)
__global__ void __launch_bounds__(FCP_TPB) k_mu_eff_wall(MeshView m, const int32_t *__restrict__ bftype, double viscos, const double *__restrict__ te,
                                                          const double *__restrict__ den, const double *__restrict__ u, const double *__restrict__ v,
                                                          const double *__restrict__ w, const double *__restrict__ dnw, double *vis, double *visw,
                                                          double *ypl, double *tau) {
  FCP_CELL_LOOP(i, m.B) {
    if (bftype[i] != FCP_BC_WALL) continue;
    const int32_t f = m.F + i, ijp = m.owner[f], ijb = m.n + i;
    const double arx = m.arx[f], ary = m.ary[f], arz = m.arz[f];
    const double are = sqrt(arx * arx + ary * ary + arz * arz);
    const double nxf = arx / are, nyf = ary / are, nzf = arz / are;
    const double Vnp = u[ijp] * nxf + v[ijp] * nyf + w[ijp] * nzf;
    const double xtp = u[ijp] - Vnp * nxf, ytp = v[ijp] - Vnp * nyf, ztp = w[ijp] - Vnp * nzf;
    const double Vtp = sqrt(xtp * xtp + ytp * ytp + ztp * ztp);
    const double yp = den[ijp] * cmu25_dev() * sqrt(te[ijp]) * dnw[ijb] / viscos;
    ypl[ijb] = yp;
    tau[ijb] = FCP_CAPPA * den[ijp] * Vtp * cmu25_dev() * sqrt(te[ijp]) / log(FCP_ELOG * yp);
    double viscw = 0.0;
    if (yp > FCP_CTRANS) viscw = yp * viscos * FCP_CAPPA / log(FCP_ELOG * yp);
    const double vw = fmax(viscos, viscw);
    visw[ijb] = vw;
    vis[ijb] = vw;
  }
}

// ---------------------------------------------------------------------------------------------
// LES sub-grid viscosity (wale_sgs.f90, vremanSGS.f90): fvxGradient's Grad(U) (the two-pass Gauss gradient with the gradco skewness
// correction, fvxGradient.f90:1549-1662, 1761-1838) + the tensorFields algebra, one thread per cell.
// ---------------------------------------------------------------------------------------------
// one pass: gnew = (sum over faces of fie S)/vol with fie interpolated with the OLD gradient gold (zero in the first pass).  Staged lists and
// gather rounds of W = 3 faces like k_grad_gauss (fvm.cu): a hexahedron costs two memory round trips instead of twelve; the arithmetic per face
// and the face order are unchanged.
template <int WS>
__global__ void __launch_bounds__(FCP_TPB, 2) k_grad_gauss_fvx(MeshView m, const double *__restrict__ u, const double *__restrict__ gold,
                                                                double *__restrict__ gnew) {
  FCP_STAGE_DYN_N(WS, 2, stage);
  FCP_STAGED_LOOP_BEGIN(stage, m, m.n, c, st)
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c], uc = u[c];
    const double vol = __ldg(m.vol + c);
    const double ocx = gold ? gold[3 * (int64_t)c] : 0.0, ocy = gold ? gold[3 * (int64_t)c + 1] : 0.0, ocz = gold ? gold[3 * (int64_t)c + 2] : 0.0;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    constexpr int W = 3;
    FCP_FACE_BATCHES_STAGED(stage, st, m, c, W) {
      FCP_BATCH_LISTS_STAGED(stage, st, WS, m, W, e_, o_, sl_);
      double ax_[W], ay_[W], az_[W], lam_[W], fx_[W], fy_[W], fz_[W], xo_[W], yo_[W], zo_[W], uo_[W], gx_[W], gy_[W], gz_[W];
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const int32_t f = (e_[k] > 0 ? e_[k] : -e_[k]) - 1;
        const bool on = e_[k] != 0, two = on && sl_[k] >= 0;
        const int64_t o = o_[k];
        ax_[k] = on ? __ldg(m.arx + f) : 0.0; ay_[k] = on ? __ldg(m.ary + f) : 0.0; az_[k] = on ? __ldg(m.arz + f) : 0.0;
        uo_[k] = on ? __ldg(u + o) : 0.0;
        lam_[k] = two ? __ldg(m.facint + f) : 0.0;
        fx_[k] = two ? __ldg(m.xf + f) : 0.0; fy_[k] = two ? __ldg(m.yf + f) : 0.0; fz_[k] = two ? __ldg(m.zf + f) : 0.0;
        xo_[k] = two ? __ldg(m.xc + o) : 0.0; yo_[k] = two ? __ldg(m.yc + o) : 0.0; zo_[k] = two ? __ldg(m.zc + o) : 0.0;
        gx_[k] = (two && gold) ? __ldg(gold + 3 * o) : 0.0; gy_[k] = (two && gold) ? __ldg(gold + 3 * o + 1) : 0.0;
        gz_[k] = (two && gold) ? __ldg(gold + 3 * o + 2) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < W; ++k) {
        if (e_[k] == 0) continue;
        const double arx = ax_[k], ary = ay_[k], arz = az_[k];
        if (sl_[k] >= 0) {                                   // gradco, in the face's orientation (P = owner, N = neighbour)
          const bool own = e_[k] > 0;
          const double fxn = lam_[k], fxp = 1.0 - fxn;
          const double xo = xo_[k], yo = yo_[k], zo = zo_[k], uo = uo_[k];
          const double oox = gx_[k], ooy = gy_[k], ooz = gz_[k];
          const double xP = own ? xc : xo, yP = own ? yc : yo, zP = own ? zc : zo, xN = own ? xo : xc, yN = own ? yo : yc, zN = own ? zo : zc;
          const double uP = own ? uc : uo, uN = own ? uo : uc;
          const double gPx = own ? ocx : oox, gPy = own ? ocy : ooy, gPz = own ? ocz : ooz;
          const double gNx = own ? oox : ocx, gNy = own ? ooy : ocy, gNz = own ? ooz : ocz;
          const double xi = xP * fxp + xN * fxn, yi = yP * fxp + yN * fxn, zi = zP * fxp + zN * fxn;
          const double dfxi = gPx * fxp + gNx * fxn, dfyi = gPy * fxp + gNy * fxn, dfzi = gPz * fxp + gNz * fxn;
          const double fie = uP * fxp + uN * fxn + dfxi * (fx_[k] - xi) + dfyi * (fy_[k] - yi) + dfzi * (fz_[k] - zi);
          const double dfxe = fie * arx, dfye = fie * ary, dfze = fie * arz;
          if (own) { sx = sx + dfxe; sy = sy + dfye; sz = sz + dfze; }
          else     { sx = sx - dfxe; sy = sy - dfye; sz = sz - dfze; }
        } else {                                             // gradbc
          const double ub = uo_[k];
          sx = sx + ub * arx; sy = sy + ub * ary; sz = sz + ub * arz;
        }
      }
    }
    const double volr = 1.0 / vol;
    gnew[3 * (int64_t)c] = sx * volr; gnew[3 * (int64_t)c + 1] = sy * volr; gnew[3 * (int64_t)c + 2] = sz * volr;
  FCP_STAGED_LOOP_END
}

// tensors as t[9] = xx xy xz yx yy yz zx zy zz (tensorFields.f90)
__device__ __forceinline__ void tf_inner(const double (&a)[9], const double (&b)[9], double (&r)[9]) {   // :490-513; quirk Q24: r[6] uses a[6] twice
  r[0] = a[0] * b[0] + a[1] * b[3] + a[2] * b[6];
  r[1] = a[0] * b[1] + a[1] * b[4] + a[2] * b[7];
  r[2] = a[0] * b[2] + a[1] * b[5] + a[2] * b[8];
  r[3] = a[3] * b[0] + a[4] * b[3] + a[5] * b[6];
  r[4] = a[3] * b[1] + a[4] * b[4] + a[5] * b[7];
  r[5] = a[3] * b[2] + a[4] * b[5] + a[5] * b[8];
  r[6] = a[6] * b[0] + a[6] * b[3] + a[8] * b[6];
  r[7] = a[6] * b[1] + a[7] * b[4] + a[8] * b[7];
  r[8] = a[6] * b[2] + a[7] * b[5] + a[8] * b[8];
}
__device__ __forceinline__ void tf_trans(const double (&a)[9], double (&r)[9]) {
  r[0] = a[0]; r[1] = a[3]; r[2] = a[6]; r[3] = a[1]; r[4] = a[4]; r[5] = a[7]; r[6] = a[2]; r[7] = a[5]; r[8] = a[8];
}
__device__ __forceinline__ double tf_tr(const double (&a)[9]) { return a[0] + a[4] + a[8]; }
__device__ __forceinline__ double tf_magsq(const double (&a)[9]) {
  return a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3] + a[4] * a[4] + a[5] * a[5] + a[6] * a[6] + a[7] * a[7] + a[8] * a[8];
}
__device__ __forceinline__ void tf_symm(const double (&a)[9], double (&r)[9]) {
  double t[9];
  tf_trans(a, t);
#pragma unroll
  for (int k = 0; k < 9; ++k) r[k] = 0.5 * (a[k] + t[k]);
}
__device__ __forceinline__ void tf_dev(const double (&a)[9], double (&r)[9]) {
  const double tr = tf_tr(a), third = 1.0 / 3.0;
#pragma unroll
  for (int k = 0; k < 9; ++k) r[k] = a[k] - third * (tr * ((k == 0 || k == 4 || k == 8) ? 1.0 : 0.0));
}
#define FCP_TF_EPS ((double)1e-30f)     // tensorFields.f90:926  `1e-30`

template <int MODEL>   // 0 WALE (wale_sgs.f90:88-105), 1 Vreman (vremanSGS.f90:86-105)
__global__ void __launch_bounds__(FCP_TPB) k_sgs_viscosity(int32_t n, double urf, double viscos, const double *__restrict__ gU, const double *__restrict__ gV,
                                                            const double *__restrict__ gW, const double *__restrict__ den, const double *__restrict__ vol,
                                                            double *__restrict__ vis) {
  FCP_CELL_LOOP(c, n) {
    const int64_t b = 3 * (int64_t)c;
    const double D[9] = {gU[b], gU[b + 1], gU[b + 2], gV[b], gV[b + 1], gV[b + 2], gW[b], gW[b + 1], gW[b + 2]};
    double musgs;
    if (MODEL == 0) {
      const double Cw = (double)0.325f, r13 = 1.0 / 3.0;
      double DD[9], S[9], Sd[9], sD[9];
      tf_inner(D, D, DD);
      tf_symm(DD, S);
      tf_dev(S, Sd);
      const double magSqrSd = tf_magsq(Sd);
      tf_symm(D, sD);
      const double t = Cw * pow(vol[c], r13);
      const double num = (den[c] * (t * t)) * pow(magSqrSd, 1.5);
      const double dnm = pow(tf_magsq(sD), 2.5) + pow(magSqrSd, 1.25);
      musgs = num / (dnm + FCP_TF_EPS);
    } else {
      const double Cvsq = 0.0681, r23 = 2.0 / 3.0;
      double Dt[9], G[9], GG[9];
      tf_trans(D, Dt);
      tf_inner(Dt, D, G);
      tf_inner(G, G, GG);
      const double trG = tf_tr(G);
      const double x = trG * trG - tf_tr(GG);            // power(.tr.G, 2.0_dp): pow(x, 2) is exactly x*x
      double mu = sqrt((0.5 * x) / (tf_magsq(G) + FCP_TF_EPS));
      mu = fmax(mu, FCP_SMALL);
      musgs = ((den[c] * Cvsq) * pow(vol[c], r23)) * mu;
    }
    vis[c] = urf * (musgs + viscos) + (1.0 - urf) * vis[c];
  }
}
// boundary values of the effective viscosity, wale_sgs.f90:110-160: wall -> visw = vis = max(viscos, 0); periodic pair -> the mean (patch-order
// rule as in k_update_boundary); every other patch -> the owner value
__global__ void __launch_bounds__(FCP_TPB) k_sgs_boundary(MeshView m, const int32_t *__restrict__ bftype, double viscos, double *vis, double *visw) {
  FCP_CELL_LOOP(i, m.B) {
    const int t = bftype[i];
    const int32_t ijp = m.owner[m.F + i], ijb = m.n + i;
    if (t == FCP_BC_WALL) {
      const double vw = fmax(viscos, 0.0);
      visw[ijb] = vw;
      vis[ijb] = vw;
    } else if (t == FCP_BC_PERIODIC) {
      const double v = 0.5 * (vis[ijp] + vis[m.per_cell[i]]);
      vis[ijb] = v;
      if (m.per_face[i] < m.F + i) vis[m.n + (m.per_face[i] - m.F)] = v;
    } else if (t == FCP_BC_EMPTY && m.per_cell && m.per_cell[i] >= 0) {
      if (m.per_face[i] < m.F + i) vis[ijb] = vis[ijp];
    } else {
      vis[ijb] = vis[ijp];
    }
  }
}

// npass passes of the gradco gradient: pass 1 interpolates with a zero gradient, pass k with the gradient of pass k-1 (ghost copies exchanged in
// between on partitions); the last pass writes g, the passes alternate between g and gtmp.  npass = 2: fvxGradient.f90:1549-1662;
// npass = nigrad: the iterative Gauss gradient of the MPI tree, src-par/gradients.f90:1547-1664.
int fvm_grad_gauss_passes(fcp_ctx *ctx, const double *u, double *gtmp, double *g, int npass) {
  if (ctx->n == 0) return FCP_OK;
  MeshView m = fcp_mesh_view(ctx);
  FCP_TRY(fcp_apply_face_variant(ctx, fcp_face_variant(FCP_FK_GRAD_GAUSS), m, true, true));   // k_grad_gauss_fvx: same operands as k_grad_gauss + the face centres
  const bool wide = ctx->max_cell_faces > 6;
  const int grid = std::max(fcp_nchunks(ctx->n), 1);
  size_t smem6 = 0, smem10 = 0;
  if (wide) FCP_TRY((fcp_stage_smem<10, 2>(k_grad_gauss_fvx<10>, &smem10)));
  else FCP_TRY((fcp_stage_smem<6, 2>(k_grad_gauss_fvx<6>, &smem6)));
  size_t tok = ctx->prof.begin(FCP_K_GRAD, ctx->stream);
  const double *gold = nullptr;
  for (int lc = 1; lc <= npass; ++lc) {
    double *out = ((npass - lc) % 2 == 0) ? g : gtmp;
    if (wide) k_grad_gauss_fvx<10><<<grid, FCP_TPB, smem10, ctx->stream>>>(m, u, gold, out);
    else k_grad_gauss_fvx<6><<<grid, FCP_TPB, smem6, ctx->stream>>>(m, u, gold, out);
    FCP_LAUNCHED();
    if (lc != npass && ctx->comm) FCP_TRY(comm_exchange(ctx, out, 3));      // the next pass interpolates this pass's gradient across process faces
    gold = out;
  }
  ctx->prof.end(tok, ctx->stream);
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_grad_gauss_fvx(fcp_ctx *ctx, const double *u, double *gtmp, double *g) { return fvm_grad_gauss_passes(ctx, u, gtmp, g, 2); }
int fvm_sgs_viscosity(fcp_ctx *ctx, int model, double urf, double viscos, const double *gU, const double *gV, const double *gW, const double *den,
                      double *vis, double *visw) {
  if (ctx->n == 0) return FCP_OK;
  if (model == 0) k_sgs_viscosity<0><<<FCP_GRID(ctx->n)>>>(ctx->n, urf, viscos, gU, gV, gW, den, ctx->vol, vis);
  else k_sgs_viscosity<1><<<FCP_GRID(ctx->n)>>>(ctx->n, urf, viscos, gU, gV, gW, den, ctx->vol, vis);
  FCP_LAUNCHED();
  if (ctx->B) {
    k_sgs_boundary<<<FCP_GRID(ctx->B)>>>(fcp_mesh_view(ctx), ctx->bftype, viscos, vis, visw);
    FCP_LAUNCHED();
  }
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

// modify_mu_eff of the SST model, k_omega_SST.f90:790-958
__global__ void __launch_bounds__(FCP_TPB) k_mu_eff_sst_cell(int32_t n, double urf, double viscos, double densit, int lowre, const double *__restrict__ magStrain,
                                                              const double *__restrict__ walldist, const double *__restrict__ te, const double *__restrict__ ed,
                                                              const double *__restrict__ den, double *__restrict__ vis) {
  FCP_CELL_LOOP(c, n) {
    const double visold = vis[c], wldist = walldist[c], tec = te[c], edc = ed[c], denc = den[c];
    const double etha = fmax(2 * sqrt(tec) / (SST_BETTAST * wldist * edc), (500 * viscos / denc) / (wldist * wldist * edc));
    const double f2 = tanh(etha * etha);
    double v = viscos + denc * SST_A1 * tec / (fmax(SST_A1 * edc, magStrain[c] * f2));
    if (lowre) {
      const double alphast = (0.024 + (densit * tec) / (6 * viscos * edc)) / (1.0 + (densit * tec) / (6 * viscos * edc));
      v = viscos + denc * tec / (edc + FCP_SMALL) * 1.0 / fmax(1.0 / alphast, magStrain[c] * f2 / (SST_A1 * edc));
    }
    vis[c] = urf * v + (1.0 - urf) * visold;
  }
}
__global__ void __launch_bounds__(FCP_TPB) k_mu_eff_sst_wall(MeshView m, const int32_t *__restrict__ bftype, double viscos, double densit,
                                                              const double *__restrict__ te, const double *__restrict__ den, const double *__restrict__ u,
                                                              const double *__restrict__ v, const double *__restrict__ w, const double *__restrict__ dnw,
                                                              double *vis, double *visw, double *ypl, double *tau) {
  FCP_CELL_LOOP(i, m.B) {
    if (bftype[i] != FCP_BC_WALL) continue;
    const int32_t f = m.F + i, ijp = m.owner[f], ijb = m.n + i;
    const double arx = m.arx[f], ary = m.ary[f], arz = m.arz[f];
    const double are = sqrt(arx * arx + ary * ary + arz * arz);
    const double nxf = arx / are, nyf = ary / are, nzf = arz / are;
    const double Vnp = u[ijp] * nxf + v[ijp] * nyf + w[ijp] * nzf;
    const double xtp = u[ijp] - Vnp * nxf, ytp = v[ijp] - Vnp * nyf, ztp = w[ijp] - Vnp * nzf;
    const double Vtp = sqrt(xtp * xtp + ytp * ytp + ztp * ztp);
    const double dn = dnw[ijb];
    const double Utau = sqrt(viscos * Vtp / (densit * dn) + cmu25_dev() * te[ijp]);
    const double yp = den[ijp] * Utau * dn / viscos;
    ypl[ijb] = yp;
    const double Utaulog = 1.0 / FCP_CAPPA * log(FCP_ELOG * yp);
    const double uv2 = yp * yp, ul2 = Utaulog * Utaulog;
    const double Upl = sqrt(sqrt(uv2 * uv2 + ul2 * ul2));
    const double viscw = den[ijp] * Utau * dn / Upl;
    tau[ijb] = den[ijp] * ((Vtp / Upl) * (Vtp / Upl));
    const double vw = fmax(viscos, viscw);
    visw[ijb] = vw;
    vis[ijb] = vw;
  }
}
int fvm_mu_eff_sst(fcp_ctx *ctx, double urf, double viscos, double densit, int lowre, const double *magStrain, const double *walldist, const double *te,
                   const double *ed, const double *den, const double *u, const double *v, const double *w, const double *dnw, double *vis, double *visw,
                   double *ypl, double *tau) {
  if (ctx->n == 0) return FCP_OK;
  k_mu_eff_sst_cell<<<FCP_GRID(ctx->n)>>>(ctx->n, urf, viscos, densit, lowre, magStrain, walldist, te, ed, den, vis);
  FCP_LAUNCHED();
  FCP_TRY(fvm_update_boundary(ctx, vis));
  if (ctx->B) {
    k_mu_eff_sst_wall<<<FCP_GRID(ctx->B)>>>(fcp_mesh_view(ctx), ctx->bftype, viscos, densit, te, den, u, v, w, dnw, vis, visw, ypl, tau);
    FCP_LAUNCHED();
  }
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_sst_blend(fcp_ctx *ctx, double viscos, const double *walldist, const double *gte, const double *gom, const double *den, const double *te,
                  const double *ed, double *fsst) {
  if (ctx->n == 0) return FCP_OK;
  k_sst_blend<<<FCP_GRID(ctx->n)>>>(ctx->n, viscos, walldist, gte, gom, den, te, ed, fsst);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

int fvm_strain(fcp_ctx *ctx, const double *gU, const double *gV, const double *gW, double *magStrain, double *vorticity) {
  if (ctx->n == 0) return FCP_OK;
  k_strain<<<FCP_GRID(ctx->n)>>>(ctx->n, gU, gV, gW, magStrain, vorticity);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

int fvm_sc_assemble(fcp_ctx *ctx, const ScParams &q) {
  if (ctx->n == 0) return FCP_OK;
  ScArgs g;
  g.kind = q.kind; g.cscheme = q.cscheme; g.tscheme = q.tscheme;
  g.gds = q.gds; g.prtr = q.prtr; g.viscos = q.viscos; g.densit = q.densit; g.timestep = q.timestep;
  g.phi = q.phi; g.phio = q.phio; g.phioo = q.phioo; g.te = q.te; g.ed = q.ed; g.den = q.den; g.vis = q.vis; g.visw = q.visw; g.dnw = q.dnw;
  g.flmass = q.flmass; g.u = q.u; g.v = q.v; g.w = q.w; g.magStrain = q.magStrain; g.su_vol = q.su_vol; g.sp_vol = q.sp_vol; g.g = q.grad;
  g.gen = q.gen; g.tau = q.tau; g.a = q.a; g.su = q.su; g.sp = q.sp; g.phi_new = q.phi_new;
  g.fsst = q.fsst; g.walldist = q.walldist; g.gte = q.gte; g.lowre = q.lowre;
  const MeshView m = fcp_mesh_view(ctx);
  size_t tok = ctx->prof.begin(FCP_K_SCALAR, ctx->stream);
  {
    const int grid = std::max(fcp_nchunks(ctx->n), 1);
    size_t smem = 0;
#define SC_LAUNCH(KIND)                                                                                   \
  do {                                                                                                    \
    if (ctx->max_cell_faces > 6) {                                                                        \
      FCP_TRY((fcp_stage_smem<10, 2>(k_sc_assemble<KIND, 10>, &smem)));                                   \
      k_sc_assemble<KIND, 10><<<grid, FCP_TPB, smem, ctx->stream>>>(m, g);                                \
    } else {                                                                                              \
      FCP_TRY((fcp_stage_smem<6, 2>(k_sc_assemble<KIND, 6>, &smem)));                                     \
      k_sc_assemble<KIND, 6><<<grid, FCP_TPB, smem, ctx->stream>>>(m, g);                                 \
    }                                                                                                     \
  } while (0)
    if (q.kind == 0) SC_LAUNCH(0);
    else if (q.kind == 1) SC_LAUNCH(1);
    else if (q.kind == 2) SC_LAUNCH(2);
    else if (q.kind == 3) SC_LAUNCH(3);
    else SC_LAUNCH(4);
#undef SC_LAUNCH
  }
  ctx->prof.end(tok, ctx->stream);
  FCP_LAUNCHED();
  k_sc_diag<<<FCP_GRID(ctx->n)>>>(m, q.a, q.sp, q.su, q.phi_new, q.phi_out, q.urf);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

int fvm_clip_small(fcp_ctx *ctx, double *phi, int32_t count) {
  if (count == 0) return FCP_OK;
  k_clip<<<FCP_GRID(count)>>>(count, phi);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

int fvm_mu_eff_rlzb(fcp_ctx *ctx, double urf, double viscos, const double *gU, const double *gV, const double *gW, const double *te, const double *ed,
                    const double *den, const double *u, const double *v, const double *w, const double *dnw, double *vis, double *visw, double *ypl,
                    double *tau) {
  if (ctx->n == 0) return FCP_OK;
  k_mu_eff_cell<<<FCP_GRID(ctx->n)>>>(ctx->n, urf, viscos, gU, gV, gW, te, ed, den, vis);
  FCP_LAUNCHED();
  FCP_TRY(fvm_update_boundary(ctx, vis));                                       // :884
  if (ctx->B) {
    k_mu_eff_wall<<<FCP_GRID(ctx->B)>>>(fcp_mesh_view(ctx), ctx->bftype, viscos, te, den, u, v, w, dnw, vis, visw, ypl, tau);
    FCP_LAUNCHED();
  }
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
