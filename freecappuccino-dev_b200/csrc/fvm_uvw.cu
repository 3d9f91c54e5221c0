// fvm_uvw.cu -- the momentum predictor calcuvw (Velocity/velocity.f90:50-750), tier "next" row f1, as cell-centric gathers
// like fvm.cu:
//   k_update_vel_bnd   updateVelocityAtBoundary                      velocity.f90:1184-1277
//   k_uvw_assemble     volume sources :187-283, facefluxuvw :754-878 (+ sngrad gradients.f90:1720-1779 and the face_value
//                      family interpolation.f90:28-650), boundary patches :330-475 (facefluxuvw_bnd :882-1034)
//   k_uvw_diag         diagonal, reciprocal diagonal apu/apv/apw and under-relaxation :602-620, :656-668, :717-730
// One thread owns one cell and walks its faces in ascending index; every face quantity is evaluated in the face's own
// orientation (P = owner, N = neighbour), so both sides see the same bits and the per-cell sums round like the
// reference's sequential loops.  Periodic patches: facefluxuvw_periodic on both cells of a pair.  Not built: Crank-Nicolson, buoyancy, MHD.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "fcp_internal.h"
#include "fvm_common.cuh"
#include "interp.cuh"

struct CellState {   // what facefluxuvw reads of one cell
  double x, y, z, u, v, w, vis;
  double gu[3], gv[3], gw[3];
};

// sngrad for one component, gradients.f90:1720-1779
__device__ __forceinline__ void sngrad_dev(double arx, double ary, double arz, double fxp, double fxn, double xpn, double ypn, double zpn, double Df,
                                           double phiP, double phiN, const double (&gP)[3], const double (&gN)[3], double &d1, double &d2, double &d3,
                                           double &e1, double &e2, double &e3) {
  const double vole = xpn * arx + ypn * ary + zpn * arz;
  d1 = gP[0] * fxp + gN[0] * fxn;
  d2 = gP[1] * fxp + gN[1] * fxn;
  d3 = gP[2] * fxp + gN[2] * fxn;
  e1 = d1 + arx / vole * (phiN - phiP - d1 * xpn - d2 * ypn - d3 * zpn);
  e2 = d2 + ary / vole * (phiN - phiP - d1 * xpn - d2 * ypn - d3 * zpn);
  e3 = d3 + arz / vole * (phiN - phiP - d1 * xpn - d2 * ypn - d3 * zpn);
  d1 = d1 * (arx - Df * xpn);
  d2 = d2 * (ary - Df * ypn);
  d3 = d3 * (arz - Df * zpn);
}

__device__ __forceinline__ void load_cell(CellState &s, const MeshView &m, const UvwArgs &g, int32_t c) {
  s.x = m.xc[c]; s.y = m.yc[c]; s.z = m.zc[c];
  s.u = g.u[c]; s.v = g.v[c]; s.w = g.w[c]; s.vis = g.vis[c];
#pragma unroll
  for (int k = 0; k < 3; ++k) { s.gu[k] = g.dUdxi[3 * (int64_t)c + k]; s.gv[k] = g.dVdxi[3 * (int64_t)c + k]; s.gw[k] = g.dWdxi[3 * (int64_t)c + k]; }
}

// updateVelocityAtBoundary, velocity.f90:1184-1277 (boundary-face parallel)
__global__ void __launch_bounds__(FCP_TPB) k_update_vel_bnd(MeshView m, const int32_t *__restrict__ bftype, double *u, double *v, double *w) {
  FCP_CELL_LOOP(i, m.B) {
    const int t = bftype[i];
    const int32_t f = m.F + i, ijp = m.owner[f], ijb = m.n + i;
    if (t == FCP_BC_EMPTY || t == FCP_BC_PERIODIC) {
      u[ijb] = u[ijp]; v[ijb] = v[ijp]; w[ijb] = w[ijp];
    } else if (t == FCP_BC_SYMMETRY) {
      const double sx = m.arx[f], sy = m.ary[f], sz = m.arz[f];
      const double Unmag = u[ijp] * sx + v[ijp] * sy + w[ijp] * sz;
      u[ijb] = u[ijp] - Unmag * sx; v[ijb] = v[ijp] - Unmag * sy; w[ijb] = w[ijp] - Unmag * sz;
    }
  }
}

// Face stage: everything facefluxuvw reads of ONE two-sided face and of the cell across it (25 doubles), copied global -> shared with cp.async
// one face ahead of the face being evaluated, so that the ~300 flop of a face overlap the gathers of the next one.  Two stages of [25][256] doubles.
// Together with the list stage (fvm_common.cuh) a cell costs one exposed memory round trip (its first face) instead of twelve.
#define FCP_UVW_NV 25
struct UvwFaceStage {
  double v[2][FCP_UVW_NV][FCP_TPB];
  __device__ __forceinline__ void fetch(int s, const MeshView &m, const UvwArgs &g, int32_t e, int32_t o, int32_t sl) {
    const int t = threadIdx.x;
    if (e == 0) return;
    const int32_t f = (e > 0 ? e : -e) - 1;
    fcp_cp_async8(&v[s][0][t], m.arx + f); fcp_cp_async8(&v[s][1][t], m.ary + f); fcp_cp_async8(&v[s][2][t], m.arz + f);
    if (sl < 0) return;
    fcp_cp_async8(&v[s][3][t], m.xf + f); fcp_cp_async8(&v[s][4][t], m.yf + f); fcp_cp_async8(&v[s][5][t], m.zf + f);
    fcp_cp_async8(&v[s][6][t], g.flmass + f); fcp_cp_async8(&v[s][7][t], m.facint + f); fcp_cp_async8(&v[s][8][t], m.Df + f);
    fcp_cp_async8(&v[s][9][t], m.xc + o); fcp_cp_async8(&v[s][10][t], m.yc + o); fcp_cp_async8(&v[s][11][t], m.zc + o);
    fcp_cp_async8(&v[s][12][t], g.u + o); fcp_cp_async8(&v[s][13][t], g.v + o); fcp_cp_async8(&v[s][14][t], g.w + o);
    fcp_cp_async8(&v[s][15][t], g.vis + o);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      fcp_cp_async8(&v[s][16 + k][t], g.dUdxi + 3 * (int64_t)o + k);
      fcp_cp_async8(&v[s][19 + k][t], g.dVdxi + 3 * (int64_t)o + k);
      fcp_cp_async8(&v[s][22 + k][t], g.dWdxi + 3 * (int64_t)o + k);
    }
  }
  __device__ __forceinline__ void cell(int s, CellState &c) const {
    const int t = threadIdx.x;
    c.x = v[s][9][t]; c.y = v[s][10][t]; c.z = v[s][11][t];
    c.u = v[s][12][t]; c.v = v[s][13][t]; c.w = v[s][14][t]; c.vis = v[s][15][t];
#pragma unroll
    for (int k = 0; k < 3; ++k) { c.gu[k] = v[s][16 + k][t]; c.gv[k] = v[s][19 + k][t]; c.gw[k] = v[s][22 + k][t]; }
  }
};

// FS: face stage on (needs 138 KB of shared memory: one CTA per SM); MINB: resident CTAs per SM the registers are allocated for
template <bool FS, int MINB>
__global__ void __launch_bounds__(FCP_TPB, MINB) k_uvw_assemble(MeshView m, UvwArgs g) {
  constexpr int WS = 6;
#ifdef FCP_EMU
  unsigned char *raw__ = emu::dyn_smem();
#else
  extern __shared__ __align__(16) unsigned char raw__[];
#endif
  ListStage<WS> &stage = *reinterpret_cast<ListStage<WS> *>(raw__);
  UvwFaceStage &fs = *reinterpret_cast<UvwFaceStage *>(raw__ + sizeof(ListStage<WS>));
  int fst = 0;               // face stage of the face about to be evaluated
  bool have_pref = false;    // its copies were issued while the previous cell was being finished
  FCP_STAGED_LOOP_BEGIN(stage, m, m.n, c, st)
    CellState me;
    load_cell(me, m, g, c);
    const double vol = m.vol[c], denc = g.den[c];
    double s1 = g.su[c], s2 = g.sv[c], s3 = g.sw[c];          // -sum p_f S_f from gradp_and_sources
    double p1 = 0.0, p2 = 0.0, p3 = 0.0;                      // spu, spv, sp (:130-132)
    // ---- volume sources, velocity.f90:187-283
    if (g.const_mflux) s1 = s1 + g.gradPcmf * vol;
    if (g.tscheme == 1) {
      const double apotime = denc * vol / g.timestep;
      s1 = s1 + apotime * g.uo[c]; s2 = s2 + apotime * g.vo[c]; s3 = s3 + apotime * g.wo[c];
      p1 = p1 + apotime; p2 = p2 + apotime; p3 = p3 + apotime;
    } else if (g.tscheme == 2) {
      const double apotime = denc * vol / g.timestep;
      s1 = s1 + apotime * (2 * g.uo[c] - 0.5 * g.uoo[c]);
      s2 = s2 + apotime * (2 * g.vo[c] - 0.5 * g.voo[c]);
      s3 = s3 + apotime * (2 * g.wo[c] - 0.5 * g.woo[c]);
      p1 = p1 + 1.5 * apotime; p2 = p2 + 1.5 * apotime; p3 = p3 + 1.5 * apotime;
    } else if (g.tscheme == 3) {
      const double apotime = denc * vol / g.timestep;
      const double third = (double)1.f / 3.0, c116 = (double)11.f / 6.0;
      s1 = s1 + apotime * (3 * g.uo[c] - 1.5 * g.uoo[c] + third * g.uooo[c]);
      s2 = s2 + apotime * (3 * g.vo[c] - 1.5 * g.voo[c] + third * g.vooo[c]);
      s3 = s3 + apotime * (3 * g.wo[c] - 1.5 * g.woo[c] + third * g.wooo[c]);
      p1 = p1 + c116 * apotime; p2 = p2 + c116 * apotime; p3 = p3 + c116 * apotime;
    }
    const int32_t flen = stage.len(st);
    const bool lstaged = flen <= WS;            // the list comes from the list stage
    const bool staged = FS && lstaged;          // ... and the face operands from the face stage
    int32_t e_[WS], o_[WS], sl_[WS];
    if (lstaged) stage.read(st, e_, o_, sl_);
    if (staged && !have_pref) { fs.fetch(fst, m, g, e_[0], o_[0], sl_[0]); fcp_cp_async_commit(); }
    const int64_t fbase = lstaged ? 0 : m.slptr[c >> 5] + (c & 31);
    for (int32_t q = 0; q < flen; ++q) {
      int32_t e, o, sl;
      if (staged) {
        e = 0; o = 0; sl = -1;
#pragma unroll
        for (int k = 0; k < WS; ++k) if (k == q) { e = e_[k]; o = o_[k]; sl = sl_[k]; }
        // the next face to be evaluated by this thread: face q + 1 of this cell, or the first face of its next cell (whose list is already staged)
        int32_t ne = 0, no = 0, nsl = -1;
        if (q + 1 < flen) {
#pragma unroll
          for (int k = 0; k < WS; ++k) if (k == q + 1) { ne = e_[k]; no = o_[k]; nsl = sl_[k]; }
        } else if (j__ + 1 < FCP_IPT && fcp_chunk_cell(j__ + 1) < m.n && stage.len(st ^ 1) <= WS && stage.len(st ^ 1) > 0) {
          ne = stage.v[st ^ 1][1][threadIdx.x]; no = stage.v[st ^ 1][1 + WS][threadIdx.x]; nsl = stage.v[st ^ 1][1 + 2 * WS][threadIdx.x];
        }
        fs.fetch(fst ^ 1, m, g, ne, no, nsl);
        fcp_cp_async_commit();
        fcp_cp_async_wait<1>();      // everything but the group just committed has landed: this face's data is in stage fst
        have_pref = ne != 0 && !(q + 1 < flen);
      } else if (lstaged) {
        e = 0; o = 0; sl = -1;
#pragma unroll
        for (int k = 0; k < WS; ++k) if (k == q) { e = e_[k]; o = o_[k]; sl = sl_[k]; }
      } else {
        e = __ldcs(m.ent + fbase + (int64_t)q * 32); o = __ldcs(m.other + fbase + (int64_t)q * 32); sl = __ldcs(m.slot + fbase + (int64_t)q * 32);
        have_pref = false;
      }
      const int32_t f = (e > 0 ? e : -e) - 1;
      const int t__ = threadIdx.x;
      const double arx = staged ? fs.v[fst][0][t__] : m.arx[f], ary = staged ? fs.v[fst][1][t__] : m.ary[f], arz = staged ? fs.v[fst][2][t__] : m.arz[f];
      if (sl >= 0) {
        // ---- facefluxuvw, velocity.f90:754-878
        CellState ot;
        if (staged) fs.cell(fst, ot);
        else load_cell(ot, m, g, o);
        const bool own = e > 0;
        const CellState &P = own ? me : ot, &N = own ? ot : me;
        const double xf = staged ? fs.v[fst][3][t__] : m.xf[f], yf = staged ? fs.v[fst][4][t__] : m.yf[f], zf = staged ? fs.v[fst][5][t__] : m.zf[f];
        const double flomass = staged ? fs.v[fst][6][t__] : g.flmass[f], lambda = staged ? fs.v[fst][7][t__] : m.facint[f], Df = staged ? fs.v[fst][8][t__] : m.Df[f];
        const double fxn = lambda, fxp = 1.0 - lambda;
        const double game = P.vis + (N.vis - P.vis) * lambda;
        const double de = game * Df;
        const double ce = fmin(flomass, 0.0), cp = fmax(flomass, 0.0);
        const double can = -de + ce, cap = -de - cp;
        const double xpn = N.x - P.x, ypn = N.y - P.y, zpn = N.z - P.z;
        double duxi, duyi, duzi, duxii, duyii, duzii, dvxi, dvyi, dvzi, dvxii, dvyii, dvzii, dwxi, dwyi, dwzi, dwxii, dwyii, dwzii;
        sngrad_dev(arx, ary, arz, fxp, fxn, xpn, ypn, zpn, Df, P.u, N.u, P.gu, N.gu, duxi, duyi, duzi, duxii, duyii, duzii);
        sngrad_dev(arx, ary, arz, fxp, fxn, xpn, ypn, zpn, Df, P.v, N.v, P.gv, N.gv, dvxi, dvyi, dvzi, dvxii, dvyii, dvzii);
        sngrad_dev(arx, ary, arz, fxp, fxn, xpn, ypn, zpn, Df, P.w, N.w, P.gw, N.gw, dwxi, dwyi, dwzi, dwxii, dwyii, dwzii);
        double fdue = game * (duxii * arx + dvxii * ary + dwxii * arz);
        double fdve = game * (duyii * arx + dvyii * ary + dwyii * arz);
        double fdwe = game * (duzii * arx + dvzii * ary + dwzii * arz);
        const double fdui = game * (duxi + duyi + duzi), fdvi = game * (dvxi + dvyi + dvzi), fdwi = game * (dwxi + dwyi + dwzi);
        fdue = fdue + fdui; fdve = fdve + fdvi; fdwe = fdwe + fdwi;
        const double fuuds = cp * P.u + ce * N.u, fvuds = cp * P.v + ce * N.v, fwuds = cp * P.w + ce * N.w;
        double ue, ve, we;
        if (flomass >= 0.0) {
          ue = face_value_dev(g.cscheme, P.u, N.u, P.gu, N.gu, P.x, P.y, P.z, N.x, N.y, N.z, xf, yf, zf, fxp);
          ve = face_value_dev(g.cscheme, P.v, N.v, P.gv, N.gv, P.x, P.y, P.z, N.x, N.y, N.z, xf, yf, zf, fxp);
          we = face_value_dev(g.cscheme, P.w, N.w, P.gw, N.gw, P.x, P.y, P.z, N.x, N.y, N.z, xf, yf, zf, fxp);
        } else {
          ue = face_value_dev(g.cscheme, N.u, P.u, N.gu, P.gu, N.x, N.y, N.z, P.x, P.y, P.z, xf, yf, zf, fxn);
          ve = face_value_dev(g.cscheme, N.v, P.v, N.gv, P.gv, N.x, N.y, N.z, P.x, P.y, P.z, xf, yf, zf, fxn);
          we = face_value_dev(g.cscheme, N.w, P.w, N.gw, P.gw, N.x, N.y, N.z, P.x, P.y, P.z, xf, yf, zf, fxn);
        }
        const double fuhigh = flomass * ue, fvhigh = flomass * ve, fwhigh = flomass * we;
        const double sup = -g.gds * (fuhigh - fuuds) + fdue;
        const double svp = -g.gds * (fvhigh - fvuds) + fdve;
        const double swp = -g.gds * (fwhigh - fwuds) + fdwe;
        if (own) { g.a[sl] = can; s1 = s1 + sup; s2 = s2 + svp; s3 = s3 + swp; }   // a(icell,jcell) = can ; su(ijp) += sup
        else     { g.a[sl] = cap; s1 = s1 - sup; s2 = s2 - svp; s3 = s3 - swp; }   // a(jcell,icell) = cap ; su(ijn) -= sup
      } else {
        const int type = -1 - sl;
        if (type == FCP_BC_INLET || type == FCP_BC_OUTLET || type == FCP_BC_PRESSURE) {
          // ---- facefluxuvw_bnd, velocity.f90:882-1034 (the callee's `can` is the caller's cb)
          const double xpn = m.xf[f] - me.x, ypn = m.yf[f] - me.y, zpn = m.zf[f] - me.z;
          const double vole = xpn * arx + ypn * ary + zpn * arz;
          const double Dfi = (arx * arx + ary * ary + arz * arz) / vole;
          const double game = g.vis[o];
          const double de = game * Dfi;
          const double cb = -de + fmin(g.flmass[f], 0.0);
          const double ub = g.u[o], vb = g.v[o], wb = g.w[o];
          double e1[3], e2[3], e3[3], d1[3], d2[3], d3[3];
          const double phb[3] = {ub, vb, wb}, php[3] = {me.u, me.v, me.w};
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const double gx = q == 0 ? me.gu[0] : q == 1 ? me.gv[0] : me.gw[0];
            const double gy = q == 0 ? me.gu[1] : q == 1 ? me.gv[1] : me.gw[1];
            const double gz = q == 0 ? me.gu[2] : q == 1 ? me.gv[2] : me.gw[2];
            e1[q] = gx + arx / vole * (phb[q] - php[q] - gx * xpn - gy * ypn - gz * zpn);
            e2[q] = gy + ary / vole * (phb[q] - php[q] - gx * xpn - gy * ypn - gz * zpn);
            e3[q] = gz + arz / vole * (phb[q] - php[q] - gx * xpn - gy * ypn - gz * zpn);
            d1[q] = gx * (arx - Dfi * xpn); d2[q] = gy * (ary - Dfi * ypn); d3[q] = gz * (arz - Dfi * zpn);
          }
          const double fdue = game * (e1[0] * arx + e1[1] * ary + e1[2] * arz);
          const double fdve = game * (e2[0] * arx + e2[1] * ary + e2[2] * arz);
          const double fdwe = game * (e3[0] * arx + e3[1] * ary + e3[2] * arz);
          const double fdui = game * (d1[0] + d2[0] + d3[0]), fdvi = game * (d1[1] + d2[1] + d3[1]), fdwi = game * (d1[2] + d2[2] + d3[2]);
          const double sup = fdue + fdui, svp = fdve + fdvi, swp = fdwe + fdwi;
          p1 = p1 - cb; p2 = p2 - cb; p3 = p3 - cb;
          s1 = s1 - cb * ub + sup;
          s2 = s2 - cb * vb + svp;
          s3 = s3 - cb * wb + swp;
        } else if (m.per_cell && (type == FCP_BC_PERIODIC || type == FCP_BC_EMPTY) && m.per_cell[f - m.F] >= 0) {
          // ---- facefluxuvw_periodic, velocity.f90:393-432 + :1038-1180: evaluated in the orientation of the PERIODIC face fp
          // (P = its owner, N = the owner of the twin face) on both sides of the pair
          const int32_t b = f - m.F, q = m.per_cell[b];
          const bool own = type == FCP_BC_PERIODIC;
          const int32_t fp = own ? f : m.per_face[b];
          CellState ot;
          load_cell(ot, m, g, q);
          const CellState &P = own ? me : ot, &N = own ? ot : me;
          const double ax = m.arx[fp], ay = m.ary[fp], az = m.arz[fp];
          const double fxn = 0.5, fxp = fxn;
          const double xpn = 2 * (m.xf[fp] - P.x), ypn = 2 * (m.yf[fp] - P.y), zpn = 2 * (m.zf[fp] - P.z);
          const double dpn = sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
          const double are = sqrt(ax * ax + ay * ay + az * az);
          const double game = P.vis * fxp + N.vis * fxn;
          const double de = game * (ax * ax + ay * ay + az * az) / (xpn * ax + ypn * ay + zpn * az);
          const double flomass = g.flmass[fp];
          const double ce = fmin(flomass, 0.0), cp = fmax(flomass, 0.0);
          const double can = -de + ce, cap = -de - cp;
          const double duxi = P.gu[0] * fxp + N.gu[0] * fxn, duyi = P.gu[1] * fxp + N.gu[1] * fxn, duzi = P.gu[2] * fxp + N.gu[2] * fxn;
          const double dvxi = P.gv[0] * fxp + N.gv[0] * fxn, dvyi = P.gv[1] * fxp + N.gv[1] * fxn, dvzi = P.gv[2] * fxp + N.gv[2] * fxn;
          const double dwxi = P.gw[0] * fxp + N.gw[0] * fxn, dwyi = P.gw[1] * fxp + N.gw[1] * fxn, dwzi = P.gw[2] * fxp + N.gw[2] * fxn;
          const double fdue = game * ((duxi + duxi) * ax + (duyi + dvxi) * ay + (duzi + dwxi) * az);
          const double fdve = game * ((duyi + dvxi) * ax + (dvyi + dvyi) * ay + (dvzi + dwyi) * az);
          const double fdwe = game * ((duzi + dwxi) * ax + (dwyi + dvzi) * ay + (dwzi + dwzi) * az);
          const double fdui = game * are / dpn * (duxi * xpn + duyi * ypn + duzi * zpn);
          const double fdvi = game * are / dpn * (dvxi * xpn + dvyi * ypn + dvzi * zpn);
          const double fdwi = game * are / dpn * (dwxi * xpn + dwyi * ypn + dwzi * zpn);
          const double fuuds = cp * P.u + ce * N.u, fvuds = cp * P.v + ce * N.v, fwuds = cp * P.w + ce * N.w;
          double ue, ve, we;                                   // face_value_cds with lambda = half, interpolation.f90:116-157
          if (flomass >= 0.0) { ue = P.u + (N.u - P.u) * fxp; ve = P.v + (N.v - P.v) * fxp; we = P.w + (N.w - P.w) * fxp; }
          else                { ue = N.u + (P.u - N.u) * fxn; ve = N.v + (P.v - N.v) * fxn; we = N.w + (P.w - N.w) * fxn; }
          const double fuhigh = flomass * ue, fvhigh = flomass * ve, fwhigh = flomass * we;
          const double sup = -g.gds * (fuhigh - fuuds) + fdue - fdui;
          const double svp = -g.gds * (fvhigh - fvuds) + fdve - fdvi;
          const double swp = -g.gds * (fwhigh - fwuds) + fdwe - fdwi;
          if (own) { g.a[m.per_slot[b]] = can; s1 = s1 + sup; s2 = s2 + svp; s3 = s3 + swp; }   // a(icell,jcell) = can ; su(ijp) += sup
          else     { g.a[m.per_slot[b]] = cap; s1 = s1 - sup; s2 = s2 - svp; s3 = s3 - swp; }   // a(jcell,icell) = cap ; su(ijn) -= sup
        } else if (type == FCP_BC_SYMMETRY) {
          // :361-392 ; quirk Q7: vis(inp) with the stale inp = numCells+1 -> the viscosity of the first boundary slot
          const double are = sqrt(arx * arx + ary * ary + arz * arz), arer = 1.0 / are;
          const double nxf = arx * arer, nyf = ary * arer, nzf = arz * arer;
          const double dpb = (m.xf[f] - me.x) * nxf + (m.yf[f] - me.y) * nyf + (m.zf[f] - me.z) * nzf;
          const double cf = 2 * g.vis[m.n] * are / dpb;
          s1 = s1 - cf * nxf * (nyf * me.v + nzf * me.w);
          s2 = s2 - cf * nyf * (nxf * me.u + nzf * me.w);
          s3 = s3 - cf * nzf * (nxf * me.u + nyf * me.v);
          p1 = p1 + cf * (nxf * nxf); p2 = p2 + cf * (nyf * nyf); p3 = p3 + cf * (nzf * nzf);
        } else if (type == FCP_BC_WALL) {
          // :434-472
          const double viss = fmax(g.viscos, g.visw[o]);
          const double are = sqrt(arx * arx + ary * ary + arz * arz), arer = 1.0 / are;
          const double nxf = arx * arer, nyf = ary * arer, nzf = arz * arer;
          const double dpb = (m.xf[f] - me.x) * nxf + (m.yf[f] - me.y) * nyf + (m.zf[f] - me.z) * nzf;
          const double vsol = viss * are / dpb;
          const double ub = g.u[o], vb = g.v[o], wb = g.w[o];
          const double upb = me.u - ub, vpb = me.v - vb, wpb = me.w - wb;
          p1 = p1 + vsol * (1. - nxf * nxf); p2 = p2 + vsol * (1. - nyf * nyf); p3 = p3 + vsol * (1. - nzf * nzf);
          s1 = s1 + vsol * (ub * (1. - nxf * nxf) + vpb * nyf * nxf + wpb * nzf * nxf);
          s2 = s2 + vsol * (upb * nxf * nyf + vb * (1. - nyf * nyf) + wpb * nzf * nyf);
          s3 = s3 + vsol * (upb * nxf * nzf + vpb * nyf * nzf + wb * (1. - nzf * nzf));
        }
      }
      if (staged) fst ^= 1;      // the stage that received the next face's copies becomes the current one
    }
    if (flen == 0) have_pref = false;
    g.su[c] = s1; g.sv[c] = s2; g.sw[c] = s3;
    g.spu[c] = p1; g.spv[c] = p2; g.sp[c] = p3;
    if (g.rU) { g.rU[c] = s1; g.rV[c] = s2; g.rW[c] = s3; }   // piso: rU = su (:564-568)
  FCP_STAGED_LOOP_END
}

// diagonal + under-relaxation of one momentum equation, velocity.f90:602-620 (U), :651-668 (V), :712-730 (W).
// zero_first: the V and W passes first reset a(diag) and su (:651-654); the row sum then runs over the row with that zero.
__global__ void __launch_bounds__(FCP_TPB) k_uvw_diag(MeshView m, double *a, const double *__restrict__ spq, const double *srcq /* may alias su */,
                                                       const double *__restrict__ phi, double *__restrict__ apq, double *su, double urf, int zero_first) {
  const double urfr = 1.0 / urf, urfm = 1.0 - urf;
  FCP_CELL_LOOP(c, m.n) {
    const int64_t base = m.a_slptr[c >> 5] + (c & 31);
    const int32_t ri = m.a_rinfo[c];
    const int32_t dpos = (ri >> 16) & 0xffff;
    const int32_t len = ri & 0xffff;     // the whole row: on a partition the halo coefficients apr(ipro) belong to the diagonal too (src-par/calcuvw.f90:259-265 adds them through spu)
    const double adiag_old = zero_first ? 0.0 : a[base + (int64_t)dpos * 32];
    double s = 0.0;
    for (int32_t k = 0; k < len; ++k) s = s + (k == dpos ? adiag_old : a[base + (int64_t)k * 32]);   // sum( a(ia(inp):ia(inp+1)-1) ), CSR order
    const double sum_off = s - adiag_old;
    double ad = spq[c] - sum_off;
    apq[c] = 1. / (ad + FCP_SMALL);
    ad = ad * urfr;
    a[base + (int64_t)dpos * 32] = ad;
    su[c] = srcq[c] + urfm * ad * phi[c];
  }
}

int fvm_update_vel_bnd(fcp_ctx *ctx, double *u, double *v, double *w) {
  if (ctx->B == 0) return FCP_OK;
  k_update_vel_bnd<<<FCP_GRID(ctx->B)>>>(fcp_mesh_view(ctx), ctx->bftype, u, v, w);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_uvw_assemble(fcp_ctx *ctx, const UvwArgs &g) {
  if (ctx->n == 0) return FCP_OK;
  // FCP_UVW = fs (default: list stage + face stage, one CTA per SM) | list (list stage only, registers for one CTA per SM) | list2 (list stage only,
  // 128 registers = two CTAs per SM): A/B measurements
  const char *e = getenv("FCP_UVW");
  const int variant = (e && !strcmp(e, "list")) ? 1 : (e && !strcmp(e, "list2")) ? 2 : 0;
  const size_t smem = sizeof(ListStage<6>) + (variant == 0 ? sizeof(UvwFaceStage) : 0);
  const int grid = std::max(fcp_nchunks(ctx->n), 1);
  if (!(ctx->uvw_smem_configured & (1 << variant))) {
    cudaError_t ce = variant == 0 ? cudaFuncSetAttribute(k_uvw_assemble<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                   : variant == 1 ? cudaFuncSetAttribute(k_uvw_assemble<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                  : cudaFuncSetAttribute(k_uvw_assemble<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FCP_CUDA(ce);
    ctx->uvw_smem_configured |= 1 << variant;
  }
  size_t tok = ctx->prof.begin(FCP_K_UVW, ctx->stream);
  if (variant == 0) k_uvw_assemble<true, 1><<<grid, FCP_TPB, smem, ctx->stream>>>(fcp_mesh_view(ctx), g);
  else if (variant == 1) k_uvw_assemble<false, 1><<<grid, FCP_TPB, smem, ctx->stream>>>(fcp_mesh_view(ctx), g);
  else k_uvw_assemble<false, 2><<<grid, FCP_TPB, smem, ctx->stream>>>(fcp_mesh_view(ctx), g);
  ctx->prof.end(tok, ctx->stream);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_uvw_diag(fcp_ctx *ctx, double *a, const double *spq, const double *srcq, const double *phi, double *apq, double *su, double urf, int zero_first) {
  if (ctx->n == 0) return FCP_OK;
  k_uvw_diag<<<FCP_GRID(ctx->n)>>>(fcp_mesh_view(ctx), a, spq, srcq, phi, apq, su, urf, zero_first);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
