// fvm.cu -- finite-volume face-loop operators of the hot path as CELL-CENTRIC GATHERS.
//
// The reference runs face loops that scatter into both cells of a face (e.g. calcp_simple.f90:82-118).  Here one
// thread owns one cell and walks the cell's faces in ascending face index (inner faces, then boundary faces), which is
// exactly the order in which the reference's sequential face loops touch that cell.  Per-face quantities are evaluated
// in the face's own orientation (P = owner, N = neighbour) by the same instruction sequence on both sides, so both
// cells see bit-identical face values, there are no atomics, results are deterministic and round like the reference.
// The face lists are SELL-32 (fcp_internal.h): a warp reads 128 contiguous bytes per list step.
#include <algorithm>
#include <cstdlib>
#include "fcp_internal.h"
#include "reduce.cuh"
#include "fvm_common.cuh"

// owner-ordered copies of the face geometry (fcp_ctx::og); a kernel that uses a face index ONLY to address these seven arrays runs on them by a
// pointer swap in its MeshView (fcp_apply_face_variant): `ent` then holds +-(position + 1) instead of +-(face + 1)
__global__ void __launch_bounds__(FCP_TPB) k_build_og(int32_t nF, int64_t og_n, const int32_t *__restrict__ gpos, const double *__restrict__ arx,
                                                       const double *__restrict__ ary, const double *__restrict__ arz, const double *__restrict__ facint,
                                                       const double *__restrict__ xf, const double *__restrict__ yf, const double *__restrict__ zf,
                                                       double *__restrict__ og) {
  for (int32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nF; f += gridDim.x * blockDim.x) {
    const int64_t g = gpos[f];
    og[g] = arx[f]; og[og_n + g] = ary[f]; og[2 * og_n + g] = arz[f]; og[3 * og_n + g] = facint[f];
    og[4 * og_n + g] = xf[f]; og[5 * og_n + g] = yf[f]; og[6 * og_n + g] = zf[f];
  }
}
int fvm_ensure_og(fcp_ctx *ctx) {
  if (ctx->og_valid) return FCP_OK;
  const int32_t nF = ctx->F + ctx->B;
  if (!ctx->og) {
    FCP_TRY(dev_alloc(&ctx->og, 7 * (size_t)ctx->og_n));
    FCP_CUDA(cudaMemsetAsync(ctx->og, 0, sizeof(double) * 7 * (size_t)ctx->og_n, ctx->stream));
  }
  if (nF > 0) {
    k_build_og<<<std::min((nF + FCP_TPB - 1) / FCP_TPB, 148 * 16), FCP_TPB, 0, ctx->stream>>>(nF, ctx->og_n, ctx->d_gpos, ctx->arx, ctx->ary, ctx->arz, ctx->facint,
                                                                                              ctx->xf, ctx->yf, ctx->zf, ctx->og);
    FCP_LAUNCHED();
    FCP_CHECK_LAUNCH();
  }
  ctx->og_valid = true;
  return FCP_OK;
}

// ---------------------------------------------------------------------------------------------
// grad_gauss   gradients.f90:1607-1693
// ---------------------------------------------------------------------------------------------
// WS = faces per cell the shared-memory list stage holds (6: hexahedra and smaller; 10: the ten-faced polyhedra of config 5); longer lists are
// read from global memory as before
// MINB = CTAs per SM asked of the compiler; PF = 0: two list stages; PF >= 1: three list stages, and the operands of cell j + 1 are prefetched
// into the L2 (no registers) while cell j waits for its own gathers -- PF = 1: the cell's own streams and its owner-side faces (first touch; the
// other faces were read by their owner cells shortly before), PF = 2: every operand
template <int WS, int NS>
__device__ __forceinline__ void grad_gauss_prefetch(const MeshView &m, const double *u, const ListStage<WS, NS> &stage, int stn, int64_t cn, int pf) {
  if (cn >= m.n) return;
  fcp_prefetch_l2(u + cn); fcp_prefetch_l2(m.vol + cn);
  const bool cl = m.kinds != nullptr;
  const int32_t flen = stage.len(stn, cl);
#pragma unroll
  for (int k = 0; k < WS; ++k) {
    if (k >= flen) break;
    const int32_t e = stage.ent(stn, k);
    if (pf < 2 && e < 0) continue;
    const int32_t f = (e > 0 ? e : -e) - 1, o = stage.oth(stn, k);
    fcp_prefetch_l2(u + o);
    fcp_prefetch_l2(m.arx + f); fcp_prefetch_l2(m.ary + f); fcp_prefetch_l2(m.arz + f);
    if (stage.slot(stn, k, cl) >= 0) fcp_prefetch_l2(m.facint + f);
  }
}
template <int WS, int MINB, int PF>
__global__ void __launch_bounds__(FCP_TPB, MINB) k_grad_gauss(MeshView m, const double *__restrict__ u, double *__restrict__ g) {
  constexpr int NS = PF ? 3 : 2;
  FCP_STAGE_DYN_N(WS, NS, stage);
  FCP_STAGED_LOOP_BEGIN_N(NS, stage, m, m.n, c, st, stn, if (PF) grad_gauss_prefetch<WS, NS>(m, u, stage, stn, fcp_chunk_cell(j__ + 1), PF))
    double gx = 0.0, gy = 0.0, gz = 0.0;
    const double uc = u[c];
    const double vol = __ldg(m.vol + c);      // with the cell's first loads, not behind the face loop (a second DRAM round trip per cell)
    constexpr int W = WS <= 6 ? 6 : 5;
    FCP_FACE_BATCHES_STAGED(stage, st, m, c, W) {
      FCP_BATCH_LISTS_STAGED(stage, st, WS, m, W, e, o, sl);
      double sx[W], sy[W], sz[W], lam[W], uo[W];
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const int32_t f = (e[k] > 0 ? e[k] : -e[k]) - 1;
        const bool on = e[k] != 0;
        uo[k] = on ? __ldg(u + o[k]) : 0.0;
        lam[k] = (on && sl[k] >= 0) ? __ldg(m.facint + f) : 0.0;
        sx[k] = on ? __ldg(m.arx + f) : 0.0; sy[k] = on ? __ldg(m.ary + f) : 0.0; sz[k] = on ? __ldg(m.arz + f) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < W; ++k) {
        if (e[k] == 0) continue;
        if (sl[k] >= 0) {
          const double uP = e[k] > 0 ? uc : uo[k], uN = e[k] > 0 ? uo[k] : uc;
          const double fie = uP + (uN - uP) * lam[k];
          const double dfx = fie * sx[k], dfy = fie * sy[k], dfz = fie * sz[k];
          if (e[k] > 0) { gx = gx + dfx; gy = gy + dfy; gz = gz + dfz; }
          else          { gx = gx - dfx; gy = gy - dfy; gz = gz - dfz; }
        } else {
          gx = gx + uo[k] * sx[k]; gy = gy + uo[k] * sy[k]; gz = gz + uo[k] * sz[k];
        }
      }
    }
    const double volr = 1.0 / vol;
    g[3 * (int64_t)c + 0] = gx * volr;
    g[3 * (int64_t)c + 1] = gy * volr;
    g[3 * (int64_t)c + 2] = gz * volr;
  FCP_STAGED_LOOP_END
}

// ---------------------------------------------------------------------------------------------
// least squares: create_matrix_lsq :660-779 / create_matrix_lsq_dm :1157-1326 ; grad_lsq :782-893 / grad_lsq_dm :1334-1486
// Dmat is stored SoA [9][n] on the device.
// ---------------------------------------------------------------------------------------------
template <bool W>
__global__ void __launch_bounds__(FCP_TPB) k_lsq_matrix(MeshView m, double *__restrict__ D) {
  FCP_CELL_LOOP(c, m.n) {
    double d11 = 0.0, d12 = 0.0, d13 = 0.0, d22 = 0.0, d23 = 0.0, d33 = 0.0;
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c];
    FCP_FACE_LOOP(m, c) {
      FCP_FACE_FETCH(m);
      double Dx, Dy, Dz;
      if (sl >= 0) {
        const double xo = m.xc[o], yo = m.yc[o], zo = m.zc[o];
        if (e > 0) { Dx = xo - xc; Dy = yo - yc; Dz = zo - zc; }
        else       { Dx = xc - xo; Dy = yc - yo; Dz = zc - zo; }
      } else {
        Dx = m.xf[f] - xc; Dy = m.yf[f] - yc; Dz = m.zf[f] - zc;
      }
      if (W) {
        const double w = 1.0 / (Dx * Dx + Dy * Dy + Dz * Dz);
        d11 = d11 + w * Dx * Dx; d22 = d22 + w * Dy * Dy; d33 = d33 + w * Dz * Dz;
        d12 = d12 + w * Dx * Dy; d13 = d13 + w * Dx * Dz; d23 = d23 + w * Dy * Dz;
      } else {
        d11 = d11 + Dx * Dx; d22 = d22 + Dy * Dy; d33 = d33 + Dz * Dz;
        d12 = d12 + Dx * Dy; d13 = d13 + Dx * Dz; d23 = d23 + Dy * Dz;
      }
    }
    const double d21 = d12, d31 = d13, d32 = d23;   // :748-777
    const double tmp = 1.0 / (d11 * d22 * d33 - d11 * d23 * d32 - d12 * d21 * d33 + d12 * d23 * d31 + d13 * d21 * d32 - d13 * d22 * d31 + FCP_SMALL);
    const int64_t n = m.n;
    D[0 * n + c] = (d22 * d33 - d23 * d32) * tmp;
    D[1 * n + c] = (d21 * d33 - d23 * d31) * tmp;
    D[2 * n + c] = (d21 * d32 - d22 * d31) * tmp;
    D[3 * n + c] = (d11 * d33 - d13 * d31) * tmp;
    D[4 * n + c] = (d12 * d33 - d13 * d32) * tmp;
    D[5 * n + c] = (d11 * d32 - d12 * d31) * tmp;
    D[6 * n + c] = (d12 * d23 - d13 * d22) * tmp;
    D[7 * n + c] = (d11 * d23 - d13 * d21) * tmp;
    D[8 * n + c] = (d11 * d22 - d12 * d21) * tmp;
  }
}

template <int WS, int NS>
__device__ __forceinline__ void grad_lsq_prefetch(const MeshView &m, const double *D, const double *phi, const ListStage<WS, NS> &stage, int stn, int64_t cn,
                                                  int pf) {
  if (cn >= m.n) return;
  fcp_prefetch_l2(phi + cn); fcp_prefetch_l2(m.xc + cn); fcp_prefetch_l2(m.yc + cn); fcp_prefetch_l2(m.zc + cn);
#pragma unroll
  for (int q = 0; q < 9; ++q) fcp_prefetch_l2(D + (int64_t)q * m.n + cn);
  const bool cl = m.kinds != nullptr;
  const int32_t flen = stage.len(stn, cl);
#pragma unroll
  for (int k = 0; k < WS; ++k) {
    if (k >= flen) break;
    const int32_t e = stage.ent(stn, k);
    if (pf < 2 && e < 0) continue;
    const int32_t f = (e > 0 ? e : -e) - 1, o = stage.oth(stn, k);
    fcp_prefetch_l2(phi + o);
    if (stage.slot(stn, k, cl) >= 0) { fcp_prefetch_l2(m.xc + o); fcp_prefetch_l2(m.yc + o); fcp_prefetch_l2(m.zc + o); }
    else { fcp_prefetch_l2(m.xf + f); fcp_prefetch_l2(m.yf + f); fcp_prefetch_l2(m.zf + f); }
  }
}
template <bool W, int WS, int MINB, int PF>
__global__ void __launch_bounds__(FCP_TPB, MINB) k_grad_lsq(MeshView m, const double *__restrict__ D, const double *__restrict__ phi,
                                                             double *__restrict__ g, int row2_reference) {
  constexpr int NS = PF ? 3 : 2;
  FCP_STAGE_DYN_N(WS, NS, stage);
  FCP_STAGED_LOOP_BEGIN_N(NS, stage, m, m.n, c, st, stn, if (PF) grad_lsq_prefetch<WS, NS>(m, D, phi, stage, stn, fcp_chunk_cell(j__ + 1), PF))
    double b1 = 0.0, b2 = 0.0, b3 = 0.0;
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c], pc = phi[c];
    const int64_t n = m.n;
    // the cell's row of the inverse matrix with the cell's first loads, not behind the face loop (a second DRAM round trip per cell)
    const double D1 = __ldg(D + 0 * n + c), D2 = __ldg(D + 1 * n + c), D3 = __ldg(D + 2 * n + c), D4 = __ldg(D + 3 * n + c), D5 = __ldg(D + 4 * n + c),
                 D6 = __ldg(D + 5 * n + c), D7 = __ldg(D + 6 * n + c), D8 = __ldg(D + 7 * n + c), D9 = __ldg(D + 8 * n + c);
    constexpr int WB = WS <= 6 ? 6 : 5;
    FCP_FACE_BATCHES_STAGED(stage, st, m, c, WB) {
      FCP_BATCH_LISTS_STAGED(stage, st, WS, m, WB, e_, o_, sl_);
      // every gather of the batch before the first use: cell centres and phi across two-sided faces, face centres and boundary values otherwise
      double x_[WB], y_[WB], z_[WB], p_[WB];
#pragma unroll
      for (int k = 0; k < WB; ++k) {
        const bool on = e_[k] != 0, two = on && sl_[k] >= 0;
        const int32_t f = (e_[k] > 0 ? e_[k] : -e_[k]) - 1;
        x_[k] = !on ? 0.0 : two ? __ldg(m.xc + o_[k]) : __ldg(m.xf + f);
        y_[k] = !on ? 0.0 : two ? __ldg(m.yc + o_[k]) : __ldg(m.yf + f);
        z_[k] = !on ? 0.0 : two ? __ldg(m.zc + o_[k]) : __ldg(m.zf + f);
        p_[k] = on ? __ldg(phi + o_[k]) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < WB; ++k) {
        if (e_[k] == 0) continue;
        const int32_t e = e_[k], sl = sl_[k];
        const int32_t f = (e > 0 ? e : -e) - 1;
      double Dx, Dy, Dz;
      if (sl >= 0) {
        const double xo = x_[k], yo = y_[k], zo = z_[k], po = p_[k];
        double dx, dy, dz, dphi;
        if (e > 0) { dx = xo - xc; dy = yo - yc; dz = zo - zc; dphi = po - pc; }
        else       { dx = xc - xo; dy = yc - yo; dz = zc - zo; dphi = pc - po; }
        if (W) {
          const double w = dphi / (dx * dx + dy * dy + dz * dz);
          Dx = w * dx; Dy = w * dy; Dz = w * dz;
        } else {
          Dx = dx * dphi; Dy = dy * dphi; Dz = dz * dphi;
        }
      } else {
        const double dx = x_[k] - xc, dy = y_[k] - yc, dz = z_[k] - zc;
        const double dphi = p_[k] - pc;
        if (W) {
          // quirk Q2 (gradients.f90:1459): the weight's denominator indexes xf with the BOUNDARY COUNTER i, not iface
          const int32_t i = f - m.F;
          const double ex = m.xf[i] - xc, ey = m.yf[i] - yc, ez = m.zf[i] - zc;
          const double w = dphi / (ex * ex + ey * ey + ez * ez);
          Dx = w * dx; Dy = w * dy; Dz = w * dz;
        } else {
          Dx = dx * dphi; Dy = dy * dphi; Dz = dz * dphi;
        }
      }
      b1 = b1 + Dx; b2 = b2 + Dy; b3 = b3 + Dz;
      }
    }
    g[3 * (int64_t)c + 0] = b1 * D1 - b2 * D2 + b3 * D3;                   // :886-888 ; row 2 is quirk Q1
    g[3 * (int64_t)c + 1] = row2_reference ? (b1 * D4 - b2 * D5 - b3 * D6) : (b2 * D4 - b1 * D5 - b3 * D6);
    g[3 * (int64_t)c + 2] = b1 * D7 - b2 * D8 + b3 * D9;
  FCP_STAGED_LOOP_END
}

// ---------------------------------------------------------------------------------------------
// laplacian(mu,phi)   fvImplicit/laplacian.f90  (src-par/fvm_laplacian.f90: process faces -> halo column)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FCP_TPB) k_laplacian(MeshView m, const double *__restrict__ mu, const double *__restrict__ phi,
                                                        double *__restrict__ a, double *__restrict__ su) {
  FCP_CELL_LOOP(c, m.n) {
    double dg = 0.0, s = su[c];
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c], muc = mu[c];
    FCP_FACE_LOOP(m, c) {
      FCP_FACE_FETCH(m);
      const double sx = m.arx[f], sy = m.ary[f], sz = m.arz[f];
      if (sl >= 0) {
        const double xo = m.xc[o], yo = m.yc[o], zo = m.zc[o], muo = mu[o];
        double xpn, ypn, zpn, muP, muN;
        if (e > 0) { xpn = xo - xc; ypn = yo - yc; zpn = zo - zc; muP = muc; muN = muo; }
        else       { xpn = xc - xo; ypn = yc - yo; zpn = zc - zo; muP = muo; muN = muc; }
        const double fxn = m.facint[f], fxp = 1.0 - fxn;
        const double smdpn = (sx * sx + sy * sy + sz * sz) / (sx * xpn + sy * ypn + sz * zpn);
        const double cap = (fxp * muP + fxn * muN) * smdpn;
        a[sl] = cap;
        dg = dg - cap;
      } else {
        const double are = sqrt(sx * sx + sy * sy + sz * sz);
        const double nxf = sx / are, nyf = sy / are, nzf = sz / are;
        const double dfn = (m.xf[f] - xc) * nxf + (m.yf[f] - yc) * nyf + (m.zf[f] - zc) * nzf;
        const double dcoef = muc * are / dfn;
        dg = dg - dcoef;
        s = s - dcoef * phi[o];
      }
    }
    a[diag_pos(m, c)] = dg;
    su[c] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// gradp_and_sources(p)   Pressure/nablap.f90:19-208 + Pressure/bpres.f90
// Everything a cell needs (its inner-face sum, its own boundary faces, both extrapolation stages) is local to the
// cell, so 'linear' and 'weighted' run as ONE kernel.  'central' needs the stage-1 gradient of the neighbours and is
// split in two kernels (k_gradp<...,1> then k_gradp_central2).
// CORRECT: fuse the velocity / pressure correction and updateVelocityAtBoundary of calcp_simple.f90:416-429.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double face_p(int scheme, double pP, double pN, double lam, double aP, double aN) {
  if (scheme == FCP_PSCHEME_WEIGHTED) return (pP * aP + pN * aN) / (aP + aN + FCP_SMALL);   // nablap.f90:87
  return pP + (pN - pP) * lam;                                                              // face_value_cds, interpolation.f90:155
}

// Generic cell body (any number of faces): the face list is walked again for every boundary stage.
template <bool CORRECT>
__device__ __forceinline__ void gradp_cell_generic(const MeshView &m, int32_t c, int scheme, int nstages, double *p, const double *__restrict__ apu,
                                                   double *__restrict__ su, double *__restrict__ sv, double *__restrict__ sw,
                                                   double *__restrict__ dPdxi, const CorrectArgs &ca) {
    const double pc = p[c];
    const double ac = scheme == FCP_PSCHEME_WEIGHTED ? apu[c] : 0.0;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    bool has_bnd = false;
    {
      constexpr int W = 6;
      FCP_FACE_BATCHES(m, c, W) {
        FCP_BATCH_LISTS(m, W, e, o, sl);
        double sx[W], sy[W], sz[W], lam[W], po[W], ao[W];
#pragma unroll
        for (int k = 0; k < W; ++k) {
          const int32_t f = (e[k] > 0 ? e[k] : -e[k]) - 1;
          const bool two = e[k] != 0 && sl[k] >= 0;
          // inner-face values of p are never written by this kernel (only boundary slots are): non-coherent loads are safe
          po[k] = two ? __ldg(p + o[k]) : 0.0;
          ao[k] = (two && scheme == FCP_PSCHEME_WEIGHTED) ? __ldg(apu + o[k]) : 0.0;
          lam[k] = two ? __ldg(m.facint + f) : 0.0;
          sx[k] = two ? __ldg(m.arx + f) : 0.0; sy[k] = two ? __ldg(m.ary + f) : 0.0; sz[k] = two ? __ldg(m.arz + f) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < W; ++k) {
          if (e[k] == 0) continue;
          if (sl[k] >= 0) {
            const double pf = e[k] > 0 ? face_p(scheme, pc, po[k], lam[k], ac, ao[k]) : face_p(scheme, po[k], pc, lam[k], ao[k], ac);
            const double dfx = pf * sx[k], dfy = pf * sy[k], dfz = pf * sz[k];
            if (e[k] > 0) { s1 = s1 - dfx; s2 = s2 - dfy; s3 = s3 - dfz; }
            else          { s1 = s1 + dfx; s2 = s2 + dfy; s3 = s3 + dfz; }
          } else {
            has_bnd = true;
          }
        }
      }
    }
    const double volr = 1.0 / m.vol[c];
    // stage 1: bpres(p,1): p_b = p_P on every patch that is not a pressure patch
    double gx = -s1, gy = -s2, gz = -s3;
    if (has_bnd) {
      FCP_FACE_LOOP(m, c) {
        FCP_FACE_FETCH(m);
        if (sl < 0) {
          const int type = -1 - sl;
          double pb;
          if (type == FCP_BC_PRESSURE) pb = p[o]; else { pb = pc; p[o] = pb; }
          gx = gx + pb * m.arx[f]; gy = gy + pb * m.ary[f]; gz = gz + pb * m.arz[f];
        }
      }
    }
    gx = gx * volr; gy = gy * volr; gz = gz * volr;
    if (nstages >= 2) {
      // stage 2: bpres(p,2): walls are linearly extrapolated with the stage-1 gradient
      if (has_bnd) {
        const double g1x = gx, g1y = gy, g1z = gz;
        gx = -s1; gy = -s2; gz = -s3;
        const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c];
        FCP_FACE_LOOP(m, c) {
          FCP_FACE_FETCH(m);
          if (sl < 0) {
            const int type = -1 - sl;
            double pb;
            if (type == FCP_BC_WALL) {
              const double xpb = m.xf[f] - xc, ypb = m.yf[f] - yc, zpb = m.zf[f] - zc;
              pb = pc + g1x * xpb + g1y * ypb + g1z * zpb;
              p[o] = pb;
            } else {
              pb = p[o];
            }
            gx = gx + pb * m.arx[f]; gy = gy + pb * m.ary[f]; gz = gz + pb * m.arz[f];
          }
        }
        gx = gx * volr; gy = gy * volr; gz = gz * volr;
      }
      // (cells without boundary faces: stage 2 recomputes the identical value)
    }
    dPdxi[3 * (int64_t)c + 0] = gx;
    dPdxi[3 * (int64_t)c + 1] = gy;
    dPdxi[3 * (int64_t)c + 2] = gz;
    if (nstages >= 2) {
      // boundary-face part of the momentum sources, nablap.f90:195-204
      if (has_bnd) {
        FCP_FACE_LOOP(m, c) {
          FCP_FACE_FETCH(m);
          if (sl < 0) {
            const double pb = p[o];
            s1 = s1 - pb * m.arx[f]; s2 = s2 - pb * m.ary[f]; s3 = s3 - pb * m.arz[f];
          }
        }
      }
      if (su) { su[c] = s1; sv[c] = s2; sw[c] = s3; }   // su == nullptr: gradient + boundary extrapolation only (calcp_piso.f90:336-344)
      if (CORRECT) {
        // calcp_simple.f90:416-419
        const double ppref = ca.ppref_src ? *ca.ppref_src : 0.0;
        const double un = ca.u[c] + s1 * ca.apu[c];
        const double vn = ca.v[c] + s2 * ca.apv[c];
        const double wn = ca.w[c] + s3 * ca.apw[c];
        ca.u[c] = un; ca.v[c] = vn; ca.w[c] = wn;
        if (ca.pres) ca.pres[c] = ca.pres[c] + ca.urfp * (pc - ppref);   // nullptr: calcp_piso.f90:429-431 corrects velocities only
        if (has_bnd) {   // updateVelocityAtBoundary, velocity.f90:1184-1277
          FCP_FACE_LOOP(m, c) {
            FCP_FACE_FETCH(m);
            if (sl < 0) {
              const int type = -1 - sl;
              if (type == FCP_BC_EMPTY || type == FCP_BC_PERIODIC) {
                ca.u[o] = un; ca.v[o] = vn; ca.w[o] = wn;
              } else if (type == FCP_BC_SYMMETRY) {
                const double sx = m.arx[f], sy = m.ary[f], sz = m.arz[f];
                const double Unmag = un * sx + vn * sy + wn * sz;
                ca.u[o] = un - Unmag * sx; ca.v[o] = vn - Unmag * sy; ca.w[o] = wn - Unmag * sz;
              }
            }
          }
        }
      }
    } else {
      su[c] = s1; sv[c] = s2; sw[c] = s3;   // stage-1 inner sums kept for the central second pass
    }
}

// Fast cell body for cells with at most W faces (every hexahedron, prism, tetrahedron): the whole face list and every
// face quantity are loaded ONCE into registers (two memory round trips per cell) and all the stages -- inner-face sum,
// bpres stage 1, bpres stage 2, boundary part of the sources, velocity correction, updateVelocityAtBoundary -- run out
// of registers.  Boundary cells then cost one extra round trip (the wall-face centres) instead of four dependent list
// walks; on a structured mesh a quarter of all warps contain a boundary cell, so this is what bounds the kernel.
// Arithmetic order per cell is unchanged (faces in ascending index).
template <bool CORRECT, bool WEIGHTED, int W>
__device__ __forceinline__ void gradp_cell_fast(const MeshView &m, int32_t c, const int32_t (&e)[W], const int32_t (&o)[W], const int32_t (&sl)[W], int scheme,
                                                int nstages, double *p, const double *__restrict__ apu, double *__restrict__ su, double *__restrict__ sv,
                                                double *__restrict__ sw, double *__restrict__ dPdxi, const CorrectArgs &ca) {
  const double pc = p[c];
  const double ac = WEIGHTED ? apu[c] : 0.0;
  const double vol = m.vol[c];
  double sx[W], sy[W], sz[W], lam[W], pv[W], ao[WEIGHTED ? W : 1];
  bool has_bnd = false, has_wall = false;
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const int32_t f = (e[k] > 0 ? e[k] : -e[k]) - 1;
    const bool on = e[k] != 0, two = on && sl[k] >= 0, bnd = on && sl[k] < 0;
    has_bnd = has_bnd || bnd;
    has_wall = has_wall || (bnd && (-1 - sl[k]) == FCP_BC_WALL);
    // p: inner-face (cell / ghost) values are never written by this kernel -> non-coherent loads; pressure-patch values likewise
    pv[k] = (two || (bnd && (-1 - sl[k]) == FCP_BC_PRESSURE)) ? __ldg(p + o[k]) : 0.0;
    if (WEIGHTED) ao[k] = two ? __ldg(apu + o[k]) : 0.0;
    lam[k] = two ? __ldg(m.facint + f) : 0.0;
    sx[k] = on ? __ldg(m.arx + f) : 0.0; sy[k] = on ? __ldg(m.ary + f) : 0.0; sz[k] = on ? __ldg(m.arz + f) : 0.0;
  }
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
  for (int k = 0; k < W; ++k) {
    if (e[k] != 0 && sl[k] >= 0) {
      const double aok = WEIGHTED ? ao[k] : 0.0;
      const int sch = WEIGHTED ? FCP_PSCHEME_WEIGHTED : FCP_PSCHEME_LINEAR;
      const double pf = e[k] > 0 ? face_p(sch, pc, pv[k], lam[k], ac, aok) : face_p(sch, pv[k], pc, lam[k], aok, ac);
      const double dfx = pf * sx[k], dfy = pf * sy[k], dfz = pf * sz[k];
      if (e[k] > 0) { s1 = s1 - dfx; s2 = s2 - dfy; s3 = s3 - dfz; }
      else          { s1 = s1 + dfx; s2 = s2 + dfy; s3 = s3 + dfz; }
    }
  }
  const double volr = 1.0 / vol;
  // stage 1: bpres(p,1): p_b = p_P on every patch that is not a pressure patch (pv[k] becomes the boundary value)
  double gx = -s1, gy = -s2, gz = -s3;
  if (has_bnd) {
#pragma unroll
    for (int k = 0; k < W; ++k) {
      if (e[k] != 0 && sl[k] < 0) {
        const int type = -1 - sl[k];
        if (type != FCP_BC_PRESSURE) {
          pv[k] = pc;
          if (!(nstages >= 2 && type == FCP_BC_WALL)) p[o[k]] = pc;   // (wall values are overwritten by stage 2 anyway)
        }
        gx = gx + pv[k] * sx[k]; gy = gy + pv[k] * sy[k]; gz = gz + pv[k] * sz[k];
      }
    }
  }
  gx = gx * volr; gy = gy * volr; gz = gz * volr;
  if (nstages >= 2) {
    if (has_bnd) {
      // stage 2: bpres(p,2): walls are linearly extrapolated with the stage-1 gradient
      const double g1x = gx, g1y = gy, g1z = gz;
      gx = -s1; gy = -s2; gz = -s3;
      if (has_wall) {
        const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c];
        double xw[W], yw[W], zw[W];
#pragma unroll
        for (int k = 0; k < W; ++k) {
          const bool wall = e[k] != 0 && sl[k] == -1 - FCP_BC_WALL;
          const int32_t f = (e[k] > 0 ? e[k] : -e[k]) - 1;
          xw[k] = wall ? __ldg(m.xf + f) : 0.0; yw[k] = wall ? __ldg(m.yf + f) : 0.0; zw[k] = wall ? __ldg(m.zf + f) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < W; ++k) {
          if (e[k] != 0 && sl[k] == -1 - FCP_BC_WALL) {
            const double xpb = xw[k] - xc, ypb = yw[k] - yc, zpb = zw[k] - zc;
            pv[k] = pc + g1x * xpb + g1y * ypb + g1z * zpb;
            p[o[k]] = pv[k];
          }
        }
      }
#pragma unroll
      for (int k = 0; k < W; ++k) {
        if (e[k] != 0 && sl[k] < 0) { gx = gx + pv[k] * sx[k]; gy = gy + pv[k] * sy[k]; gz = gz + pv[k] * sz[k]; }
      }
      gx = gx * volr; gy = gy * volr; gz = gz * volr;
    }
  }
  dPdxi[3 * (int64_t)c + 0] = gx;
  dPdxi[3 * (int64_t)c + 1] = gy;
  dPdxi[3 * (int64_t)c + 2] = gz;
  if (nstages >= 2) {
    if (has_bnd) {   // boundary-face part of the momentum sources, nablap.f90:195-204
#pragma unroll
      for (int k = 0; k < W; ++k) {
        if (e[k] != 0 && sl[k] < 0) { s1 = s1 - pv[k] * sx[k]; s2 = s2 - pv[k] * sy[k]; s3 = s3 - pv[k] * sz[k]; }
      }
    }
    if (su) { su[c] = s1; sv[c] = s2; sw[c] = s3; }
    if (CORRECT) {
      // calcp_simple.f90:416-419
      const double ppref = ca.ppref_src ? *ca.ppref_src : 0.0;
      const double un = ca.u[c] + s1 * ca.apu[c];
      const double vn = ca.v[c] + s2 * ca.apv[c];
      const double wn = ca.w[c] + s3 * ca.apw[c];
      ca.u[c] = un; ca.v[c] = vn; ca.w[c] = wn;
      if (ca.pres) ca.pres[c] = ca.pres[c] + ca.urfp * (pc - ppref);
      if (has_bnd) {   // updateVelocityAtBoundary, velocity.f90:1184-1277
#pragma unroll
        for (int k = 0; k < W; ++k) {
          if (e[k] != 0 && sl[k] < 0) {
            const int type = -1 - sl[k];
            if (type == FCP_BC_EMPTY || type == FCP_BC_PERIODIC) {
              ca.u[o[k]] = un; ca.v[o[k]] = vn; ca.w[o[k]] = wn;
            } else if (type == FCP_BC_SYMMETRY) {
              const double Unmag = un * sx[k] + vn * sy[k] + wn * sz[k];
              ca.u[o[k]] = un - Unmag * sx[k]; ca.v[o[k]] = vn - Unmag * sy[k]; ca.w[o[k]] = wn - Unmag * sz[k];
            }
          }
        }
      }
    }
  } else {
    su[c] = s1; sv[c] = s2; sw[c] = s3;   // stage-1 inner sums kept for the central second pass
  }
}

// the streams of the fused velocity / pressure correction are only read behind the face loop: asked into the L2 with the cell's first loads, so that
// the tail costs an L2 hit instead of a second DRAM round trip per cell
template <bool CORRECT>
__device__ __forceinline__ void gradp_prefetch_tail(const CorrectArgs &ca, int64_t c) {
  if (!CORRECT) return;
  fcp_prefetch_l2(ca.u + c); fcp_prefetch_l2(ca.v + c); fcp_prefetch_l2(ca.w + c);
  fcp_prefetch_l2(ca.apu + c); fcp_prefetch_l2(ca.apv + c); fcp_prefetch_l2(ca.apw + c);
  if (ca.pres) fcp_prefetch_l2(ca.pres + c);
}
template <bool CORRECT, bool WEIGHTED, int WS, int NS>
__device__ __forceinline__ void gradp_prefetch(const MeshView &m, const double *p, const double *apu, const CorrectArgs &ca, const ListStage<WS, NS> &stage,
                                               int stn, int64_t cn, int pf) {
  if (cn >= m.n) return;
  fcp_prefetch_l2(p + cn); fcp_prefetch_l2(m.vol + cn);
  if (WEIGHTED) fcp_prefetch_l2(apu + cn);
  gradp_prefetch_tail<CORRECT>(ca, cn);
  const bool cl = m.kinds != nullptr;
  const int32_t flen = stage.len(stn, cl);
#pragma unroll
  for (int k = 0; k < WS; ++k) {
    if (k >= flen) break;
    const int32_t e = stage.ent(stn, k);
    if (pf < 2 && e < 0) continue;
    const int32_t f = (e > 0 ? e : -e) - 1, o = stage.oth(stn, k), sl = stage.slot(stn, k, cl);
    fcp_prefetch_l2(m.arx + f); fcp_prefetch_l2(m.ary + f); fcp_prefetch_l2(m.arz + f);
    if (sl >= 0) {
      fcp_prefetch_l2(p + o); fcp_prefetch_l2(m.facint + f);
      if (WEIGHTED) fcp_prefetch_l2(apu + o);
    } else if (-1 - sl == FCP_BC_PRESSURE) {
      fcp_prefetch_l2(p + o);
    }
  }
}
template <bool CORRECT, bool WEIGHTED, int PF>
__global__ void __launch_bounds__(FCP_TPB, 2) k_gradp(MeshView m, int nstages, double *p, const double *__restrict__ apu,
                                                       double *__restrict__ su, double *__restrict__ sv, double *__restrict__ sw,
                                                       double *__restrict__ dPdxi, CorrectArgs ca) {
  constexpr int W = 6;
  constexpr int NS = PF ? 3 : 2;
  constexpr int scheme = WEIGHTED ? FCP_PSCHEME_WEIGHTED : FCP_PSCHEME_LINEAR;
  FCP_STAGE_DYN_N(W, NS, stage);      // the face list of the NEXT cell(s) travels global -> shared while this cell's gathers are in flight
  FCP_STAGED_LOOP_BEGIN_N(NS, stage, m, m.n, c, st, stn,
                          if (PF) gradp_prefetch<CORRECT, WEIGHTED, W, NS>(m, p, apu, ca, stage, stn, fcp_chunk_cell(j__ + 1), PF))
    if (!PF && nstages >= 2) gradp_prefetch_tail<CORRECT>(ca, c);
    const bool cl = m.kinds != nullptr;
    if (stage.len(st, cl) <= W) {
      int32_t e[W], o[W], sl[W];
      stage.read(st, e, o, sl, cl);
      gradp_cell_fast<CORRECT, WEIGHTED, W>(m, c, e, o, sl, scheme, nstages, p, apu, su, sv, sw, dPdxi, ca);
    } else {
      gradp_cell_generic<CORRECT>(m, c, scheme, nstages, p, apu, su, sv, sw, dPdxi, ca);
    }
  FCP_STAGED_LOOP_END
}

// 'central' stage 2 (nablap.f90:129-160): inner-face sum recomputed with face_value_central (interpolation.f90:218-264)
// using the stage-1 gradients g1 of both cells; writes the final gradient to gout (!= g1).
template <bool CORRECT>
__global__ void __launch_bounds__(FCP_TPB) k_gradp_central2(MeshView m, double *p, const double *__restrict__ g1, double *__restrict__ su,
                                                             double *__restrict__ sv, double *__restrict__ sw, double *__restrict__ gout,
                                                             CorrectArgs ca) {
  FCP_CELL_LOOP(c, m.n) {
    const double pc = p[c];
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c];
    const double gcx = g1[3 * (int64_t)c], gcy = g1[3 * (int64_t)c + 1], gcz = g1[3 * (int64_t)c + 2];
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    bool has_bnd = false;
    {
      FCP_FACE_LOOP(m, c) {
        FCP_FACE_FETCH(m);
        if (sl >= 0) {
          const double po = p[o];
          const double xo = m.xc[o], yo = m.yc[o], zo = m.zc[o];
          const double gox = g1[3 * (int64_t)o], goy = g1[3 * (int64_t)o + 1], goz = g1[3 * (int64_t)o + 2];
          const double xfa = m.xf[f], yfa = m.yf[f], zfa = m.zf[f];
          double gradfidr, pP, pN;
          if (e > 0) {
            gradfidr = gcx * (xfa - xc) + gcy * (yfa - yc) + gcz * (zfa - zc) + gox * (xfa - xo) + goy * (yfa - yo) + goz * (zfa - zo);
            pP = pc; pN = po;
          } else {
            gradfidr = gox * (xfa - xo) + goy * (yfa - yo) + goz * (zfa - zo) + gcx * (xfa - xc) + gcy * (yfa - yc) + gcz * (zfa - zc);
            pP = po; pN = pc;
          }
          const double pf = 0.5 * (pP + pN + gradfidr);
          const double dfx = pf * m.arx[f], dfy = pf * m.ary[f], dfz = pf * m.arz[f];
          if (e > 0) { s1 = s1 - dfx; s2 = s2 - dfy; s3 = s3 - dfz; }
          else       { s1 = s1 + dfx; s2 = s2 + dfy; s3 = s3 + dfz; }
        } else {
          has_bnd = true;
        }
      }
    }
    const double volr = 1.0 / m.vol[c];
    double gx = -s1, gy = -s2, gz = -s3;
    if (has_bnd) {
      FCP_FACE_LOOP(m, c) {
        FCP_FACE_FETCH(m);
        if (sl < 0) {
          const int type = -1 - sl;
          double pb;
          if (type == FCP_BC_WALL) {
            const double xpb = m.xf[f] - xc, ypb = m.yf[f] - yc, zpb = m.zf[f] - zc;
            pb = pc + gcx * xpb + gcy * ypb + gcz * zpb;
            p[o] = pb;
          } else {
            pb = p[o];
          }
          gx = gx + pb * m.arx[f]; gy = gy + pb * m.ary[f]; gz = gz + pb * m.arz[f];
          s1 = s1 - pb * m.arx[f]; s2 = s2 - pb * m.ary[f]; s3 = s3 - pb * m.arz[f];
        }
      }
    }
    gout[3 * (int64_t)c + 0] = gx * volr;
    gout[3 * (int64_t)c + 1] = gy * volr;
    gout[3 * (int64_t)c + 2] = gz * volr;
    su[c] = s1; sv[c] = s2; sw[c] = s3;
    if (CORRECT) {
      const double ppref = ca.ppref_src ? *ca.ppref_src : 0.0;
      const double un = ca.u[c] + s1 * ca.apu[c];
      const double vn = ca.v[c] + s2 * ca.apv[c];
      const double wn = ca.w[c] + s3 * ca.apw[c];
      ca.u[c] = un; ca.v[c] = vn; ca.w[c] = wn;
      if (ca.pres) ca.pres[c] = ca.pres[c] + ca.urfp * (pc - ppref);   // nullptr: calcp_piso.f90:429-431 corrects velocities only
      if (has_bnd) {
        FCP_FACE_LOOP(m, c) {
          FCP_FACE_FETCH(m);
          if (sl < 0) {
            const int type = -1 - sl;
            if (type == FCP_BC_EMPTY || type == FCP_BC_PERIODIC) {
              ca.u[o] = un; ca.v[o] = vn; ca.w[o] = wn;
            } else if (type == FCP_BC_SYMMETRY) {
              const double sx = m.arx[f], sy = m.ary[f], sz = m.arz[f];
              const double Unmag = un * sx + vn * sy + wn * sz;
              ca.u[o] = un - Unmag * sx; ca.v[o] = vn - Unmag * sy; ca.w[o] = wn - Unmag * sz;
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// pressure-correction assembly   calcp_simple.f90:69-234 + facefluxmass2 faceflux_mass.f90:175-249
// ---------------------------------------------------------------------------------------------
// PISO = true: facefluxmass_piso (faceflux_mass.f90:389-459): the flux is the plain interpolated HbyA flux (no Rhie-Chow
// pressure term) and pressure patches do not reset pp (calcp_piso.f90:140-240).
// MPIF = true: inner faces take the MPI tree's `facefluxmass` (quirk Q10, src-par/faceflux_mass.f90:28-180; process faces keep facefluxmass2 like
// src-par/calcp_simple.f90:96-118): gradient-corrected central velocities, per-component (Vol/Ap)_f, the P'/E' pressure correction with its sign
// quirk Q26.  A switchable variant, not the benchmarked path: its extra operands are plain dependent loads.
// PF >= 1: three list stages, and the operands of cell j + 1 are asked into the L2 while cell j is evaluated (see k_grad_gauss), so that the W-face
// gather rounds of a cell cost L2 hits instead of DRAM round trips
template <int WS, int NS>
__device__ __forceinline__ void assemble_pcorr_prefetch(const MeshView &m, const AsmArgs &g, const ListStage<WS, NS> &stage, int stn, int64_t cn, int pf) {
  if (cn >= m.n) return;
  fcp_prefetch_l2(m.xc + cn); fcp_prefetch_l2(m.yc + cn); fcp_prefetch_l2(m.zc + cn); fcp_prefetch_l2(g.den + cn); fcp_prefetch_l2(m.vol + cn);
  fcp_prefetch_l2(g.apu + cn); fcp_prefetch_l2(g.u + cn); fcp_prefetch_l2(g.v + cn); fcp_prefetch_l2(g.w + cn); fcp_prefetch_l2(g.p + cn);
  fcp_prefetch_l2(g.dPdxi + 3 * cn); fcp_prefetch_l2(g.dPdxi + 3 * cn + 2); fcp_prefetch_l2(m.a_rinfo + cn);
  const int32_t flen = stage.len(stn);
#pragma unroll
  for (int k = 0; k < WS; ++k) {
    if (k >= flen) break;
    const int32_t e = stage.ent(stn, k);
    if (pf < 2 && e < 0) continue;
    const int32_t f = (e > 0 ? e : -e) - 1;
    const int64_t o = stage.oth(stn, k);
    fcp_prefetch_l2(m.arx + f); fcp_prefetch_l2(m.ary + f); fcp_prefetch_l2(m.arz + f);
    if (stage.slot(stn, k, false) >= 0) {
      fcp_prefetch_l2(m.facint + f); fcp_prefetch_l2(m.Df + f);
      fcp_prefetch_l2(m.xc + o); fcp_prefetch_l2(m.yc + o); fcp_prefetch_l2(m.zc + o); fcp_prefetch_l2(g.den + o); fcp_prefetch_l2(m.vol + o);
      fcp_prefetch_l2(g.apu + o); fcp_prefetch_l2(g.u + o); fcp_prefetch_l2(g.v + o); fcp_prefetch_l2(g.w + o); fcp_prefetch_l2(g.p + o);
      fcp_prefetch_l2(g.dPdxi + 3 * o); fcp_prefetch_l2(g.dPdxi + 3 * o + 2);
    }
  }
}
template <bool PISO, int W, bool MPIF = false, int PF = 0>
__global__ void __launch_bounds__(FCP_TPB, ((W >= 3 || MPIF) ? 1 : W == 2 ? 2 : 3)) k_assemble_pcorr(MeshView m, AsmArgs g) {
  constexpr int WS = 6;
  constexpr int NS = PF ? 3 : 2;
  FCP_STAGE_DYN_N(WS, NS, stage);     // the face list of the NEXT cell(s) travels global -> shared while this cell's gathers are in flight
  FCP_STAGED_LOOP_BEGIN_N(NS, stage, m, m.n, c, st, stn, if (PF) assemble_pcorr_prefetch<WS, NS>(m, g, stage, stn, fcp_chunk_cell(j__ + 1), PF))
    const int64_t dpos = diag_pos(m, c);    // with the cell's first loads (a_rinfo is a DRAM stream), not in front of the last store
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c];
    const double denc = g.den[c], kc = m.vol[c] * g.apu[c];
    const double uc = g.u[c], vc = g.v[c], wc = g.w[c], pc = g.p[c];
    const double gcx = g.dPdxi[3 * (int64_t)c], gcy = g.dPdxi[3 * (int64_t)c + 1], gcz = g.dPdxi[3 * (int64_t)c + 2];
    double dg = 0.0, s = 0.0;
    FCP_FACE_BATCHES_STAGED(stage, st, m, c, W) {
      FCP_BATCH_LISTS_STAGED(stage, st, WS, m, W, e_, o_, sl_);
      double sx_[W], sy_[W], sz_[W], lam_[W], Df_[W], xo_[W], yo_[W], zo_[W], deno_[W], volo_[W], apuo_[W], uo_[W], vo_[W], wo_[W], po_[W],
          gox_[W], goy_[W], goz_[W];
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const int32_t f = (e_[k] > 0 ? e_[k] : -e_[k]) - 1;
        const bool on = e_[k] != 0, two = on && sl_[k] >= 0;
        const int32_t o = o_[k];
        sx_[k] = on ? __ldg(m.arx + f) : 0.0; sy_[k] = on ? __ldg(m.ary + f) : 0.0; sz_[k] = on ? __ldg(m.arz + f) : 0.0;
        lam_[k] = two ? __ldg(m.facint + f) : 0.0; Df_[k] = two ? __ldg(m.Df + f) : 0.0;
        xo_[k] = two ? __ldg(m.xc + o) : 0.0; yo_[k] = two ? __ldg(m.yc + o) : 0.0; zo_[k] = two ? __ldg(m.zc + o) : 0.0;
        deno_[k] = two ? __ldg(g.den + o) : 0.0; volo_[k] = two ? __ldg(m.vol + o) : 0.0; apuo_[k] = two ? __ldg(g.apu + o) : 0.0;
        // u, v, w, p cell values are not written by this kernel (only pressure-patch boundary slots are)
        uo_[k] = two ? __ldg(g.u + o) : 0.0; vo_[k] = two ? __ldg(g.v + o) : 0.0; wo_[k] = two ? __ldg(g.w + o) : 0.0; po_[k] = two ? __ldg(g.p + o) : 0.0;
        gox_[k] = two ? __ldg(g.dPdxi + 3 * (int64_t)o) : 0.0; goy_[k] = two ? __ldg(g.dPdxi + 3 * (int64_t)o + 1) : 0.0;
        goz_[k] = two ? __ldg(g.dPdxi + 3 * (int64_t)o + 2) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < W; ++k) {
        if (e_[k] == 0) continue;
        const int32_t e = e_[k], o = o_[k], sl = sl_[k];
        const int32_t f = (e > 0 ? e : -e) - 1;
        const double sx = sx_[k], sy = sy_[k], sz = sz_[k];
        if (sl >= 0) {
          const double xo = xo_[k], yo = yo_[k], zo = zo_[k];
          const double deno = deno_[k], ko = volo_[k] * apuo_[k];
          const double uo = uo_[k], vo = vo_[k], wo = wo_[k], po = po_[k];
          const double gox = gox_[k], goy = goy_[k], goz = goz_[k];
          const double lam = lam_[k], fxn = lam, fxp = 1.0 - lam;
          const bool own = e > 0;
          if (MPIF && o < m.n) {
            const int32_t cP = own ? c : o, cN = own ? o : c;
            const double xf = m.xf[f], yf = m.yf[f], zf = m.zf[f];
            const double xP = m.xc[cP], yP = m.yc[cP], zP = m.zc[cP], xN = m.xc[cN], yN = m.yc[cN], zN = m.zc[cN];
            const double xpn = xN - xP, ypn = yN - yP, zpn = zN - zP;
            const double are = sqrt(sx * sx + sy * sy + sz * sz);
            const double nxx = sx / are, nyy = sy / are, nzz = sz / are;
            const double volP = m.vol[cP], volN = m.vol[cN];
            const double Dpu = (fxn * volN * g.apu[cN] + fxp * volP * g.apu[cP]);
            const double Dpv = (fxn * volN * g.apv[cN] + fxp * volP * g.apv[cP]);
            const double Dpw = (fxn * volN * g.apw[cN] + fxp * volP * g.apw[cP]);
            const double dene = g.den[cP] * fxp + g.den[cN] * fxn;
            const double sfdpnr = 1. / (sx * xpn + sy * ypn + sz * zpn);
            const double smdpn = (sx * sx + sy * sy + sz * sz) * sfdpnr;
            const double cap = -dene * Dpu * smdpn;
            const double *gq[3] = {g.gU, g.gV, g.gW};
            const double *fq[3] = {g.u, g.v, g.w};
            double vi[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {            // face_value_central, src-par/interpolation.f90:148-176
              const double *gg = gq[q];
              const double gradfidr = gg[3 * (int64_t)cP] * (xf - xP) + gg[3 * (int64_t)cP + 1] * (yf - yP) + gg[3 * (int64_t)cP + 2] * (zf - zP) +
                                      gg[3 * (int64_t)cN] * (xf - xN) + gg[3 * (int64_t)cN + 1] * (yf - yN) + gg[3 * (int64_t)cN + 2] * (zf - zN);
              vi[q] = 0.5 * (fq[q][cP] + fq[q][cN] + gradfidr);
            }
            const double gPx = g.dPdxi[3 * (int64_t)cP], gPy = g.dPdxi[3 * (int64_t)cP + 1], gPz = g.dPdxi[3 * (int64_t)cP + 2];
            const double gNx = g.dPdxi[3 * (int64_t)cN], gNy = g.dPdxi[3 * (int64_t)cN + 1], gNz = g.dPdxi[3 * (int64_t)cN + 2];
            const double dpxi = Dpu * (fxn * gNx + fxp * gPx) * xpn * nxx;
            const double dpyi = Dpv * (fxn * gNy + fxp * gPy) * ypn * nyy;
            const double dpzi = Dpw * (fxn * gNz + fxp * gPz) * zpn * nzz;
            double xpp = xf - (xf - xP) * nxx, ypp = yf - (yf - yP) * nyy, zpp = zf - (zf - zP) * nzz;
            double xep = xf - (xf - xN) * nxx, yep = yf - (yf - yN) * nyy, zep = zf - (zf - zN) * nzz;
            xpp = xpp - xP; ypp = ypp - yP; zpp = zpp - zP;
            xep = xep - xN; yep = yep - yN; zep = zep - zN;
            double dpe = (g.p[cN] - g.p[cP]);
            const double dpecorr = (gNx * xep + gNy * yep + gNz * zep - gPx * xpp + gPy * ypp + gPz * zpp);     // quirk Q26 (:158-159)
            dpe = dpe + dpecorr;
            const double dpex = Dpu * dpe * sfdpnr * sx, dpey = Dpv * dpe * sfdpnr * sy, dpez = Dpw * dpe * sfdpnr * sz;
            const double ue = vi[0] - dpex + dpxi, ve = vi[1] - dpey + dpyi, we = vi[2] - dpez + dpzi;
            const double flm = dene * (ue * sx + ve * sy + we * sz);
            g.a[sl] = cap;
            dg = dg - cap;
            if (own) { s = s - flm; g.flmass[f] = flm; }
            else     { s = s + flm; }
            continue;
          }
          // P = owner side, N = neighbour side of the face, whichever this cell is
          const double xpn = own ? xo - xc : xc - xo, ypn = own ? yo - yc : yc - yo, zpn = own ? zo - zc : zc - zo;
          const double denP = own ? denc : deno, denN = own ? deno : denc;
          const double kP = own ? kc : ko, kN = own ? ko : kc;
          const double uP = own ? uc : uo, uN = own ? uo : uc;
          const double vP = own ? vc : vo, vN = own ? vo : vc;
          const double wP = own ? wc : wo, wN = own ? wo : wc;
          const double pP = own ? pc : po, pN = own ? po : pc;
          const double gPx = own ? gcx : gox, gPy = own ? gcy : goy, gPz = own ? gcz : goz;
          const double gNx = own ? gox : gcx, gNy = own ? goy : gcy, gNz = own ? goz : gcz;
          const double dene = denP * fxp + denN * fxn;
          const double Kj = kP * fxp + kN * fxn;
          const double cap = -dene * Kj * Df_[k];
          const double ui = uP + (uN - uP) * lam;
          const double vi = vP + (vN - vP) * lam;
          const double wi = wP + (wN - wP) * lam;
          const double dpxi = (gNx * fxp + gPx * fxn) * xpn;     // weights swapped in the reference (faceflux_mass.f90:236-238), kept
          const double dpyi = (gNy * fxp + gPy * fxn) * ypn;
          const double dpzi = (gNz * fxp + gPz * fxn) * zpn;
          const double flm = PISO ? dene * (ui * sx + vi * sy + wi * sz)
                                  : dene * (ui * sx + vi * sy + wi * sz) + cap * (pN - pP - dpxi - dpyi - dpzi);
          g.a[sl] = cap;
          dg = dg - cap;
          if (own) { s = s - flm; g.flmass[f] = flm; }
          else     { s = s + flm; }
        } else {
          const int type = -1 - sl;
          if (type == FCP_BC_INLET || type == FCP_BC_OUTLET) {
            s = s - g.flmass[f];                                // calcp_simple.f90:131-160
          } else if (type == FCP_BC_PRESSURE) {                 // facefluxmassPressBnd faceflux_mass.f90:765-831
            const double xpn = m.xf[f] - xc, ypn = m.yf[f] - yc, zpn = m.zf[f] - zc;
            const double capp = kc / (sx * xpn + sy * ypn + sz * zpn);
            const double dpcor = g.p[o] - pc - (gcx * xpn + gcy * ypn + gcz * zpn);
            const double ub = uc - sx * capp * dpcor, vb = vc - sy * capp * dpcor, wb = wc - sz * capp * dpcor;
            g.ub[o] = ub; g.vb[o] = vb; g.wb[o] = wb;
            const double flm = denc * (ub * sx + vb * sy + wb * sz);
            g.flmass[f] = flm;
            const double cap = -denc * (sx * sx + sy * sy + sz * sz) * capp;
            dg = dg - cap;
            s = s - flm;
            if (!PISO) g.pp[o] = 0.0;
          } else if (m.per_cell && (type == FCP_BC_PERIODIC || type == FCP_BC_EMPTY) && m.per_cell[f - m.F] >= 0) {
            // ---- facefluxmass2_periodic, faceflux_mass.f90:313-384 (calcp_simple.f90:185-228, calcp_piso.f90:248-293): evaluated
            // in the orientation of the PERIODIC face fp (P = its owner, N = the owner of the twin face) on both sides of the pair
            const int32_t b = f - m.F, q = m.per_cell[b];
            const bool own = type == FCP_BC_PERIODIC;
            const int32_t fp = own ? f : m.per_face[b];
            const int32_t cP = own ? c : q, cN = own ? q : c;
            const double ax = m.arx[fp], ay = m.ary[fp], az = m.arz[fp];
            const double fxn = 0.5, fxp = 1.0 - 0.5;
            const double xpn = 2 * (m.xf[fp] - m.xc[cP]), ypn = 2 * (m.yf[fp] - m.yc[cP]), zpn = 2 * (m.zf[fp] - m.zc[cP]);
            const double apuP = g.apu[cP], apuN = g.apu[cN];
            const double dene = g.den[cP] * fxp + g.den[cN] * fxn;
            double Kj = m.vol[cP] * apuP * fxp + m.vol[cN] * apuN * fxn;
            const double cap = -dene * Kj * m.per_df[b];      // quirk Q21: Df(i), i = the face's ordinal inside its patch
            Kj = (apuP + apuN + FCP_SMALL);
            const double ui = (g.u[cP] * apuN + g.u[cN] * apuP) / Kj;
            const double vi = (g.v[cP] * g.apv[cN] + g.v[cN] * g.apv[cP]) / Kj;
            const double wi = (g.w[cP] * g.apw[cN] + g.w[cN] * g.apw[cP]) / Kj;
            const double dpxi = (g.dPdxi[3 * (int64_t)cN] * fxp + g.dPdxi[3 * (int64_t)cP] * fxn) * xpn;
            const double dpyi = (g.dPdxi[3 * (int64_t)cN + 1] * fxp + g.dPdxi[3 * (int64_t)cP + 1] * fxn) * ypn;
            const double dpzi = (g.dPdxi[3 * (int64_t)cN + 2] * fxp + g.dPdxi[3 * (int64_t)cP + 2] * fxn) * zpn;
            const double flm = dene * (ui * ax + vi * ay + wi * az) + cap * (g.p[cN] - g.p[cP] - dpxi - dpyi - dpzi);
            g.a[m.per_slot[b]] = cap;
            dg = dg - cap;
            if (own) { s = s - flm; g.flmass[fp] = flm; }
            else     { s = s + flm; }
          }
        }
      }
    }
    g.a[dpos] = dg;
    g.su[c] = s;
  FCP_STAGED_LOOP_END
}

// adjustMassFlow faceflux_mass.f90:833-916 (src-par/adjustMassFlow.f90: `call global_sum(flowo)` between the two loops).  Outlet patches are
// small; ONE CTA evaluates the outlet fluxes in parallel and thread 0 adds them in face order (the reference's order) into *flowo; after the
// cross-rank sum (rank order, comm.cu) the second kernel scales.  A rank without outlet faces runs both with nout = 0 and contributes 0.
__global__ void __launch_bounds__(FCP_TPB) k_adjust_mass_flow_flux(int32_t nout, const int32_t *__restrict__ oface, int32_t n, int32_t F,
                                                                    const int32_t *__restrict__ owner, const double *__restrict__ arx,
                                                                    const double *__restrict__ ary, const double *__restrict__ arz,
                                                                    const double *__restrict__ den, double *u, double *v, double *w,
                                                                    double *flmass, double *flowo_out) {
  for (int32_t i = threadIdx.x; i < nout; i += blockDim.x) {
    const int32_t f = oface[i], ijp = owner[f], ijb = n + (f - F);
    const double ub = u[ijp], vb = v[ijp], wb = w[ijp];
    u[ijb] = ub; v[ijb] = vb; w[ijb] = wb;
    flmass[f] = den[ijp] * (ub * arx[f] + vb * ary[f] + wb * arz[f]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double flowo = 0.0;
    for (int32_t i = 0; i < nout; ++i) flowo = flowo + flmass[oface[i]];
    *flowo_out = flowo;
  }
}
__global__ void __launch_bounds__(FCP_TPB) k_adjust_mass_flow_scale(int32_t nout, const int32_t *__restrict__ oface, int32_t n, int32_t F,
                                                                     double *u, double *v, double *w, double *flmass, double flomas,
                                                                     const double *__restrict__ flowo) {
  const double fac = flomas / (*flowo + FCP_SMALL);
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nout; i += gridDim.x * blockDim.x) {
    const int32_t f = oface[i], ijb = n + (f - F);
    flmass[f] = flmass[f] * fac;
    u[ijb] = u[ijb] * fac; v[ijb] = v[ijb] * fac; w[ijb] = w[ijb] * fac;
  }
}

// flux correction  calcp_simple.f90:331-341 : face-parallel, no scatter
__global__ void __launch_bounds__(FCP_TPB) k_correct_flux(int32_t F, const int32_t *__restrict__ owner, const int32_t *__restrict__ neigh,
                                                           const int32_t *__restrict__ kPN, const double *__restrict__ a,
                                                           const double *__restrict__ pp, double *__restrict__ flmass) {
  FCP_CELL_LOOP(f, F) { flmass[f] = flmass[f] + a[kPN[f]] * (pp[neigh[f]] - pp[owner[f]]); }
}
// periodic pairs  calcp_simple.f90:350-373 / calcp_piso.f90:441-460 : flmass(if) += a(k) (x(ijn) - x(ijp)), flmass(iftwin) = flmass(if)
__global__ void __launch_bounds__(FCP_TPB) k_correct_flux_periodic(MeshView m, const int32_t *__restrict__ bftype, const double *__restrict__ a,
                                                                    const double *__restrict__ x, double *__restrict__ flmass) {
  FCP_CELL_LOOP(i, m.B) {
    if (bftype[i] != FCP_BC_PERIODIC) continue;
    const int32_t f = m.F + i;
    const double fl = flmass[f] + a[m.per_slot[i]] * (x[m.per_cell[i]] - x[m.owner[f]]);
    flmass[f] = fl;
    flmass[m.per_face[i]] = fl;
  }
}
// pressure patches  calcp_simple.f90:345-391 + facefluxmassCorrPressBnd faceflux_mass.f90:699-762
__global__ void __launch_bounds__(FCP_TPB) k_correct_pressure_bnd(MeshView m, const int32_t *__restrict__ bftype, const double *__restrict__ den,
                                                                   const double *__restrict__ apu, const double *__restrict__ pp,
                                                                   double *u, double *v, double *w, double *flmass) {
  FCP_CELL_LOOP(i, m.B) {
    if (bftype[i] != FCP_BC_PRESSURE) continue;
    const int32_t f = m.F + i, ijp = m.owner[f], ijb = m.n + i;
    const double sx = m.arx[f], sy = m.ary[f], sz = m.arz[f];
    const double xpn = m.xf[f] - m.xc[ijp], ypn = m.yf[f] - m.yc[ijp], zpn = m.zf[f] - m.zc[ijp];
    const double cap = m.vol[ijp] * apu[ijp] / (sx * xpn + sy * ypn + sz * zpn);
    const double dpcor = -pp[ijp];
    u[ijb] = u[ijb] - sx * cap * dpcor;
    v[ijb] = v[ijb] - sy * cap * dpcor;
    w[ijb] = w[ijb] - sz * cap * dpcor;
    flmass[f] = flmass[f] - den[ijp] * (sx * sx + sy * sy + sz * sz) * cap * dpcor;
  }
}

// non-orthogonal corrector  calcp_simple.f90:433-455 + fluxmc2 faceflux_mass.f90:650-696
__global__ void __launch_bounds__(FCP_TPB) k_nonorth(MeshView m, const double *__restrict__ den, const double *__restrict__ apu,
                                                      const double *__restrict__ dPdxi, double *__restrict__ su, double *__restrict__ flmass) {
  FCP_CELL_LOOP(c, m.n) {
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c];
    const double rc = apu[c] * den[c];
    const double gcx = dPdxi[3 * (int64_t)c], gcy = dPdxi[3 * (int64_t)c + 1], gcz = dPdxi[3 * (int64_t)c + 2];
    double s = 0.0;
    FCP_FACE_LOOP(m, c) {
      FCP_FACE_FETCH(m);
      if (sl >= 0) {
        const bool own = e > 0;
        const double sx = m.arx[f], sy = m.ary[f], sz = m.arz[f];
        const double xo = m.xc[o], yo = m.yc[o], zo = m.zc[o];
        const double ro = apu[o] * den[o];
        const double gox = dPdxi[3 * (int64_t)o], goy = dPdxi[3 * (int64_t)o + 1], goz = dPdxi[3 * (int64_t)o + 2];
        const double xpn = own ? xo - xc : xc - xo, ypn = own ? yo - yc : yc - yo, zpn = own ? zo - zc : zc - zo;
        const double rP = own ? rc : ro, rN = own ? ro : rc;
        const double gPx = own ? gcx : gox, gPy = own ? gcy : goy, gPz = own ? gcz : goz;
        const double gNx = own ? gox : gcx, gNy = own ? goy : gcy, gNz = own ? goz : gcz;
        const double s2 = sx * sx + sy * sy + sz * sz;
        const double dn = xpn * sx + ypn * sy + zpn * sz;
        const double rapr = -0.5 * (rP + rN);
        const double dpx = 0.5 * (gNx + gPx), dpy = 0.5 * (gNy + gPy), dpz = 0.5 * (gNz + gPz);
        const double fmcor = rapr * ((dn * sx - xpn * s2) * dpx + (dn * sy - ypn * s2) * dpy + (dn * sz - zpn * s2) * dpz);
        if (own) { flmass[f] = flmass[f] + fmcor; s = s - fmcor; }
        else     { s = s + fmcor; }
      }
    }
    su[c] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// host wrappers (called from api.cu)
// ---------------------------------------------------------------------------------------------

int fvm_grad_gauss(fcp_ctx *ctx, const double *u, double *g) {
  if (ctx->B) FCP_CUDA(cudaMemsetAsync(g + 3 * (size_t)ctx->n, 0, sizeof(double) * 3 * (size_t)ctx->B, ctx->stream));
  if (ctx->n == 0) return FCP_OK;
  size_t smem = 0;
  const int grid = std::max(fcp_nchunks(ctx->n), 1);
  MeshView mv = fcp_mesh_view(ctx);
  const FaceVariant fv = fcp_face_variant(FCP_FK_GRAD_GAUSS);
  FCP_TRY(fcp_apply_face_variant(ctx, fv, mv, true, true));
#define GG_LAUNCH(WS, MINB, PF)                                                                                                      \
  do {                                                                                                                               \
    FCP_TRY((fcp_stage_smem<WS, (PF) ? 3 : 2>(k_grad_gauss<WS, MINB, PF>, &smem)));                                                  \
    FCP_PROF(&ctx->prof, FCP_K_GRAD, ctx->stream, (k_grad_gauss<WS, MINB, PF><<<grid, FCP_TPB, smem, ctx->stream>>>(mv, u, g)));     \
  } while (0)
#define GG_LAUNCH_WS(WS)                                            \
  do {                                                              \
    if (fv.occ >= 3) {                                              \
      if (fv.pf == 2) GG_LAUNCH(WS, 3, 2);                          \
      else if (fv.pf == 1) GG_LAUNCH(WS, 3, 1);                     \
      else GG_LAUNCH(WS, 3, 0);                                     \
    } else {                                                        \
      if (fv.pf == 2) GG_LAUNCH(WS, 2, 2);                          \
      else if (fv.pf == 1) GG_LAUNCH(WS, 2, 1);                     \
      else GG_LAUNCH(WS, 2, 0);                                     \
    }                                                               \
  } while (0)
  if (ctx->max_cell_faces > 6) GG_LAUNCH_WS(10);
  else GG_LAUNCH_WS(6);
#undef GG_LAUNCH_WS
#undef GG_LAUNCH
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_lsq_matrix(fcp_ctx *ctx, bool weighted, double *D) {
  if (ctx->n == 0) return FCP_OK;
  if (weighted) k_lsq_matrix<true><<<FCP_GRID(ctx->n)>>>(fcp_mesh_view(ctx), D);
  else k_lsq_matrix<false><<<FCP_GRID(ctx->n)>>>(fcp_mesh_view(ctx), D);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_grad_lsq(fcp_ctx *ctx, bool weighted, const double *D, const double *phi, double *g, int row2_reference) {
  if (ctx->B) FCP_CUDA(cudaMemsetAsync(g + 3 * (size_t)ctx->n, 0, sizeof(double) * 3 * (size_t)ctx->B, ctx->stream));
  if (ctx->n == 0) return FCP_OK;
  size_t tok = ctx->prof.begin(FCP_K_GRAD, ctx->stream);
  const bool wide = ctx->max_cell_faces > 6;
  size_t smem = 0;
  const int grid = std::max(fcp_nchunks(ctx->n), 1);
  MeshView mv = fcp_mesh_view(ctx);
  const FaceVariant fv = fcp_face_variant(FCP_FK_GRAD_LSQ);
  FCP_TRY(fcp_apply_face_variant(ctx, fv, mv, true, false));   // (quirk Q2 does arithmetic on the face index: no owner-ordered geometry)
#define LSQ_LAUNCH(WT, WS, MINB, PF)                                                                        \
  do {                                                                                                      \
    FCP_TRY((fcp_stage_smem<WS, (PF) ? 3 : 2>(k_grad_lsq<WT, WS, MINB, PF>, &smem)));                       \
    k_grad_lsq<WT, WS, MINB, PF><<<grid, FCP_TPB, smem, ctx->stream>>>(mv, D, phi, g, row2_reference);      \
  } while (0)
#define LSQ_LAUNCH_V(WT, WS)                                         \
  do {                                                              \
    if (fv.occ >= 3) {                                              \
      if (fv.pf == 2) LSQ_LAUNCH(WT, WS, 3, 2);                     \
      else if (fv.pf == 1) LSQ_LAUNCH(WT, WS, 3, 1);                \
      else LSQ_LAUNCH(WT, WS, 3, 0);                                \
    } else {                                                        \
      if (fv.pf == 2) LSQ_LAUNCH(WT, WS, 2, 2);                     \
      else if (fv.pf == 1) LSQ_LAUNCH(WT, WS, 2, 1);                \
      else LSQ_LAUNCH(WT, WS, 2, 0);                                \
    }                                                               \
  } while (0)
  if (weighted && wide) LSQ_LAUNCH_V(true, 10);
  else if (weighted) LSQ_LAUNCH_V(true, 6);
  else if (wide) LSQ_LAUNCH_V(false, 10);
  else LSQ_LAUNCH_V(false, 6);
#undef LSQ_LAUNCH_V
#undef LSQ_LAUNCH
  ctx->prof.end(tok, ctx->stream);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_laplacian(fcp_ctx *ctx, const double *mu, const double *phi, double *a, double *su) {
  if (ctx->n == 0) return FCP_OK;
  FCP_PROF(&ctx->prof, FCP_K_LAPLACIAN, ctx->stream, (k_laplacian<<<FCP_GRID(ctx->n)>>>(fcp_mesh_view(ctx), mu, phi, a, su)));
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
// p: the field whose gradient is taken (p or pp); correct != nullptr fuses calcp_simple.f90:416-429
int fvm_gradp(fcp_ctx *ctx, int scheme, double *p, const double *apu, double *su, double *sv, double *sw, double *dPdxi, double *gtmp,
              const CorrectArgs *correct) {
  if (ctx->n == 0) return FCP_OK;
  MeshView m = fcp_mesh_view(ctx);
  CorrectArgs ca{};
  if (correct) ca = *correct;
  size_t smem = 0;
  const int grid = std::max(fcp_nchunks(ctx->n), 1);
  const FaceVariant fv = fcp_face_variant(FCP_FK_GRADP);
  MeshView mg = m;       // the view of k_gradp (compact lists, owner-ordered geometry); k_gradp_central2 keeps the plain one
  FCP_TRY(fcp_apply_face_variant(ctx, fv, mg, true, true));
#define GRADP_LAUNCH(C, WG, PF, NST, OUT)                                                                          \
  do {                                                                                                             \
    FCP_TRY((fcp_stage_smem<6, (PF) ? 3 : 2>(k_gradp<C, WG, PF>, &smem)));                                         \
    k_gradp<C, WG, PF><<<grid, FCP_TPB, smem, ctx->stream>>>(mg, NST, p, apu, su, sv, sw, OUT, ca);                 \
  } while (0)
#define GRADP_LAUNCH_V(C, WG, NST, OUT)                             \
  do {                                                              \
    if (fv.pf == 2) GRADP_LAUNCH(C, WG, 2, NST, OUT);               \
    else if (fv.pf == 1) GRADP_LAUNCH(C, WG, 1, NST, OUT);          \
    else GRADP_LAUNCH(C, WG, 0, NST, OUT);                          \
  } while (0)
  if (scheme == FCP_PSCHEME_CENTRAL) {
    GRADP_LAUNCH_V(false, false, 1, gtmp);
    FCP_LAUNCHED();
    if (correct) k_gradp_central2<true><<<FCP_GRID(ctx->n)>>>(m, p, gtmp, su, sv, sw, dPdxi, ca);
    else k_gradp_central2<false><<<FCP_GRID(ctx->n)>>>(m, p, gtmp, su, sv, sw, dPdxi, ca);
    FCP_LAUNCHED();
  } else {
    size_t tok = ctx->prof.begin(FCP_K_GRADP, ctx->stream);
    const bool wgt = scheme == FCP_PSCHEME_WEIGHTED;
    if (correct && wgt) GRADP_LAUNCH_V(true, true, 2, dPdxi);
    else if (correct) GRADP_LAUNCH_V(true, false, 2, dPdxi);
    else if (wgt) GRADP_LAUNCH_V(false, true, 2, dPdxi);
    else GRADP_LAUNCH_V(false, false, 2, dPdxi);
    ctx->prof.end(tok, ctx->stream);
    FCP_LAUNCHED();
  }
#undef GRADP_LAUNCH_V
#undef GRADP_LAUNCH
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_assemble_pcorr(fcp_ctx *ctx, const AsmArgs &g, bool piso) {
  if (ctx->n == 0) return FCP_OK;
  int w = 2;           // faces per register batch: FCP_ASM_W = 1 | 2 | 3 (A/B measurements)
  if (const char *e = getenv("FCP_ASM_W")) { w = atoi(e); if (w < 1 || w > 3) w = 2; }
  const FaceVariant fv = fcp_face_variant(FCP_FK_ASSEMBLE);
  size_t smem = 0;
  const int grid = std::max(fcp_nchunks(ctx->n), 1);
  size_t tok = ctx->prof.begin(FCP_K_ASSEMBLE, ctx->stream);
  MeshView mv = fcp_mesh_view(ctx);
  FCP_TRY(fcp_apply_face_variant(ctx, fv, mv, false, false));
#define ASM_LAUNCH(PISO, W, MPIF, PF)                                                                   \
  do {                                                                                                  \
    FCP_TRY((fcp_stage_smem<6, (PF) ? 3 : 2>(k_assemble_pcorr<PISO, W, MPIF, PF>, &smem)));             \
    k_assemble_pcorr<PISO, W, MPIF, PF><<<grid, FCP_TPB, smem, ctx->stream>>>(mv, g);                   \
  } while (0)
#define ASM_LAUNCH_V(PISO, W)                                       \
  do {                                                              \
    if (fv.pf == 2) ASM_LAUNCH(PISO, W, false, 2);                  \
    else if (fv.pf == 1) ASM_LAUNCH(PISO, W, false, 1);             \
    else ASM_LAUNCH(PISO, W, false, 0);                             \
  } while (0)
  if (g.gU) {          // quirk Q10 switch (fcp_set_flux_variant): SIMPLE only, one face per gather round
    ASM_LAUNCH(false, 1, true, 0);
  } else if (piso) {
    if (w == 1) ASM_LAUNCH_V(true, 1);
    else if (w == 2) ASM_LAUNCH_V(true, 2);
    else ASM_LAUNCH(true, 3, false, 0);
  } else {
    if (w == 1) ASM_LAUNCH_V(false, 1);
    else if (w == 2) ASM_LAUNCH_V(false, 2);
    else ASM_LAUNCH(false, 3, false, 0);
  }
#undef ASM_LAUNCH_V
#undef ASM_LAUNCH
  ctx->prof.end(tok, ctx->stream);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_adjust_mass_flow(fcp_ctx *ctx, int32_t nout, const int32_t *d_oface, const double *den, double *u, double *v, double *w,
                         double *flmass, double flomas) {
  if (!ctx->d_flowo) FCP_TRY(dev_alloc(&ctx->d_flowo, 4));
  k_adjust_mass_flow_flux<<<1, FCP_TPB, 0, ctx->stream>>>(nout, d_oface, ctx->n, ctx->F, ctx->owner, ctx->arx, ctx->ary, ctx->arz, den, u, v, w,
                                                           flmass, ctx->d_flowo);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  if (ctx->comm) FCP_TRY(comm_allgather_sum(ctx->comm, ctx->d_flowo, 1, ctx->stream));   // src-par/adjustMassFlow.f90:55  call global_sum(flowo)
  k_adjust_mass_flow_scale<<<std::max(1, std::min(64, (nout + FCP_TPB - 1) / FCP_TPB)), FCP_TPB, 0, ctx->stream>>>(nout, d_oface, ctx->n, ctx->F, u, v, w, flmass,
                                                                                                                     flomas, ctx->d_flowo);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_correct_flux(fcp_ctx *ctx, const double *a, const double *pp, double *flmass) {
  if (ctx->F == 0) return FCP_OK;
  FCP_PROF(&ctx->prof, FCP_K_CORRECT_FLUX, ctx->stream, (k_correct_flux<<<FCP_GRID(ctx->F)>>>(ctx->F, ctx->owner, ctx->neigh, ctx->kPN, a, pp, flmass)));
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_correct_flux_periodic(fcp_ctx *ctx, const double *a, const double *x, double *flmass) {
  if (!ctx->nper) return FCP_OK;
  k_correct_flux_periodic<<<FCP_GRID(ctx->B)>>>(fcp_mesh_view(ctx), ctx->bftype, a, x, flmass);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_correct_pressure_bnd(fcp_ctx *ctx, const double *den, const double *apu, const double *pp, double *u, double *v, double *w,
                             double *flmass) {
  if (ctx->B == 0) return FCP_OK;
  k_correct_pressure_bnd<<<FCP_GRID(ctx->B)>>>(fcp_mesh_view(ctx), ctx->bftype, den, apu, pp, u, v, w, flmass);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_nonorth(fcp_ctx *ctx, const double *den, const double *apu, const double *dPdxi, double *su, double *flmass) {
  if (ctx->n == 0) return FCP_OK;
  k_nonorth<<<FCP_GRID(ctx->n)>>>(fcp_mesh_view(ctx), den, apu, dPdxi, su, flmass);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
