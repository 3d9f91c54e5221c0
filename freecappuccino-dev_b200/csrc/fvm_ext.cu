// fvm_ext.cu -- the remaining cell-parallel operators of the hot path (same gather design as fvm.cu):
//   slope limiters          gradients.f90:288-656      (Barth-Jespersen, Venkatakrishnan, 'R3', multidimensional)
//   QR least-squares grad   gradients.f90:900-1152     (+ misc/matrix.f90:137-167 inv, :366-419 mgs_qr)
//   calcp_piso pieces       Pressure/calcp_piso.f90:81-489, fluxes/faceflux_mass.f90:564-647 (fluxmc)
// Every loop of the reference that looks face-sequential here is per-cell independent: a cell's value is only ever
// modified by its own faces, in ascending face order -- which is exactly what a thread walking the cell's SELL face
// list reproduces.
#include "fcp_internal.h"
#include "reduce.cuh"
#include "fvm_common.cuh"

// ---------------------------------------------------------------------------------------------
// global extrema of phi(1:numCells)   gradients.f90:317-318 (minval / maxval are exact: any order gives the same bits)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FCP_TPB) k_minmax_part(int32_t n, const double *__restrict__ phi, double *__restrict__ part) {
  double lo = INFINITY, hi = -INFINITY;
  FCP_CELL_LOOP(c, n) {
    const double v = phi[c];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
  __shared__ double slo[FCP_TPB / 32], shi[FCP_TPB / 32];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < FCP_TPB / 32; ++w) { lo = fmin(lo, slo[w]); hi = fmax(hi, shi[w]); }
    part[2 * blockIdx.x] = lo;
    part[2 * blockIdx.x + 1] = hi;
  }
}
__global__ void __launch_bounds__(FCP_TPB) k_minmax_final(int nparts, const double *__restrict__ part, double *__restrict__ mm) {
  double lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < nparts; i += FCP_TPB) { lo = fmin(lo, part[2 * i]); hi = fmax(hi, part[2 * i + 1]); }
  __shared__ double slo[FCP_TPB / 32], shi[FCP_TPB / 32];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < FCP_TPB / 32; ++w) { lo = fmin(lo, slo[w]); hi = fmax(hi, shi[w]); }
    mm[0] = lo;
    mm[1] = hi;
  }
}

// ---------------------------------------------------------------------------------------------
// Barth-Jespersen :288-373, Venkatakrishnan :378-461, 'R3' :464-552 (active formula: R4).  One thread = one cell, walks
// its CSR row (SELL, columns ascending) skipping the diagonal -- the reference's `do k=ia(inp),ia(inp+1)-1` loop.
// ---------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ double limiter_fn(double r) {
  if (KIND == FCP_LIMITER_BARTH_JESPERSEN) return r;
  const double r2 = r * r;
  if (KIND == FCP_LIMITER_VENKATAKRISHNAN) return (r2 + 2.0 * r) / (r2 + r + 2.0);
  const double r3 = r2 * r, r4 = r2 * r2;
  return (r4 + 2.0 * r3 - 4.0 * r2 + 8.0 * r) / (r4 + r3 + 2.0 * r2 - 4.0 * r + 8.0);
}

template <int KIND>
__global__ void __launch_bounds__(FCP_TPB) k_limiter_cell(MeshView m, const double *__restrict__ phi, const double *__restrict__ mm,
                                                           double *__restrict__ g) {
  const double fimin = mm[0], fimax = mm[1];
  const double eps = (double)1.e-6f;   // `1.e-6`: default-real literal
  FCP_CELL_LOOP(c, m.n) {
    const double gx = g[3 * (int64_t)c], gy = g[3 * (int64_t)c + 1], gz = g[3 * (int64_t)c + 2];
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c], pc = phi[c];
    const double deltamax = fimax - pc, deltamin = fimin - pc;
    const int64_t base = m.a_slptr[c >> 5] + (c & 31);
    const int32_t ri = m.a_rinfo[c];
    const int32_t dpos = (ri >> 16) & 0xffff;
    const int32_t len = ri & 0xffff;     // the whole row: on a partition the cells across process faces are neighbours like any other (ghost copies)
    double slopelimit = 1.0;
    // (a version that gathered a whole row before the arithmetic -- 8 column indices, then 24 centre loads -- was measured SLOWER on a B200, 0.91 ms
    // against 0.79 ms at 256^3: 102 registers instead of 48 cost more occupancy than the batching bought)
    for (int32_t k = 0; k < len; ++k) {
      if (k == dpos) continue;
      const int32_t j = m.a_ja[base + (int64_t)k * 32];
      const double delta_face = gx * (m.xc[j] - xc) + gy * (m.yc[j] - yc) + gz * (m.zc[j] - zc);
      double r;
      if (fabs(delta_face) < eps) r = 1.0;
      else if (delta_face > 0.0) r = deltamax / delta_face;
      else r = deltamin / delta_face;
      slopelimit = fmin(slopelimit, limiter_fn<KIND>(r));
    }
    g[3 * (int64_t)c] = slopelimit * gx;
    g[3 * (int64_t)c + 1] = slopelimit * gy;
    g[3 * (int64_t)c + 2] = slopelimit * gz;
  }
}

// multidimensional limiter :556-656: local extrema over the CSR row (the cell itself included), then the cell's INNER
// faces in ascending face order, each possibly replacing the running gradient.
__global__ void __launch_bounds__(FCP_TPB) k_limiter_mdl(MeshView m, const double *__restrict__ phi, double *__restrict__ g) {
  FCP_CELL_LOOP(c, m.n) {
    const int64_t base = m.a_slptr[c >> 5] + (c & 31);
    const int32_t ri = m.a_rinfo[c];
    const int32_t len = ri & 0xffff;     // whole row, halo columns included
    const double pc = phi[c];
    double phimax = phi[m.a_ja[base]], phimin = phimax;
    for (int32_t k = 1; k < len; ++k) {
      const double v = phi[m.a_ja[base + (int64_t)k * 32]];
      phimax = fmax(phimax, v);
      phimin = fmin(phimin, v);
    }
    const double dPhimax = phimax - pc, dPhimin = phimin - pc;
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c];
    double gx = g[3 * (int64_t)c], gy = g[3 * (int64_t)c + 1], gz = g[3 * (int64_t)c + 2];
    FCP_FACE_LOOP(m, c) {
      FCP_FACE_FETCH(m);
      if (f >= m.F && sl < 0) continue;   // inner faces only (:588); a process face is an inner face of the unpartitioned mesh
      const double xpn = m.xf[f] - xc, ypn = m.yf[f] - yc, zpn = m.zf[f] - zc;
      const double dpn = sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
      const double nx = xpn / dpn, ny = ypn / dpn, nz = zpn / dpn;
      const double gn = gx * nx + gy * ny + gz * nz;
      const double gtx = gx - gn * nx, gty = gy - gn * ny, gtz = gz - gn * nz;
      const double dPhi = gx * xpn + gy * ypn + gz * zpn;
      if (phimax > pc && dPhi > dPhimax) { gx = gtx + nx * dPhimax; gy = gty + ny * dPhimax; gz = gtz + nz * dPhimax; }
      if (phimin < pc && dPhi < dPhimin) { gx = gtx + nx * dPhimin; gy = gty + ny * dPhimin; gz = gtz + nz * dPhimin; }
    }
    g[3 * (int64_t)c] = gx;
    g[3 * (int64_t)c + 1] = gy;
    g[3 * (int64_t)c + 2] = gz;
  }
}

// global minimum / maximum of phi(1:numCells): mm_out points at the device pair {min, max}
int fvm_minmax(fcp_ctx *ctx, const double *phi, double **mm_out, int32_t count) {
  const int32_t cnt = count < 0 ? ctx->n : count;
  const int nparts = std::max(fcp_nchunks(cnt), 1);
  const int cap = std::max(fcp_nchunks(ctx->nT), 1);          // the partial buffer is sized for the longest field
  if (!ctx->d_mmpart) FCP_TRY(dev_alloc(&ctx->d_mmpart, (size_t)2 * cap + 2));
  double *mm = ctx->d_mmpart + (size_t)2 * cap;
  k_minmax_part<<<nparts, FCP_TPB, 0, ctx->stream>>>(cnt, phi, ctx->d_mmpart);
  k_minmax_final<<<1, FCP_TPB, 0, ctx->stream>>>(nparts, ctx->d_mmpart, mm);
  FCP_LAUNCHED(); FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  if (ctx->comm) FCP_TRY(comm_allreduce_minmax(ctx->comm, mm, ctx->stream));
  *mm_out = mm;
  return FCP_OK;
}

int fvm_slope_limiter(fcp_ctx *ctx, int limiter, const double *phi, double *g) {
  if (limiter == FCP_LIMITER_NONE || ctx->n == 0 && !ctx->comm) return FCP_OK;
  MeshView m = fcp_mesh_view(ctx);
  size_t tok = ctx->prof.begin(FCP_K_LIMITER, ctx->stream);
  if (limiter == FCP_LIMITER_MULTIDIMENSIONAL) {
    if (ctx->n) { k_limiter_mdl<<<FCP_GRID(ctx->n)>>>(m, phi, g); FCP_LAUNCHED(); }
  } else {
    double *mm = nullptr;                                    // global extrema (quirk Q3), rank-reduced inside
    FCP_TRY(fvm_minmax(ctx, phi, &mm));
    if (ctx->n) {
      if (limiter == FCP_LIMITER_BARTH_JESPERSEN) k_limiter_cell<FCP_LIMITER_BARTH_JESPERSEN><<<FCP_GRID(ctx->n)>>>(m, phi, mm, g);
      else if (limiter == FCP_LIMITER_VENKATAKRISHNAN) k_limiter_cell<FCP_LIMITER_VENKATAKRISHNAN><<<FCP_GRID(ctx->n)>>>(m, phi, mm, g);
      else k_limiter_cell<FCP_LIMITER_R3><<<FCP_GRID(ctx->n)>>>(m, phi, mm, g);
      FCP_LAUNCHED();
    }
  }
  ctx->prof.end(tok, ctx->stream);
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

// ---------------------------------------------------------------------------------------------
// QR least squares   create_matrix_lsq_qr gradients.f90:900-1052, grad_lsq_qr :1057-1152.
// D(3,6,numCells) of the reference is stored SoA here: D[(l*3 + i) * n + c].
// Quirk Q20: the reference's call of mgs_qr passes mis-shaped actuals, its R1 is partly uninitialised; the
// documented algorithm (thin QR by modified Gram-Schmidt, R1^-1 Q1^T) is what is built.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void inv3(const double (&a)[3][3], double (&r)[3][3]) {   // misc/matrix.f90:146-165, literal
#define A(i, j) a[i - 1][j - 1]
#define DET (A(1,1)*A(2,2)*A(3,3) - A(1,1)*A(2,3)*A(3,2) - A(1,2)*A(2,1)*A(3,3) + A(1,2)*A(2,3)*A(3,1) + A(1,3)*A(2,1)*A(3,2) - A(1,3)*A(2,2)*A(3,1))
  r[0][0] = (A(2,2)*A(3,3) - A(2,3)*A(3,2)) / DET;
  r[0][1] = -(A(1,2)*A(3,3) - A(1,3)*A(3,2)) / DET;
  r[0][2] = (A(1,2)*A(2,3) - A(1,3)*A(2,2)) / DET;
  r[1][0] = -(A(2,1)*A(3,3) - A(2,3)*A(3,1)) / DET;
  r[1][1] = (A(1,1)*A(3,3) - A(1,3)*A(3,1)) / DET;
  r[1][2] = -(A(1,1)*A(2,3) - A(1,3)*A(2,1)) / DET;
  r[2][0] = (A(2,1)*A(3,2) - A(2,2)*A(3,1)) / DET;
  r[2][1] = -(A(1,1)*A(3,2) - A(1,2)*A(3,1)) / DET;
  r[2][2] = (A(1,1)*A(2,2) - A(1,2)*A(2,1)) / DET;
#undef DET
#undef A
}

__global__ void __launch_bounds__(FCP_TPB) k_lsq_qr_matrix(MeshView m, double *__restrict__ D) {
  FCP_CELL_LOOP(c, m.n) {
    double q[6][3];
#pragma unroll
    for (int i = 0; i < 6; ++i) { q[i][0] = 0.0; q[i][1] = 0.0; q[i][2] = 0.0; }
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c];
    {
      FCP_FACE_LOOP(m, c) {   // :962-991: neighbour_index order = ascending face index (inner faces, then boundary faces)
        FCP_FACE_FETCH(m);
        double dx, dy, dz;
        if (sl >= 0) { dx = m.xc[o] - xc; dy = m.yc[o] - yc; dz = m.zc[o] - zc; }
        else         { dx = m.xf[f] - xc; dy = m.yf[f] - yc; dz = m.zf[f] - zc; }
#pragma unroll
        for (int i = 0; i < 6; ++i)
          if (i == q__) { q[i][0] = dx; q[i][1] = dy; q[i][2] = dz; }
      }
    }
    double r[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
    for (int j = 0; j < 3; ++j) {   // mgs_qr, misc/matrix.f90:393-416
      double z = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) z = z + q[i][j] * q[i][j];
      r[j][j] = sqrt(z);
#pragma unroll
      for (int i = 0; i < 6; ++i) q[i][j] = q[i][j] / r[j][j];
#pragma unroll
      for (int k = j + 1; k < 3; ++k) {
        z = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) z = z + q[i][j] * q[i][k];
        r[j][k] = z;
#pragma unroll
        for (int i = 0; i < 6; ++i) q[i][k] = q[i][k] - r[j][k] * q[i][j];
      }
    }
    double ri[3][3];
    inv3(r, ri);
    const int64_t n = m.n;
#pragma unroll
    for (int l = 0; l < 6; ++l)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double x = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) x = x + ri[i][k] * q[l][k];
        D[(int64_t)(l * 3 + i) * n + c] = x;
      }
  }
}

__global__ void __launch_bounds__(FCP_TPB) k_grad_lsq_qr(MeshView m, const double *__restrict__ D, const double *__restrict__ phi,
                                                          double *__restrict__ g) {
  FCP_CELL_LOOP(c, m.n) {
    const double pc = phi[c];
    const int64_t n = m.n;
    double g1 = 0.0, g2 = 0.0, g3 = 0.0;
    FCP_FACE_LOOP(m, c) {
      const int32_t o = __ldcs(m.other + fbase__ + (int64_t)q__ * 32);
      const double b = phi[o] - pc;                        // :1106-1126 (inner: phi(other)-phi(cell); boundary: phi(ijb)-phi(cell))
      g1 = g1 + D[(int64_t)(q__ * 3 + 0) * n + c] * b;     // :1139-1141 sum(D(i,1:l)*b(1:l))
      g2 = g2 + D[(int64_t)(q__ * 3 + 1) * n + c] * b;
      g3 = g3 + D[(int64_t)(q__ * 3 + 2) * n + c] * b;
    }
    g[3 * (int64_t)c] = g1;
    g[3 * (int64_t)c + 1] = g2;
    g[3 * (int64_t)c + 2] = g3;
  }
}

int fvm_lsq_qr_matrix(fcp_ctx *ctx, double *D) {
  if (ctx->n == 0) return FCP_OK;
  k_lsq_qr_matrix<<<FCP_GRID(ctx->n)>>>(fcp_mesh_view(ctx), D);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_grad_lsq_qr(fcp_ctx *ctx, const double *D, const double *phi, double *g) {
  if (ctx->B) FCP_CUDA(cudaMemsetAsync(g + 3 * (size_t)ctx->n, 0, sizeof(double) * 3 * (size_t)ctx->B, ctx->stream));
  if (ctx->n == 0) return FCP_OK;
  FCP_PROF(&ctx->prof, FCP_K_GRAD, ctx->stream, (k_grad_lsq_qr<<<FCP_GRID(ctx->n)>>>(fcp_mesh_view(ctx), D, phi, g)));
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

// ---------------------------------------------------------------------------------------------
// calcp_piso pieces
// ---------------------------------------------------------------------------------------------
// H(U) = rU - sum_faces h(k) U_other  (calcp_piso.f90:102-124), per cell in ascending face order
__global__ void __launch_bounds__(FCP_TPB) k_piso_H(MeshView m, const double *__restrict__ h, const double *__restrict__ rU,
                                                     const double *__restrict__ rV, const double *__restrict__ rW,
                                                     const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ w,
                                                     double *__restrict__ su, double *__restrict__ sv, double *__restrict__ sw) {
  FCP_CELL_LOOP(c, m.n) {
    double s1 = rU[c], s2 = rV[c], s3 = rW[c];
    FCP_FACE_LOOP(m, c) {
      FCP_FACE_FETCH(m);
      if (sl >= 0) {
        const double hk = h[sl];
        s1 = s1 - hk * u[o];
        s2 = s2 - hk * v[o];
        s3 = s3 - hk * w[o];
      }
    }
    su[c] = s1; sv[c] = s2; sw[c] = s3;
  }
}
// HbyA (:127-129)
__global__ void __launch_bounds__(FCP_TPB) k_piso_hbya(int32_t n, const double *__restrict__ apu, const double *__restrict__ apv,
                                                        const double *__restrict__ apw, const double *__restrict__ su,
                                                        const double *__restrict__ sv, const double *__restrict__ sw,
                                                        double *__restrict__ u, double *__restrict__ v, double *__restrict__ w) {
  FCP_CELL_LOOP(c, n) {
    u[c] = apu[c] * su[c];
    v[c] = apv[c] * sv[c];
    w[c] = apw[c] * sw[c];
  }
}
// sum(pp(1:numCells)) with the fixed tree (reduce.cuh); total -> out[0]
__global__ void __launch_bounds__(FCP_TPB) k_sum(int32_t n, const double *__restrict__ x, double *partials, int stride, unsigned int *counter,
                                                  double *out) {
  double s[1] = {0.0};
  FCP_CELL_LOOP(c, n) { s[0] = s[0] + x[c]; }
  double total[1];
  if (fcp_grid_reduce<1>(s, partials, stride, counter, total))
    if (threadIdx.x == 0) out[0] = total[0];
}
// p = (1-urfP) p + urfP (pp - pavg), pavg = sum/dble(numCells)   (:330-333)
__global__ void __launch_bounds__(FCP_TPB) k_piso_pupdate(int32_t n, double ncells_global, double urfp, const double *__restrict__ sum,
                                                           const double *__restrict__ pp, double *__restrict__ p) {
  const double pavg = sum[0] / ncells_global;
  FCP_CELL_LOOP(c, n) { p[c] = (1.0 - urfp) * p[c] + urfp * (pp[c] - pavg); }
}
// fluxmc faceflux_mass.f90:564-647, accumulated into su (:349-364)
__global__ void __launch_bounds__(FCP_TPB) k_piso_fluxmc(MeshView m, const double *__restrict__ den, const double *__restrict__ apu,
                                                          const double *__restrict__ dPdxi, double *__restrict__ su) {
  FCP_CELL_LOOP(c, m.n) {
    const double xc = m.xc[c], yc = m.yc[c], zc = m.zc[c];
    const double kc = apu[c] * den[c] * m.vol[c];
    const double gcx = dPdxi[3 * (int64_t)c], gcy = dPdxi[3 * (int64_t)c + 1], gcz = dPdxi[3 * (int64_t)c + 2];
    double s = su[c];
    FCP_FACE_LOOP(m, c) {
      FCP_FACE_FETCH(m);
      if (sl < 0) continue;
      const bool own = e > 0;
      const double arx = m.arx[f], ary = m.ary[f], arz = m.arz[f], xf = m.xf[f], yf = m.yf[f], zf = m.zf[f];
      const double xo = m.xc[o], yo = m.yc[o], zo = m.zc[o];
      const double ko = apu[o] * den[o] * m.vol[o];
      const double gox = dPdxi[3 * (int64_t)o], goy = dPdxi[3 * (int64_t)o + 1], goz = dPdxi[3 * (int64_t)o + 2];
      const double fxn = m.facint[f], fxp = 1.0 - fxn;
      const double xP = own ? xc : xo, yP = own ? yc : yo, zP = own ? zc : zo;
      const double xN = own ? xo : xc, yN = own ? yo : yc, zN = own ? zo : zc;
      const double kP = own ? kc : ko, kN = own ? ko : kc;
      const double gPx = own ? gcx : gox, gPy = own ? gcy : goy, gPz = own ? gcz : goz;
      const double gNx = own ? gox : gcx, gNy = own ? goy : gcy, gNz = own ? goz : gcz;
      const double xpn = xN - xP, ypn = yN - yP, zpn = zN - zP;
      const double are = sqrt(arx * arx + ary * ary + arz * arz);
      const double nxx = arx / are, nyy = ary / are, nzz = arz / are;
      double xpp = xf - (xf - xP) * nxx, ypp = yf - (yf - yP) * nyy, zpp = zf - (zf - zP) * nzz;
      double xep = xf - (xf - xN) * nxx, yep = yf - (yf - yN) * nyy, zep = zf - (zf - zN) * nzz;
      xpp = xpp - xP; ypp = ypp - yP; zpp = zpp - zP;
      xep = xep - xN; yep = yep - yN; zep = zep - zN;
      const double rapr = -((kP * fxp + kN * fxn) * are / (xpn * nxx + ypn * nyy + zpn * nzz));
      const double fmcor = rapr * ((gNx * xep - gPx * xpp) + (gNy * yep - gPy * ypp) + (gNz * zep - gPz * zpp));
      if (own) s = s - fmcor; else s = s + fmcor;
    }
    su[c] = s;
  }
}

int fvm_piso_hbya(fcp_ctx *ctx, const double *h, const double *rU, const double *rV, const double *rW, const double *apu, const double *apv,
                  const double *apw, double *u, double *v, double *w, double *su, double *sv, double *sw) {
  if (ctx->n == 0) return FCP_OK;
  FCP_PROF(&ctx->prof, FCP_K_PISO_H, ctx->stream, (k_piso_H<<<FCP_GRID(ctx->n)>>>(fcp_mesh_view(ctx), h, rU, rV, rW, u, v, w, su, sv, sw)));
  k_piso_hbya<<<FCP_GRID(ctx->n)>>>(ctx->n, apu, apv, apw, su, sv, sw, u, v, w);
  FCP_LAUNCHED(); FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_sum(fcp_ctx *ctx, const double *x, double *d_out) {
  FCP_TRY(krylov_ws_alloc(ctx->ws, ctx->pat.n, ctx->pat.ncols));
  k_sum<<<std::max(fcp_nchunks(ctx->n), 1), FCP_TPB, 0, ctx->stream>>>(ctx->n, x, ctx->ws.partials, ctx->ws.maxchunks, ctx->ws.counter, d_out);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
// ---- constant_mass_flow_forcing, src/cappuccino/constant_mass_flow_forcing.f90 (+ volumeWeightedAverage, fieldManipulation.f90:41-70)
// sums: vol*U, vol, vol*APU with the fixed tree; out[0..2]
__global__ void __launch_bounds__(FCP_TPB) k_cmf_sums(int32_t n, const double *__restrict__ vol, const double *__restrict__ u, const double *__restrict__ apu,
                                                       double *partials, int stride, unsigned int *counter, double *out) {
  double s[3] = {0.0, 0.0, 0.0};
  FCP_CELL_LOOP(c, n) {
    const double vc = vol[c];
    s[0] = s[0] + (vc * u[c]);
    s[1] = s[1] + vc;
    s[2] = s[2] + (vc * apu[c]);
  }
  double total[3];
  if (fcp_grid_reduce<3>(s, partials, stride, counter, total))
    if (threadIdx.x == 0) { out[0] = total[0]; out[1] = total[1]; out[2] = total[2]; }
}
// magUbarStar = sum(vol U)/sum(vol); rUAw = sum(vol APU)/sum(vol); gragPplus = (magUbar - magUbarStar)/rUAw; U += APU*gragPplus.
// Every thread recomputes the three scalars from the sums (same operations, same bits); thread 0 of CTA 0 publishes them in out[3], out[4].
__global__ void __launch_bounds__(FCP_TPB) k_cmf_apply(int32_t n, double magUbar, double *sums, const double *__restrict__ apu, double *__restrict__ u) {
  const double ustar = sums[0] / sums[1];
  const double ruaw = sums[2] / sums[1];
  const double gplus = (magUbar - ustar) / ruaw;
  FCP_CELL_LOOP(c, n) { u[c] = u[c] + apu[c] * gplus; }
  if (blockIdx.x == 0 && threadIdx.x == 0) { sums[3] = ustar; sums[4] = gplus; }
}
int fvm_cmf_forcing(fcp_ctx *ctx, double magUbar, const double *apu, double *u, double *d_sums /* [8] */) {
  FCP_TRY(krylov_ws_alloc(ctx->ws, ctx->pat.n, ctx->pat.ncols));
  k_cmf_sums<<<std::max(fcp_nchunks(ctx->n), 1), FCP_TPB, 0, ctx->stream>>>(ctx->n, ctx->vol, u, apu, ctx->ws.partials, ctx->ws.maxchunks, ctx->ws.counter, d_sums);
  FCP_LAUNCHED();
  if (ctx->comm) FCP_TRY(comm_allgather_sum(ctx->comm, d_sums, 3, ctx->stream));
  k_cmf_apply<<<std::max(fcp_nchunks(ctx->n), 1), FCP_TPB, 0, ctx->stream>>>(ctx->n, magUbar, d_sums, apu, u);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
// ---- updateBoundary(phi), boundary/updateBoundary.f90 (boundary-face parallel; a periodic face and its twin get the same mean)
__global__ void __launch_bounds__(FCP_TPB) k_update_boundary(MeshView m, const int32_t *__restrict__ bftype, double *phi) {
  FCP_CELL_LOOP(i, m.B) {
    const int t = bftype[i];
    const int32_t ijp = m.owner[m.F + i], ijb = m.n + i;
    // the reference walks the patches in order: a twin ('empty') patch listed BEFORE its periodic patch first copies the owner value and
    // is then overwritten with the pair's mean (:62-66); listed AFTER it, its own copy is the last write and stays
    if (t == FCP_BC_PERIODIC) {
      const double v = 0.5 * (phi[ijp] + phi[m.per_cell[i]]);
      phi[ijb] = v;
      if (m.per_face[i] < m.F + i) phi[m.n + (m.per_face[i] - m.F)] = v;
    } else if (t == FCP_BC_EMPTY && m.per_cell && m.per_cell[i] >= 0) {
      if (m.per_face[i] < m.F + i) phi[ijb] = phi[ijp];
    } else if (t == FCP_BC_OUTLET || t == FCP_BC_SYMMETRY || t == FCP_BC_PRESSURE || t == FCP_BC_EMPTY) {
      phi[ijb] = phi[ijp];
    }
  }
}
int fvm_update_boundary(fcp_ctx *ctx, double *phi) {
  if (ctx->B == 0) return FCP_OK;
  k_update_boundary<<<FCP_GRID(ctx->B)>>>(fcp_mesh_view(ctx), ctx->bftype, phi);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
// ---- wall distance, src/mesh/wall_distance.f90:75-133 (the pieces that are not laplacian / csrsolve / grad_gauss)
__global__ void __launch_bounds__(FCP_TPB) k_neg_vol(int32_t n, const double *__restrict__ vol, double *__restrict__ q) {
  FCP_CELL_LOOP(c, n) { q[c] = -vol[c]; }                                  // :83  q = -Vol(1:numCells)
}
__global__ void __launch_bounds__(FCP_TPB) k_copy_owner_nonwall(MeshView m, const int32_t *__restrict__ bftype, double *phi) {
  FCP_CELL_LOOP(i, m.B) {                                                   // :107-118  every patch but 'wall' takes the owner value
    if (bftype[i] != FCP_BC_WALL) phi[m.n + i] = phi[m.owner[m.F + i]];
  }
}
__global__ void __launch_bounds__(FCP_TPB) k_wall_distance(int32_t n, const double *__restrict__ g, const double *__restrict__ phi, double *__restrict__ wd) {
  FCP_CELL_LOOP(c, n) {                                                     // :124-125
    const double gx = g[3 * (int64_t)c], gy = g[3 * (int64_t)c + 1], gz = g[3 * (int64_t)c + 2];
    wd[c] = -sqrt(gx * gx + gy * gy + gz * gz) + sqrt(gx * gx + gy * gy + gz * gz + 2 * phi[c]);
  }
}
int fvm_neg_vol(fcp_ctx *ctx, double *q) {
  if (ctx->n == 0) return FCP_OK;
  k_neg_vol<<<FCP_GRID(ctx->n)>>>(ctx->n, ctx->vol, q);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_wall_distance_finish(fcp_ctx *ctx, int stage, double *phi, const double *g, double *wd) {
  if (stage == 0) {
    if (ctx->B == 0) return FCP_OK;
    k_copy_owner_nonwall<<<FCP_GRID(ctx->B)>>>(fcp_mesh_view(ctx), ctx->bftype, phi);
  } else {
    if (ctx->n == 0) return FCP_OK;
    k_wall_distance<<<FCP_GRID(ctx->n)>>>(ctx->n, g, phi, wd);
  }
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_piso_pupdate(fcp_ctx *ctx, double ncells_global, double urfp, const double *d_sum, const double *pp, double *p) {
  if (ctx->n == 0) return FCP_OK;
  k_piso_pupdate<<<FCP_GRID(ctx->n)>>>(ctx->n, ncells_global, urfp, d_sum, pp, p);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int fvm_piso_fluxmc(fcp_ctx *ctx, const double *den, const double *apu, const double *dPdxi, double *su) {
  if (ctx->n == 0) return FCP_OK;
  k_piso_fluxmc<<<FCP_GRID(ctx->n)>>>(fcp_mesh_view(ctx), den, apu, dPdxi, su);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
