// pattern.cu -- sparsity-pattern layer: CSR (host-visible, the reference's ia/ja/diag) <-> SELL-32 (device),
// level schedules for the IC(0)/ILU(0) sweeps.  Reference: src/sparseMatrix/sparse_matrix.f90:86-296.
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <thread>
#include "fcp_internal.h"

static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_fcp_launches{0};
void fcp_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char *fcp_last_error(void) { return g_err; }
extern "C" int fcp_version(void) { return 100; }
extern "C" int64_t fcp_launch_count(void) { return g_fcp_launches.load(); }

template <class T> int dev_alloc(T **dptr, size_t count) {
  *dptr = nullptr;
  if (count == 0) count = 1;
  FCP_CUDA(cudaMalloc((void **)dptr, count * sizeof(T)));
  return FCP_OK;
}
template <class T> int dev_upload(T **dptr, const T *h, size_t count) {
  FCP_TRY(dev_alloc(dptr, count));
  if (count) {
    // a pageable-source cudaMemcpy may return while the DMA is still in flight and the library's streams are cudaStreamNonBlocking
    // (no implicit ordering with the legacy stream): wait for it here, uploads happen at set-up time only
    FCP_CUDA(cudaMemcpy(*dptr, h, count * sizeof(T), cudaMemcpyHostToDevice));
    FCP_CUDA(cudaStreamSynchronize(0));
  }
  return FCP_OK;
}
template int dev_alloc<double>(double **, size_t);
template int dev_alloc<int32_t>(int32_t **, size_t);
template int dev_alloc<int64_t>(int64_t **, size_t);
template int dev_alloc<unsigned int>(unsigned int **, size_t);
template int dev_alloc<unsigned long long>(unsigned long long **, size_t);
template int dev_upload<double>(double **, const double *, size_t);
template int dev_upload<int32_t>(int32_t **, const int32_t *, size_t);
template int dev_upload<int64_t>(int64_t **, const int64_t *, size_t);
template int dev_upload<unsigned long long>(unsigned long long **, const unsigned long long *, size_t);
template int dev_upload<unsigned long long *>(unsigned long long ***, unsigned long long *const *, size_t);

int sell_from_csr(SellPattern &p, int32_t n, int32_t ncols, const int32_t *ia1, const int32_t *ja1, const int32_t *diag1,
                  const std::vector<std::vector<int32_t>> *halo) {
  p.n = n;
  p.ncols = ncols;
  p.nnz = (int64_t)ia1[n] - 1;
  p.h_ia.assign(ia1, ia1 + n + 1);
  p.h_ja.assign(ja1, ja1 + p.nnz);
  p.h_diag.assign(diag1, diag1 + n);
  p.nslices = (n + 31) / 32;
  std::vector<int64_t> slptr(p.nslices + 1, 0);
  std::vector<int32_t> rinfo(n), llen(n), ia0(n + 1);
  bool any_halo = false;
  int64_t next = 0;
  for (int32_t r = 0; r < n; ++r) {
    int32_t ll = ia1[r + 1] - ia1[r];
    int32_t len = ll + (halo ? (int32_t)(*halo)[r].size() : 0);
    int32_t dpos = diag1[r] - ia1[r];
    if (len >= 65536 || dpos < 0 || dpos >= ll) {
      fcp_set_error("sell_from_csr: row %d has len %d, diagonal offset %d (unsupported)", r + 1, len, dpos);
      return FCP_EINVAL;
    }
    any_halo |= (len != ll);
    rinfo[r] = len | (dpos << 16);
    llen[r] = ll;
    ia0[r] = ia1[r] - 1;
    next += len;
  }
  ia0[n] = ia1[n] - 1;
  p.nnz_ext = next;
  for (int32_t s = 0; s < p.nslices; ++s) {
    int32_t w = 0;
    for (int32_t r = s * 32; r < std::min(n, s * 32 + 32); ++r) w = std::max(w, rinfo[r] & 0xffff);
    slptr[s + 1] = slptr[s] + (int64_t)w * 32;
  }
  p.nnzp = slptr[p.nslices];
  p.tile_cap = 0;
  for (int32_t s = 0; s < p.nslices; s += 8) p.tile_cap = std::max<int64_t>(p.tile_cap, slptr[std::min(s + 8, p.nslices)] - slptr[s]);
  if (p.nnzp >= (int64_t)2147483647) {
    fcp_set_error("sell_from_csr: padded nnz %lld exceeds int32 positions", (long long)p.nnzp);
    return FCP_EINVAL;
  }
  std::vector<int32_t> ja(p.nnzp);
  for (int32_t s = 0; s < p.nslices; ++s) {
    int32_t w = (int32_t)((slptr[s + 1] - slptr[s]) / 32);
    for (int32_t l = 0; l < 32; ++l) {
      int32_t r = s * 32 + l;
      int64_t base = slptr[s] + l;
      if (r >= n) {
        for (int32_t j = 0; j < w; ++j) ja[base + (int64_t)j * 32] = 0;
        continue;
      }
      int32_t ll = llen[r], len = rinfo[r] & 0xffff;
      for (int32_t j = 0; j < ll; ++j) ja[base + (int64_t)j * 32] = ja1[ia1[r] - 1 + j] - 1;
      for (int32_t j = ll; j < len; ++j) ja[base + (int64_t)j * 32] = (*halo)[r][j - ll];
      for (int32_t j = len; j < w; ++j) ja[base + (int64_t)j * 32] = r;
    }
  }
  FCP_TRY(dev_upload(&p.slptr, slptr.data(), slptr.size()));
  FCP_TRY(dev_upload(&p.rinfo, rinfo.data(), rinfo.size()));
  FCP_TRY(dev_upload(&p.ja, ja.data(), ja.size()));
  FCP_TRY(dev_upload(&p.ia0, ia0.data(), ia0.size()));
  if (any_halo) FCP_TRY(dev_upload(&p.llen, llen.data(), llen.size()));
  return FCP_OK;
}

void sell_free(SellPattern &p) {
  cudaFree(p.slptr); cudaFree(p.rinfo); cudaFree(p.ja); cudaFree(p.ia0); cudaFree(p.llen);
  cudaFree(p.lev_ptr); cudaFree(p.lev_rows); cudaFree(p.blev_ptr); cudaFree(p.blev_rows); cudaFree(p.tpos);
  cudaFree(p.plev_rows); cudaFree(p.pblev_rows); cudaFree(p.ready);
  for (TriTiles &t : p.tri) { cudaFree(t.tptr); cudaFree(t.tcol); cudaFree(t.tsrc); cudaFree(t.ttsrc); cudaFree(t.tval); cudaFree(t.ttval); cudaFree(t.dtile); }
  for (auto *q : p.zll) cudaFree(q);
  p = SellPattern();
}

// ---- value conversion: host-visible CSR a(nnz)  <->  SELL values -----------------------------------------------
__global__ void k_csr_to_sell(int32_t n, const int64_t *__restrict__ slptr, const int32_t *__restrict__ ia0,
                              const double *__restrict__ a_csr, double *__restrict__ a_sell) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int64_t base = slptr[r >> 5] + (r & 31);
  int32_t k0 = ia0[r], k1 = ia0[r + 1];
  for (int32_t k = k0; k < k1; ++k) a_sell[base + (int64_t)(k - k0) * 32] = a_csr[k];
}
__global__ void k_sell_to_csr(int32_t n, const int64_t *__restrict__ slptr, const int32_t *__restrict__ ia0,
                              const double *__restrict__ a_sell, double *__restrict__ a_csr) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int64_t base = slptr[r >> 5] + (r & 31);
  int32_t k0 = ia0[r], k1 = ia0[r + 1];
  for (int32_t k = k0; k < k1; ++k) a_csr[k] = a_sell[base + (int64_t)(k - k0) * 32];
}
int sell_values_from_csr(const SellPattern &p, const double *d_a_csr, double *d_a_sell, cudaStream_t st) {
  if (p.n == 0) return FCP_OK;
  k_csr_to_sell<<<(p.n + 255) / 256, 256, 0, st>>>(p.n, p.slptr, p.ia0, d_a_csr, d_a_sell);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int sell_values_to_csr(const SellPattern &p, const double *d_a_sell, double *d_a_csr, cudaStream_t st) {
  if (p.n == 0) return FCP_OK;
  k_sell_to_csr<<<(p.n + 255) / 256, 256, 0, st>>>(p.n, p.slptr, p.ia0, d_a_sell, d_a_csr);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

// ---- level schedules -------------------------------------------------------------------------------------------
// forward sweep: row i depends on the rows of its entries before the diagonal position; backward sweep: on those
// after it (local columns only: block-Jacobi across ranks, src-par/iccg.f90:93-127).  Rows of one level are listed in
// ascending row order; executing level after level is arithmetically identical to the sequential sweep
// (linear_solvers.f90:439-445, 458-475) because every row still sums its own entries in CSR order.
int sell_build_levels(SellPattern &p, cudaStream_t st) {
  (void)st;
  if (p.levels_built) return FCP_OK;
  const int32_t n = p.n;
  const int32_t *ia = p.h_ia.data(), *ja = p.h_ja.data(), *dg = p.h_diag.data();
  std::vector<int32_t> lev(n, 0), rows(n);
  auto finish = [&](std::vector<int32_t> &ptr, int32_t nlev) {
    ptr.assign(nlev + 1, 0);
    for (int32_t i = 0; i < n; ++i) ptr[lev[i] + 1]++;
    for (int32_t l = 0; l < nlev; ++l) ptr[l + 1] += ptr[l];
    std::vector<int32_t> pos(ptr.begin(), ptr.end() - 1);
    for (int32_t i = 0; i < n; ++i) rows[pos[lev[i]]++] = i;
  };
  int32_t nlev = 0;
  for (int32_t i = 0; i < n; ++i) {
    int32_t l = 0;
    for (int32_t k = ia[i]; k < dg[i]; ++k) l = std::max(l, lev[ja[k - 1] - 1] + 1);
    lev[i] = l;
    nlev = std::max(nlev, l + 1);
  }
  if (n == 0) nlev = 0;
  finish(p.h_lev_ptr, nlev);
  p.nlevels = nlev;
  FCP_TRY(dev_upload(&p.lev_ptr, p.h_lev_ptr.data(), p.h_lev_ptr.size()));
  FCP_TRY(dev_upload(&p.lev_rows, rows.data(), rows.size()));
  auto padded = [&](const std::vector<int32_t> &ptr, int32_t nl) {      // every level starts on a warp boundary: no lane ever waits on its own warp
    std::vector<int32_t> out;
    out.reserve((size_t)n + 32 * (size_t)nl);
    for (int32_t l = 0; l < nl; ++l) {
      for (int32_t q = ptr[l]; q < ptr[l + 1]; ++q) out.push_back(rows[q]);
      while (out.size() % 32) out.push_back(-1);
    }
    return out;
  };
  {
    std::vector<int32_t> pl = padded(p.h_lev_ptr, nlev);
    p.nplev = (int32_t)pl.size();
    FCP_TRY(dev_upload(&p.plev_rows, pl.data(), pl.size()));
  }
  nlev = 0;
  for (int32_t i = n - 1; i >= 0; --i) {
    int32_t l = 0;
    for (int32_t k = dg[i] + 1; k <= ia[i + 1] - 1; ++k) l = std::max(l, lev[ja[k - 1] - 1] + 1);   // all have column > i? see below
    lev[i] = l;
    nlev = std::max(nlev, l + 1);
  }
  if (n == 0) nlev = 0;
  finish(p.h_blev_ptr, nlev);
  p.nblevels = nlev;
  FCP_TRY(dev_upload(&p.blev_ptr, p.h_blev_ptr.data(), p.h_blev_ptr.size()));
  FCP_TRY(dev_upload(&p.blev_rows, rows.data(), rows.size()));
  {
    std::vector<int32_t> pl = padded(p.h_blev_ptr, nlev);
    p.npblev = (int32_t)pl.size();
    FCP_TRY(dev_upload(&p.pblev_rows, pl.data(), pl.size()));
    std::vector<int32_t> zero((size_t)std::max(n, 1), 0);
    FCP_TRY(dev_upload(&p.ready, zero.data(), zero.size()));
    p.sweep_epoch = 0;
  }
  p.levels_built = true;
  return FCP_OK;
}

// The two triangles in level-tile order for the flag-in-data sweeps (fcp_internal.h: TriTiles).  with_transposed: also the positions of the
// transposed entries of the lower triangle (the ILU(0) factor of bicgstab, needs sell_build_tpos).
int sell_build_tiles(SellPattern &p, bool with_transposed, cudaStream_t st) {
  FCP_TRY(sell_build_levels(p, st));
  if (with_transposed) FCP_TRY(sell_build_tpos(p, st));
  const bool have = p.tri[0].tptr != nullptr;
  if (have && (!with_transposed || p.tri[0].ttsrc)) return FCP_OK;
  const int32_t n = p.n;
  const int32_t *ia = p.h_ia.data(), *ja = p.h_ja.data(), *dg = p.h_diag.data();
  std::vector<int64_t> slptr(p.nslices + 1);
  FCP_CUDA(cudaMemcpy(slptr.data(), p.slptr, slptr.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
  std::vector<int32_t> tpos;
  if (with_transposed) {
    tpos.resize(p.nnzp);
    FCP_CUDA(cudaMemcpy(tpos.data(), p.tpos, tpos.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  for (int dir = 0; dir < 2; ++dir) {
    TriTiles &t = p.tri[dir];
    const int32_t np = dir == 0 ? p.nplev : p.npblev;
    std::vector<int32_t> prow((size_t)std::max(np, 1), -1);
    if (np) FCP_CUDA(cudaMemcpy(prow.data(), dir == 0 ? p.plev_rows : p.pblev_rows, sizeof(int32_t) * (size_t)np, cudaMemcpyDeviceToHost));
    const int32_t ntiles = np / 32;
    std::vector<int64_t> tptr((size_t)ntiles + 1, 0);
    int32_t maxlen = 0;
    auto k0 = [&](int32_t i) { return dir == 0 ? ia[i] : dg[i] + 1; };           // 1-based CSR range of the triangle in row i
    auto k1 = [&](int32_t i) { return dir == 0 ? dg[i] : ia[i + 1]; };           // (exclusive)
    for (int32_t tt = 0; tt < ntiles; ++tt) {
      int32_t len = 0;
      for (int l = 0; l < 32; ++l) {
        const int32_t i = prow[(size_t)tt * 32 + l];
        if (i >= 0) len = std::max(len, k1(i) - k0(i));
      }
      maxlen = std::max(maxlen, len);
      tptr[tt + 1] = tptr[tt] + 32 * (int64_t)len;
    }
    const int64_t nent = tptr[ntiles];
    if (!have) {
      std::vector<int32_t> tcol((size_t)std::max<int64_t>(nent, 1), -1), tsrc((size_t)std::max<int64_t>(nent, 1), -1);
      for (int32_t tt = 0; tt < ntiles; ++tt)
        for (int l = 0; l < 32; ++l) {
          const int32_t i = prow[(size_t)tt * 32 + l];
          if (i < 0) continue;
          const int64_t sbase = slptr[i >> 5] + (i & 31);
          for (int32_t k = k0(i); k < k1(i); ++k) {
            const int64_t q = tptr[tt] + 32 * (int64_t)(k - k0(i)) + l;
            tcol[q] = ja[k - 1] - 1;
            tsrc[q] = (int32_t)(sbase + (int64_t)(k - ia[i]) * 32);
          }
        }
      t.ntiles = ntiles; t.maxlen = maxlen; t.nent = nent;
      t.prow = dir == 0 ? p.plev_rows : p.pblev_rows;
      FCP_TRY(dev_upload(&t.tptr, tptr.data(), tptr.size()));
      FCP_TRY(dev_upload(&t.tcol, tcol.data(), tcol.size()));
      FCP_TRY(dev_upload(&t.tsrc, tsrc.data(), tsrc.size()));
      FCP_TRY(dev_alloc(&t.tval, (size_t)std::max<int64_t>(nent, 1)));
      FCP_TRY(dev_alloc(&t.dtile, (size_t)std::max(np, 1)));
    }
    if (dir == 0 && with_transposed && !t.ttsrc) {
      std::vector<int32_t> ttsrc((size_t)std::max<int64_t>(nent, 1), -1);
      for (int32_t tt = 0; tt < ntiles; ++tt)
        for (int l = 0; l < 32; ++l) {
          const int32_t i = prow[(size_t)tt * 32 + l];
          if (i < 0) continue;
          const int64_t sbase = slptr[i >> 5] + (i & 31);
          for (int32_t k = ia[i]; k < dg[i]; ++k) ttsrc[tptr[tt] + 32 * (int64_t)(k - ia[i]) + l] = tpos[sbase + (int64_t)(k - ia[i]) * 32];
        }
      FCP_TRY(dev_upload(&t.ttsrc, ttsrc.data(), ttsrc.size()));
      FCP_TRY(dev_alloc(&t.ttval, (size_t)std::max<int64_t>(nent, 1)));
    }
  }
  if (!have) {
    for (auto *&q : p.zll) {
      FCP_TRY(dev_alloc(&q, (size_t)2 * std::max(n, 1)));
      FCP_CUDA(cudaMemset(q, 0, sizeof(unsigned long long) * 2 * (size_t)std::max(n, 1)));
    }
    FCP_CUDA(cudaStreamSynchronize(0));
    p.ll_epoch = 0;
  }
  return FCP_OK;
}

// position (SELL) of the transposed entry a(j,i) for every entry a(i,j) before the diagonal: bicgstab's ILU(0)
// diagonal (linear_solvers.f90:613-624) searches row j upward from diag(j) for column i; when the pattern has no such
// entry the Fortran DO loop leaves l = ia(j+1), i.e. the first entry of the next row -- reproduced here.
int sell_build_tpos(SellPattern &p, cudaStream_t st) {
  (void)st;
  if (p.tpos) return FCP_OK;
  const int32_t n = p.n;
  const int32_t *ia = p.h_ia.data(), *ja = p.h_ja.data(), *dg = p.h_diag.data();
  std::vector<int64_t> slptr(p.nslices + 1);
  FCP_CUDA(cudaMemcpy(slptr.data(), p.slptr, slptr.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
  auto sellpos = [&](int32_t k1 /*1-based CSR position*/) -> int64_t {
    if (k1 > p.nnz) return -1;   // past the end of the matrix: the reference would read out of bounds
    int32_t r = (int32_t)(std::upper_bound(ia, ia + n + 1, k1) - ia) - 1;
    return slptr[r >> 5] + (int64_t)(k1 - ia[r]) * 32 + (r & 31);
  };
  std::vector<int32_t> tpos(p.nnzp, -1);
  for (int32_t i = 0; i < n; ++i) {
    for (int32_t k = ia[i]; k < dg[i]; ++k) {
      int32_t j = ja[k - 1];
      int32_t l;
      for (l = dg[j - 1]; l <= ia[j] - 1; ++l)
        if (ja[l - 1] == i + 1) break;
      int64_t pos = slptr[i >> 5] + (int64_t)(k - ia[i]) * 32 + (i & 31);
      tpos[pos] = (int32_t)sellpos(l);
    }
  }
  FCP_TRY(dev_upload(&p.tpos, tpos.data(), tpos.size()));
  return FCP_OK;
}
