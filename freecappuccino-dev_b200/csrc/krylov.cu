// krylov.cu -- SELL-32 SpMV and the three Krylov solvers of the reference
//   dpcg      src/linearSolvers/linear_solvers.f90:206-359   (src-par/dpcg.f90 for the halo term)
//   iccg      :364-545        IC(0), diagonal-only factor (:439-445), apply :458-475
//   bicgstab  :548-786        ILU(0), diagonal-only factor (:613-624)
// Arithmetic order per row and per reduction is fixed (fcp_internal.h), no FMA contraction (-fmad=false), so the
// iterates are bitwise those of the CPU restatement in its TREE summation mode.
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "fcp_internal.h"
#include "reduce.cuh"
#include "p2p.cuh"
namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
int krylov_ws_alloc(KrylovWS &ws, int32_t n, int32_t ncols) {
  if (ws.n == n && ws.ncols == ncols && ws.res) return FCP_OK;
  krylov_ws_free(ws);
  ws.n = n;
  ws.ncols = ncols;
  FCP_TRY(dev_alloc(&ws.res, (size_t)n));
  FCP_TRY(dev_alloc(&ws.pk, (size_t)ncols));
  FCP_TRY(dev_alloc(&ws.zk, (size_t)ncols));
  FCP_TRY(dev_alloc(&ws.adiag, (size_t)n));
  ws.maxchunks = fcp_nchunks(n) + 1;
  FCP_TRY(dev_alloc(&ws.partials, (size_t)4 * ws.maxchunks));
  FCP_TRY(dev_alloc(&ws.counter, 1));
  FCP_CUDA(cudaMemset(ws.counter, 0, sizeof(unsigned int)));
  FCP_CUDA(cudaStreamSynchronize(0));   // cudaMemset is asynchronous; the solver streams do not order with the legacy stream
  FCP_CUDA(cudaMalloc((void **)&ws.sc, sizeof(KrylovScalars)));
  FCP_CUDA(cudaMallocHost((void **)&ws.h_sc, sizeof(KrylovScalars)));
  FCP_CUDA(cudaEventCreateWithFlags(&ws.ev[0], cudaEventDisableTiming));
  FCP_CUDA(cudaEventCreateWithFlags(&ws.ev[1], cudaEventDisableTiming));
  return FCP_OK;
}
static int krylov_ws_need(KrylovWS &ws, double **p, size_t count) {
  if (*p) return FCP_OK;
  FCP_TRY(dev_alloc(p, count));
  FCP_CUDA(cudaMemset(*p, 0, std::max<size_t>(count, 1) * sizeof(double)));
  FCP_CUDA(cudaStreamSynchronize(0));
  (void)ws;
  return FCP_OK;
}
void krylov_ws_free(KrylovWS &ws) {
  cudaFree(ws.res); cudaFree(ws.pk); cudaFree(ws.zk); cudaFree(ws.adiag); cudaFree(ws.d);
  cudaFree(ws.reso); cudaFree(ws.uk); cudaFree(ws.vk); cudaFree(ws.tmp);
  cudaFree(ws.partials); cudaFree(ws.counter); cudaFree(ws.sc);
  if (ws.h_sc) cudaFreeHost(ws.h_sc);
  if (ws.ev[0]) cudaEventDestroy(ws.ev[0]);
  if (ws.ev[1]) cudaEventDestroy(ws.ev[1]);
  ws = KrylovWS();
}

// ---------------------------------------------------------------------------------------------
// row kernels on SELL-32
// ---------------------------------------------------------------------------------------------
struct SellView {
  const int64_t *slptr;
  const int32_t *rinfo;
  const int32_t *ja;
  const double *a;
  const int32_t *llen;   // [n] number of local (non-ghost) entries per row; nullptr when the pattern has no halo columns
};

// s <- s (+|-) sum_k a(r,k) x(ja(r,k)), entries in CSR order (linear_solvers.f90:256-261 with SUB, :308-313 without)
template <bool SUB>
__device__ __forceinline__ double sell_row_sum(const SellView &m, const double *__restrict__ x, int32_t r, double s) {
  const int64_t base = __ldg(&m.slptr[r >> 5]) + (r & 31);
  const int32_t len = __ldg(&m.rinfo[r]) & 0xffff;
  const double *__restrict__ ap = m.a + base;
  const int32_t *__restrict__ jp = m.ja + base;
#pragma unroll 4
  for (int32_t k = 0; k < len; ++k) {
    const double av = __ldcs(ap + (int64_t)k * 32);
    const int32_t c = __ldcs(jp + (int64_t)k * 32);
    const double t = av * __ldg(x + c);
    s = SUB ? (s - t) : (s + t);
  }
  return s;
}

// Row sum for the rows of a chunk that owns process faces, on the peer-memory path: a ghost column (>= n) is not read
// from x but from this rank's LL slots, where the neighbour's k_cg_pk stored it during this very iteration; the word's
// embedded sequence number tells when it has arrived (src-par/dpcg.f90:118 `call exchange(pk)` + :129-143 halo term).
template <bool SUB>
__device__ __forceinline__ double sell_row_sum_halo(const SellView &m, const double *__restrict__ x, int32_t r, double s, const CommDev *cd,
                                                    unsigned int seq) {
  const int64_t base = __ldg(&m.slptr[r >> 5]) + (r & 31);
  const int32_t len = __ldg(&m.rinfo[r]) & 0xffff;
  const int32_t n = cd->n;
  for (int32_t k = 0; k < len; ++k) {
    const double av = __ldcs(m.a + base + (int64_t)k * 32);
    const int32_t c = __ldcs(m.ja + base + (int64_t)k * 32);
    const double xv = c >= n ? p2p_ll_load(cd->ll + 2 * (size_t)__ldg(cd->ghost_ord + (c - n)), seq, cd->hdr) : __ldg(x + c);
    const double t = av * xv;
    s = SUB ? (s - t) : (s + t);
  }
  return s;
}

#define FCP_ROW_LOOP(r, n)                                                                       \
  for (int j__ = 0; j__ < FCP_IPT; ++j__)                                                         \
    for (int32_t r = (int32_t)((int64_t)blockIdx.x * FCP_CHUNK + j__ * FCP_TPB + threadIdx.x), once__ = 1; \
         once__ && r < (n); once__ = 0)

__global__ void __launch_bounds__(FCP_TPB) k_spmv(int32_t n, SellView m, const double *__restrict__ x, double *__restrict__ y) {
  FCP_ROW_LOOP(r, n) { y[r] = sell_row_sum<false>(m, x, r, 0.0); }
}
int sell_spmv(const SellPattern &p, const double *a, const double *x, double *y, cudaStream_t st) {
  if (p.n == 0) return FCP_OK;
  SellView m{p.slptr, p.rinfo, p.ja, a, p.llen};
  k_spmv<<<fcp_nchunks(p.n), FCP_TPB, 0, st>>>(p.n, m, x, y);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

// ---------------------------------------------------------------------------------------------
// scalar epilogues.  Single GPU: run by thread 0 of the last CTA of the producing kernel.  Multi GPU: the raw local
// sums go through the rank-ordered cross-rank sum (comm.cu) and k_epilogue runs them (src-par/global_sum_mpi.f90).
// ---------------------------------------------------------------------------------------------
enum { EPI_NONE = 0, EPI_INIT_CG, EPI_PKAPK, EPI_CG_UPDATE, EPI_SK, EPI_INIT_BICG, EPI_UKRESO, EPI_VK, EPI_BICG_UPDATE, EPI_GS };

__device__ __forceinline__ void conv_check(KrylovScalars *sc) {
  // linear_solvers.f90:336-346
  if (sc->iters == 1) {
    sc->factor = sc->red[2] + FCP_SMALL;
    sc->resor = sc->res0 / sc->factor;
  }
  const double rsm = sc->resl / (sc->res0 + FCP_SMALL);
  if (rsm < sc->tol_rel || sc->resl < sc->tol_abs || sc->iters >= sc->itr_max) sc->done = 1;
}

__device__ void krylov_epilogue(int which, KrylovScalars *sc) {
  switch (which) {
    case EPI_INIT_CG:   // red0 = sum|res|, red1 = sum res*zk (dpcg only; iccg gets sk from EPI_SK)
      sc->res0 = sc->red[0];
      sc->resl = sc->red[0];
      sc->iters = 0;
      sc->factor = 0.0;
      sc->resor = sc->red[0];
      sc->done = (sc->red[0] < sc->tol_abs || sc->itr_max <= 0) ? 1 : 0;   // :266-270
      sc->s0 = (double)1.e20f;                                           // :276  s0=1.e20
      sc->sk = sc->red[1];
      sc->bet = sc->sk / sc->s0;
      break;
    case EPI_SK:        // iccg: sk = sum res*zk after the preconditioner apply
      sc->sk = sc->red[0];
      sc->bet = sc->sk / sc->s0;
      break;
    case EPI_PKAPK:
      sc->pkapk = sc->red[0];
      sc->alf = sc->sk / sc->pkapk;
      break;
    case EPI_CG_UPDATE:  // red0 = sum|res|, red1 = next sk (dpcg), red2 = sum|a_ii fi_i| (first iteration)
      sc->resl = sc->red[0];
      sc->s0 = sc->sk;
      sc->sk = sc->red[1];
      sc->bet = sc->sk / sc->s0;
      sc->iters += 1;
      conv_check(sc);
      break;
    case EPI_INIT_BICG:  // red0 = sum|res| ; red1 = sum res*reso (= sum res*res)
      sc->res0 = sc->red[0];
      sc->resl = sc->red[0];
      sc->iters = 0;
      sc->factor = 0.0;
      sc->resor = sc->red[0];
      sc->done = (sc->red[0] < sc->tol_abs || sc->itr_max <= 0) ? 1 : 0;
      sc->alf = 1.0; sc->beto = 1.0; sc->gam = 1.0;                        // :640-642
      sc->bet = sc->red[1];
      sc->om = sc->bet * sc->gam / (sc->alf * sc->beto + FCP_SMALL);       // :655
      sc->beto = sc->bet;
      break;
    case EPI_UKRESO:
      sc->ukreso = sc->red[0];
      sc->gam = sc->bet / sc->ukreso;                                      // :701
      break;
    case EPI_VK:
      sc->vkres = sc->red[0];
      sc->vkvk = sc->red[1];
      sc->alf = sc->vkres / (sc->vkvk + FCP_SMALL);                        // :747
      break;
    case EPI_BICG_UPDATE:  // red0 = sum|res|, red1 = next bet = sum res*reso, red2 = factor sum
      sc->resl = sc->red[0];
      sc->iters += 1;
      conv_check(sc);
      sc->bet = sc->red[1];
      sc->om = sc->bet * sc->gam / (sc->alf * sc->beto + FCP_SMALL);
      sc->beto = sc->bet;
      break;
    case EPI_GS:           // Gauss-Seidel, linear_solvers.f90:147-179: red0 = sum|res| of the sweep, red1 = sum|a_ii fi_i| (first sweep)
      if (sc->iters == 0) {
        sc->res0 = sc->red[0];
        if (sc->res0 < sc->tol_abs) {          // :151-157 -- AFTER the first sweep has updated fi
          sc->resl = sc->res0; sc->resor = sc->res0; sc->factor = 0.0; sc->iters = 1; sc->done = 1;
          break;
        }
      }
      sc->resl = sc->red[0];
      sc->iters += 1;
      if (sc->iters == 1) {
        sc->factor = sc->red[1] + FCP_SMALL;
        sc->resor = sc->res0 / sc->factor;
      }
      if (sc->resl / (sc->res0 + FCP_SMALL) < sc->tol_rel || sc->resl < sc->tol_abs || sc->iters >= sc->itr_max) sc->done = 1;
      break;
    default: break;
  }
}
// (after convergence the producing kernels return early and leave sc->red stale: the epilogue must not run again)
__global__ void k_epilogue(int which, KrylovScalars *sc) {
  if (sc->done && which != EPI_INIT_CG && which != EPI_INIT_BICG) return;
  krylov_epilogue(which, sc);
}

struct RedArgs {
  double *partials;
  int stride;
  unsigned int *counter;
  KrylovScalars *sc;
  int epi;    // epilogue id
  int fuse;   // 1: run the epilogue in the last CTA (single GPU); 0: only store the local sums in sc->red (NCCL path);
              // 2: peer-memory all-reduce inside the last CTA, then the epilogue
  const CommDev *cd;
  const int32_t *chunk_info;   // peer-memory path: [gridDim.x] chunk index | halo bit 31, in launch order (halo chunks first); one load per CTA
};

// Last CTA of a reducing kernel; `total` is valid in thread 0.  fuse == 2: thread (r, k) stores partial sum k into rank
// r's window as LL words tagged with the reduction's sequence number, then reads the partial of rank r from its own
// window (spinning until the tag matches), and thread 0 adds the P partials in RANK ORDER (src-par/global_sum_mpi.f90
// semantics; deterministic, identical bits on every rank).  Two slots by sequence parity: a rank can be at most one
// reduction ahead of any other.
template <int NS>
__device__ __forceinline__ void finish_epilogue(double (&total)[NS], const RedArgs &ra) {
  if (ra.fuse == 2) {
    const CommDev *cd = ra.cd;
    WinHeader *hdr = cd->hdr;
    __shared__ double tot_s[4];
    __shared__ double part_s[FCP_MAXR][4];
    __shared__ unsigned long long seq_s;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < NS; ++k) tot_s[k] = total[k];
      seq_s = ++hdr->red_seq;
    }
    __syncthreads();
    const unsigned int seq = (unsigned int)seq_s;
    const int par = (int)(seq & 1u);
    const int r = (int)threadIdx.x / NS, k = (int)threadIdx.x % NS;
    if (r < cd->nranks) {
      p2p_ll_store(&cd->peer_hdr[r]->rll[par][cd->rank][2 * k], tot_s[k], seq);
      part_s[r][k] = p2p_ll_load(&hdr->rll[par][r][2 * k], seq, hdr);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int kk = 0; kk < NS; ++kk) {
        double sum = part_s[0][kk];
        for (int rr = 1; rr < cd->nranks; ++rr) sum = sum + part_s[rr][kk];
        ra.sc->red[kk] = sum;
      }
      krylov_epilogue(ra.epi, ra.sc);
    }
    return;
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) ra.sc->red[k] = total[k];
    if (ra.fuse) krylov_epilogue(ra.epi, ra.sc);
  }
}
template <int NS>
__device__ __forceinline__ void finish_reduce_part(double (&s)[NS], const RedArgs &ra, int part, int nparts) {
  double total[NS];
  if (fcp_grid_reduce_part<NS>(s, ra.partials, ra.stride, ra.counter, part, nparts, total)) finish_epilogue<NS>(total, ra);
}
template <int NS>
__device__ __forceinline__ void finish_reduce(double (&s)[NS], const RedArgs &ra) {
  double total[NS];
  if (fcp_grid_reduce<NS>(s, ra.partials, ra.stride, ra.counter, total)) finish_epilogue<NS>(total, ra);
}

// ---------------------------------------------------------------------------------------------
// DPCG / shared CG kernels
// ---------------------------------------------------------------------------------------------
// res = rhs - A fi ; adiag = a(diag) ; pk = 0 ; sums: |res| , res*(res/adiag)
template <bool JACOBI>
__global__ void __launch_bounds__(FCP_TPB) k_cg_init(int32_t n, SellView m, const double *__restrict__ fi, const double *__restrict__ rhs,
                                                      double *__restrict__ res, double *__restrict__ adiag, double *__restrict__ pk, RedArgs ra) {
  double s[2] = {0.0, 0.0};
  FCP_ROW_LOOP(r, n) {
    const double rr = sell_row_sum<true>(m, fi, r, rhs[r]);
    res[r] = rr;
    const int64_t base = m.slptr[r >> 5] + (r & 31);
    const int32_t dpos = (m.rinfo[r] >> 16) & 0xffff;
    const double ad = m.a[base + (int64_t)dpos * 32];
    adiag[r] = ad;
    pk[r] = 0.0;
    s[0] = s[0] + fabs(rr);
    if (JACOBI) s[1] = s[1] + rr * (rr / ad);
  }
  finish_reduce<2>(s, ra);
}

// pk = zk + bet*pk with zk = res/adiag (dpcg :288-303) or the stored zk (iccg).
// cd != nullptr (multi-GPU, peer-memory path): `call exchange(pk)` (src-par/dpcg.f90:118) is fused in -- after its chunk is
// written the CTA stores the pk values of its process-face cells straight into the neighbours' LL slots over NVLink
// (sequence number = seq_base + iteration).  No fence, no flag, no extra kernel; chunks that own process faces are
// launched first so that the stores are under way while the bulk of the vector is still being updated.
template <bool JACOBI>
__global__ void __launch_bounds__(FCP_TPB) k_cg_pk(int32_t n, const double *__restrict__ res, const double *__restrict__ adiag,
                                                    const double *__restrict__ zk, double *pk, const KrylovScalars *sc, const CommDev *cd,
                                                    const int32_t *__restrict__ chunk_info, unsigned int seq_base) {
  const int32_t info = chunk_info ? __ldg(chunk_info + blockIdx.x) : (int32_t)blockIdx.x;
  if (sc->done) return;
  const double bet = sc->bet;
  const int chunk = info & 0x7fffffff;
  const int64_t base = (int64_t)chunk * FCP_CHUNK + threadIdx.x;
  if (base + (FCP_IPT - 1) * FCP_TPB < n) {
    // full chunk: all loads of the thread's 8 rows are issued before the first use (memory-level parallelism)
    double z_[FCP_IPT], d_[FCP_IPT], p_[FCP_IPT];
#pragma unroll
    for (int j = 0; j < FCP_IPT; ++j) {
      const int64_t r = base + j * FCP_TPB;
      if (JACOBI) { z_[j] = res[r]; d_[j] = adiag[r]; } else { z_[j] = zk[r]; d_[j] = 1.0; }
      p_[j] = pk[r];
    }
#pragma unroll
    for (int j = 0; j < FCP_IPT; ++j) {
      const double z = JACOBI ? (z_[j] / d_[j]) : z_[j];
      pk[base + j * FCP_TPB] = z + bet * p_[j];
    }
  } else {
    for (int j = 0; j < FCP_IPT; ++j) {
      const int64_t r = base + j * FCP_TPB;
      if (r < n) {
        const double z = JACOBI ? (res[r] / adiag[r]) : zk[r];
        pk[r] = z + bet * pk[r];
      }
    }
  }
  if (info >= 0) return;   // no halo bit: this chunk owns no process face
  const int32_t j0 = cd->chunk_ptr[chunk], j1 = cd->chunk_ptr[chunk + 1];
  __syncthreads();   // the chunk's pk values are written
  const unsigned int seq = seq_base + (unsigned int)sc->iters + 1u;
  for (int32_t j = j0 + (int32_t)threadIdx.x; j < j1; j += FCP_TPB) p2p_ll_store(cd->push_dst[j], pk[cd->push_cell[j]], seq);
}

// y = A x ; sums: sum v1*y [, sum v2*y | sum y*y]
template <int NS, bool SQ>
__global__ void __launch_bounds__(FCP_TPB) k_spmv_dot(int32_t n, SellView m, const double *__restrict__ x, double *__restrict__ y,
                                                       const double *__restrict__ v1, const KrylovScalars *sc, RedArgs ra, int fused, unsigned int seq_base) {
  const int32_t info = ra.chunk_info ? __ldg(ra.chunk_info + blockIdx.x) : (int32_t)blockIdx.x;
  if (sc->done) return;
  double s[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.0;
  const CommDev *cd = ra.cd;
  const int chunk = info & 0x7fffffff;
  const bool halo = fused && info < 0;
  const unsigned int seq = seq_base + (unsigned int)sc->iters + 1u;
  for (int j = 0; j < FCP_IPT; ++j) {
    const int64_t r64 = (int64_t)chunk * FCP_CHUNK + j * FCP_TPB + threadIdx.x;
    if (r64 >= n) continue;
    const int32_t r = (int32_t)r64;
    const double yr = halo ? sell_row_sum_halo<false>(m, x, r, 0.0, cd, seq) : sell_row_sum<false>(m, x, r, 0.0);
    y[r] = yr;
    s[0] = s[0] + v1[r] * yr;
    if (NS > 1 && SQ) s[NS - 1] = s[NS - 1] + yr * yr;
  }
  finish_reduce_part<NS>(s, ra, chunk, (int)gridDim.x);
}

// ---------------------------------------------------------------------------------------------
// k_spmv_dot_pipe: software-pipelined load/use SpMV.  A thread walks its 8 rows (stride 256); while the values and the
// x gathers of row j are in flight it already fetches the column indices of row j+1, so each row costs ONE memory
// round trip (indices of the next row, values and gathers of this row all overlap) instead of index -> gather chains.
// Register window of W entries per row; longer rows (polyhedral cells beyond W neighbours) finish in a plain loop.
// ---------------------------------------------------------------------------------------------
#ifndef FCP_PIPE_MINB
#define FCP_PIPE_MINB 4
#endif
template <int NS, bool SQ, int W>
__global__ void __launch_bounds__(FCP_TPB, (W <= 8 ? FCP_PIPE_MINB : 2)) k_spmv_dot_pipe(int32_t n, SellView m, const double *__restrict__ x, double *__restrict__ y,
                                                            const double *__restrict__ v1, const KrylovScalars *sc, RedArgs ra, int fused,
                                                            unsigned int seq_base) {
  const int32_t info = ra.chunk_info ? __ldg(ra.chunk_info + blockIdx.x) : (int32_t)blockIdx.x;
  if (sc->done) return;
  double s[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.0;
  const CommDev *cd = ra.cd;
  const int chunk = info & 0x7fffffff;
  const bool halo = fused && info < 0;
  const int64_t base = (int64_t)chunk * FCP_CHUNK + threadIdx.x;
  if (base + (FCP_IPT - 1) * FCP_TPB < n) {
    // chunks that own process faces (halo, CTA-uniform): the pipelined part covers the LOCAL entries of a row; its ghost
    // entries (stored last in the row, src-par/dpcg.f90:129-143) are added afterwards from the LL slots
    const unsigned int seq = seq_base + (unsigned int)sc->iters + 1u;
    const int lane = threadIdx.x & 31;
    int64_t pos = __ldg(&m.slptr[base >> 5]) + lane;      // SELL position of the current row's first entry
    int32_t len = halo ? __ldg(&m.llen[base]) : (__ldg(&m.rinfo[base]) & 0xffff);
    int32_t c[W];
#pragma unroll
    for (int k = 0; k < W; ++k) c[k] = (k < len) ? __ldcs(m.ja + pos + (int64_t)k * 32) : 0;
#pragma unroll 1
    for (int j = 0; j < FCP_IPT; ++j) {
      const int64_t r = base + (int64_t)j * FCP_TPB;
      // this row: values and gathers (all independent)
      double av[W], xv[W];
#pragma unroll
      for (int k = 0; k < W; ++k) av[k] = (k < len) ? __ldcs(m.a + pos + (int64_t)k * 32) : 0.0;
#pragma unroll
      for (int k = 0; k < W; ++k) xv[k] = (k < len) ? __ldg(x + c[k]) : 0.0;
      const double vv = v1[r];
      // next row: meta data and column indices
      int64_t npos = 0;
      int32_t nlen = 0;
      int32_t cn[W];
      if (j + 1 < FCP_IPT) {
        const int64_t rn = r + FCP_TPB;
        npos = __ldg(&m.slptr[rn >> 5]) + lane;
        nlen = halo ? __ldg(&m.llen[rn]) : (__ldg(&m.rinfo[rn]) & 0xffff);
#pragma unroll
        for (int k = 0; k < W; ++k) cn[k] = (k < nlen) ? __ldcs(m.ja + npos + (int64_t)k * 32) : 0;
      }
      double yr = 0.0;
#pragma unroll
      for (int k = 0; k < W; ++k)
        if (k < len) yr = yr + av[k] * xv[k];
      for (int32_t k = W; k < len; ++k) yr = yr + __ldcs(m.a + pos + (int64_t)k * 32) * __ldg(x + __ldcs(m.ja + pos + (int64_t)k * 32));
      if (halo) {
        const int32_t full = __ldg(&m.rinfo[r]) & 0xffff;
        for (int32_t k = len; k < full; ++k) {
          const int32_t cg = __ldcs(m.ja + pos + (int64_t)k * 32);
          yr = yr + __ldcs(m.a + pos + (int64_t)k * 32) * p2p_ll_load(cd->ll + 2 * (size_t)__ldg(cd->ghost_ord + (cg - n)), seq, cd->hdr);
        }
      }
      y[r] = yr;
      s[0] = s[0] + vv * yr;
      if (NS > 1 && SQ) s[NS - 1] = s[NS - 1] + yr * yr;
      pos = npos;
      len = nlen;
#pragma unroll
      for (int k = 0; k < W; ++k) c[k] = cn[k];
    }
  } else {
    // the few chunks that own process faces (launched first) and the ragged last chunk
    const unsigned int seq = seq_base + (unsigned int)sc->iters + 1u;
    for (int j = 0; j < FCP_IPT; ++j) {
      const int64_t r64 = base + (int64_t)j * FCP_TPB;
      if (r64 >= n) continue;
      const int32_t r = (int32_t)r64;
      const double yr = halo ? sell_row_sum_halo<false>(m, x, r, 0.0, cd, seq) : sell_row_sum<false>(m, x, r, 0.0);
      y[r] = yr;
      s[0] = s[0] + v1[r] * yr;
      if (NS > 1 && SQ) s[NS - 1] = s[NS - 1] + yr * yr;
    }
  }
  finish_reduce_part<NS>(s, ra, chunk, (int)gridDim.x);
}

// ---------------------------------------------------------------------------------------------
// k_spmv_dot_tma: the same SpMV + dot with the matrix stream (93 % of the bytes) moved by the TMA engine.
// A CTA owns one chunk = 8 tiles of 256 rows (8 SELL slices each).  The values and column indices of a tile are two
// contiguous, 256-/128-byte aligned blocks, fetched with cp.async.bulk into a ring of shared-memory stages that
// complete on an mbarrier (expect_tx); threads read them back with conflict-free LDS (lane l -> word l) and only the
// x gathers go through LSU/L1.  No registers or LSU request slots are tied up by the matrix stream, so many more bytes
// are in flight per SM than the load/use version can keep.  Arithmetic order per row is unchanged (CSR order).
// ---------------------------------------------------------------------------------------------
#ifdef FCP_EMU   // tests/emu: the mbarrier / bulk-copy primitives restated for the CPU emulation (test infrastructure only)
struct EmuMbar { uint32_t tx; uint8_t count, pending, phase, pad; };   // fits the 8-byte barrier word
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { EmuMbar *m = (EmuMbar *)bar; m->count = m->pending = count; m->tx = 0; m->phase = 0; }
__device__ __forceinline__ void emu_mbar_complete(EmuMbar *m) { if (m->pending == 0 && m->tx == 0) { m->phase ^= 1u; m->pending = m->count; } }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { EmuMbar *m = (EmuMbar *)bar; m->tx += bytes; m->pending -= 1; emu_mbar_complete(m); }
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) { EmuMbar *m = (EmuMbar *)bar; memcpy(dst, src, bytes); m->tx -= bytes; emu_mbar_complete(m); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { EmuMbar *m = (EmuMbar *)bar; while (m->phase == parity) emu::yield(); }
__device__ __forceinline__ void mbar_fence_init() {}
#define FCP_DYN_SMEM(name) unsigned char *name = emu::dyn_smem()
#else
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

#define FCP_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

#define FCP_TILE_ROWS FCP_TPB                      // 256 rows = 8 slices per tile
#define FCP_TILES (FCP_CHUNK / FCP_TILE_ROWS)      // 8 tiles per chunk
#define FCP_MAX_STAGES 4


// Persistent CTAs: CTA b owns chunks b, b + gridDim.x, ... and keeps ONE pipeline running across its chunks (the ring is
// refilled for the next chunk while the current one is still being consumed), so the start-up latency (slice pointers,
// first TMA round trip) is paid once per CTA instead of once per chunk.
template <int NS, bool SQ>
__global__ void __launch_bounds__(FCP_TPB) k_spmv_dot_tma(int32_t n, int32_t nslices, SellView m, const double *__restrict__ x, double *__restrict__ y,
                                                           const double *__restrict__ v1, const KrylovScalars *sc, RedArgs ra, int cap, int nstages) {
  if (sc->done) return;
  FCP_DYN_SMEM(smem_raw);
  // layout: [nstages][cap] int32 column indices | barriers | slice pointers of the current and the next chunk
  int32_t *sj = reinterpret_cast<int32_t *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)nstages * cap * 4);
  int64_t(*soff)[FCP_CHUNK / 32 + 1] = reinterpret_cast<int64_t(*)[FCP_CHUNK / 32 + 1]>(full + FCP_MAX_STAGES);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunks = (int)(((int64_t)n + FCP_CHUNK - 1) / FCP_CHUNK);
  const int my_chunks = ((int)blockIdx.x < nchunks) ? (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total_tiles = my_chunks * FCP_TILES;
  if (my_chunks == 0) return;
  if (tid <= FCP_CHUNK / 32) soff[0][tid] = m.slptr[min((int)blockIdx.x * (FCP_CHUNK / 32) + tid, nslices)];
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  int32_t ri_next = ((int64_t)blockIdx.x * FCP_CHUNK + tid < n) ? __ldg(&m.rinfo[(int64_t)blockIdx.x * FCP_CHUNK + tid]) : 0;
  __syncthreads();
  auto issue = [&](int q) {   // thread 0 only
    const int st = q % nstages, ci = q >> 3, t = q & 7;
    const int64_t b = soff[ci & 1][t * 8], e = soff[ci & 1][t * 8 + 8];
    const uint32_t cnt = (uint32_t)(e - b);
    if (cnt == 0) { mbar_expect_tx(&full[st], 0); return; }
    mbar_expect_tx(&full[st], cnt * 4u);
    tma_bulk_g2s(sj + (size_t)st * cap, m.ja + b, cnt * 4u, &full[st]);
  };
  if (tid == 0)
    for (int q = 0; q < nstages - 1 && q < total_tiles; ++q) issue(q);

  double s[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.0;
#pragma unroll 1
  for (int q = 0; q < total_tiles; ++q) {
    const int st = q % nstages, ci = q >> 3, t = q & 7;
    const int chunk = (int)blockIdx.x + ci * (int)gridDim.x;
    if (t == 0 && ci + 1 < my_chunks && tid <= FCP_CHUNK / 32)   // slice pointers of the next chunk (visible after this tile's barrier)
      soff[(ci + 1) & 1][tid] = m.slptr[min((chunk + (int)gridDim.x) * (FCP_CHUNK / 32) + tid, nslices)];
    if (tid == 0 && q + nstages - 1 < total_tiles) issue(q + nstages - 1);   // its stage was released by the barrier that ended tile q-1
    const int64_t r = (int64_t)chunk * FCP_CHUNK + (int64_t)t * FCP_TILE_ROWS + tid;
    const bool live = r < n;
    const double vv = live ? v1[r] : 0.0;
    const int32_t len = ri_next & 0xffff;
    {
      const int qn = q + 1, cn = (int)blockIdx.x + (qn >> 3) * (int)gridDim.x;
      const int64_t rn = (int64_t)cn * FCP_CHUNK + (int64_t)(qn & 7) * FCP_TILE_ROWS + tid;
      ri_next = (qn < total_tiles && rn < n) ? __ldg(&m.rinfo[rn]) : 0;
    }
    const int32_t lbase = (int32_t)(soff[ci & 1][t * 8 + warp] - soff[ci & 1][t * 8]) + lane;
    const double *__restrict__ ap = m.a + soff[ci & 1][t * 8 + warp] + lane;   // values: plain coalesced loads, independent of the indices
    const int32_t *jp = sj + (size_t)st * cap + lbase;
    mbar_wait(&full[st], (uint32_t)((q / nstages) & 1));
    double yr = 0.0;
#pragma unroll 8
    for (int32_t k = 0; k < len; ++k) {
      const double av = __ldcs(ap + (int64_t)k * 32);
      const int32_t c = jp[k * 32];
      yr = yr + av * __ldg(x + c);
    }
    if (live) {
      y[r] = yr;
      s[0] = s[0] + vv * yr;
      if (NS > 1 && SQ) s[NS - 1] = s[NS - 1] + yr * yr;
    }
    __syncthreads();   // every thread is done with stage st
    if (t == FCP_TILES - 1) {
      finish_reduce_part<NS>(s, ra, chunk, nchunks);
#pragma unroll
      for (int k = 0; k < NS; ++k) s[k] = 0.0;
    }
  }
}

// fi += alf*pk ; res -= alf*zk ; sums: |res| , [res*(res/adiag)] , [|adiag*fi|]      (:321-340)
template <bool JACOBI>
__global__ void __launch_bounds__(FCP_TPB) k_cg_update(int32_t n, double *__restrict__ fi, double *__restrict__ res, const double *__restrict__ pk,
                                                        const double *__restrict__ zk, const double *__restrict__ adiag,
                                                        const KrylovScalars *sc, RedArgs ra) {
  if (sc->done) return;
  const double alf = sc->alf;
  const bool first = (sc->iters == 0);
  double s[3] = {0.0, 0.0, 0.0};
  const int64_t base = (int64_t)blockIdx.x * FCP_CHUNK + threadIdx.x;
  if (base + (FCP_IPT - 1) * FCP_TPB < n) {
    // full chunk, two halves of 4 rows: 20 independent loads in flight per thread
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      double f_[4], r_[4], p_[4], z_[4], d_[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t r = base + (h * 4 + j) * FCP_TPB;
        f_[j] = fi[r]; r_[j] = res[r]; p_[j] = pk[r]; z_[j] = zk[r];
        d_[j] = (JACOBI || first) ? adiag[r] : 1.0;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t r = base + (h * 4 + j) * FCP_TPB;
        const double f = f_[j] + alf * p_[j];
        const double rr = r_[j] - alf * z_[j];
        fi[r] = f;
        res[r] = rr;
        s[0] = s[0] + fabs(rr);
        if (JACOBI) s[1] = s[1] + rr * (rr / d_[j]);
        if (first) s[2] = s[2] + fabs(d_[j] * f);
      }
    }
  } else {
    FCP_ROW_LOOP(r, n) {
      const double f = fi[r] + alf * pk[r];
      const double rr = res[r] - alf * zk[r];
      fi[r] = f;
      res[r] = rr;
      s[0] = s[0] + fabs(rr);
      const double ad = adiag[r];
      if (JACOBI) s[1] = s[1] + rr * (rr / ad);
      if (first) s[2] = s[2] + fabs(ad * f);
    }
  }
  finish_reduce<3>(s, ra);
}

// sum a*b (iccg sk = sum res*zk)
__global__ void __launch_bounds__(FCP_TPB) k_dot(int32_t n, const double *__restrict__ a, const double *__restrict__ b, const KrylovScalars *sc, RedArgs ra) {
  if (sc->done) return;
  double s[1] = {0.0};
  FCP_ROW_LOOP(r, n) { s[0] = s[0] + a[r] * b[r]; }
  finish_reduce<1>(s, ra);
}

// ---------------------------------------------------------------------------------------------
// IC(0)/ILU(0) diagonal-only factor and the preconditioner apply, level scheduled.
// One persistent cooperative grid; grid.sync() between levels.
// ---------------------------------------------------------------------------------------------
struct LevelView {
  const int32_t *lev_ptr, *lev_rows, *blev_ptr, *blev_rows;
  int32_t nlevels, nblevels;
};

// d(i) = 1/(a_ii - sum_{k<diag} a_k^2 d(ja_k))                       iccg   :439-445
// d(i) = 1/(a_ii - sum_{k<diag} a_k d(ja_k) a_T(k))                  bicgstab :613-624
template <bool ILU>
__global__ void __launch_bounds__(FCP_TPB) k_factor_diag(SellView m, LevelView lv, const int32_t *__restrict__ tpos, double *d) {
  cg::grid_group grid = cg::this_grid();
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int32_t L = 0; L < lv.nlevels; ++L) {
    const int32_t b = lv.lev_ptr[L], e = lv.lev_ptr[L + 1];
    for (int64_t q = b + gtid; q < e; q += gsz) {
      const int32_t i = lv.lev_rows[q];
      const int64_t base = m.slptr[i >> 5] + (i & 31);
      const int32_t dpos = (m.rinfo[i] >> 16) & 0xffff;
      double di = m.a[base + (int64_t)dpos * 32];
      for (int32_t k = 0; k < dpos; ++k) {
        const int64_t pos = base + (int64_t)k * 32;
        const double ak = m.a[pos];
        const double dj = __ldcg(&d[m.ja[pos]]);
        if (ILU) {
          const int32_t tp = tpos[pos];
          const double at = tp >= 0 ? m.a[tp] : 0.0;
          di = di - ak * dj * at;
        } else {
          di = di - ak * ak * dj;
        }
      }
      d[i] = 1.0 / di;
    }
    grid.sync();
  }
}

// zk = M^-1 rhs : forward sweep, zk/(d+small), backward sweep      (:458-475 ; quirk Q4 kept)
__global__ void __launch_bounds__(FCP_TPB) k_precond_apply(SellView m, LevelView lv, const int32_t *__restrict__ llen, const double *__restrict__ d,
                                                            const double *__restrict__ rhs, double *zk, const KrylovScalars *sc) {
  if (sc->done) return;
  cg::grid_group grid = cg::this_grid();
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int32_t L = 0; L < lv.nlevels; ++L) {
    const int32_t b = lv.lev_ptr[L], e = lv.lev_ptr[L + 1];
    for (int64_t q = b + gtid; q < e; q += gsz) {
      const int32_t i = lv.lev_rows[q];
      const int64_t base = m.slptr[i >> 5] + (i & 31);
      const int32_t dpos = (m.rinfo[i] >> 16) & 0xffff;
      double z = rhs[i];
      for (int32_t k = 0; k < dpos; ++k) {
        const int64_t pos = base + (int64_t)k * 32;
        z = z - m.a[pos] * __ldcg(&zk[m.ja[pos]]);
      }
      zk[i] = z * d[i];
    }
    grid.sync();
  }
  for (int32_t L = 0; L < lv.nblevels; ++L) {
    const int32_t b = lv.blev_ptr[L], e = lv.blev_ptr[L + 1];
    for (int64_t q = b + gtid; q < e; q += gsz) {
      const int32_t i = lv.blev_rows[q];
      const int64_t base = m.slptr[i >> 5] + (i & 31);
      const int32_t ri = m.rinfo[i];
      const int32_t dpos = (ri >> 16) & 0xffff;
      const int32_t len = llen ? llen[i] : (ri & 0xffff);   // local columns only: block-Jacobi across ranks
      const double di = d[i];
      double z = __ldcg(&zk[i]) / (di + FCP_SMALL);
      for (int32_t k = dpos + 1; k < len; ++k) {
        const int64_t pos = base + (int64_t)k * 32;
        z = z - m.a[pos] * __ldcg(&zk[m.ja[pos]]);
      }
      zk[i] = z * di;
    }
    grid.sync();
  }
}

// ---------------------------------------------------------------------------------------------
// The same preconditioner apply WITHOUT a grid barrier per level (FCP_SWEEP=flags; default is the barrier version above until this one
// has been timed on a B200).  Rows are taken in level order, 32 consecutive list positions per warp, tiles dealt round-robin to the resident
// warps; a row spins on the ready flag (= epoch of the sweep that finished it) of every row it depends on.  Every level starts on a warp
// boundary (pattern.cu), so a dependency always lies in an EARLIER tile: the lowest unfinished tile never waits, all warps are co-resident
// (cooperative launch), hence no deadlock.  One grid barrier remains between the forward and the backward sweep (the backward result
// overwrites zk(i), which later forward rows still read).  Per-row arithmetic and summation order are those of k_precond_apply: same bits.
// ---------------------------------------------------------------------------------------------
struct FlagView {
  const int32_t *prow, *pbrow;
  int32_t np, nbp;
  int32_t *ready;
  int32_t epoch;
};
// spins until *flag >= target; gives up after ~2 s (a scheduling assumption that does not hold must not hang the GPU): raises sc->pad, which
// krylov_solve turns into an error
__device__ __forceinline__ void sweep_wait(const int32_t *flag, int32_t target, const KrylovScalars *sc) {
#ifdef FCP_EMU
  (void)sc;
  while (__atomic_load_n(flag, __ATOMIC_ACQUIRE) < target) { emu::yield(); emu::os_yield(); }
#else
  int32_t v;
  unsigned int spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= target) break;
    if ((++spins & 1023u) == 0u) {
      if (*(volatile const int32_t *)&sc->pad) break;
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (!t0) t0 = t;
      else if (t - t0 > 2000000000ull) { *(volatile int32_t *)&const_cast<KrylovScalars *>(sc)->pad = 1; break; }
    }
  }
#endif
}
__device__ __forceinline__ void sweep_post(int32_t *flag, int32_t epoch) {
#ifdef FCP_EMU
  __atomic_store_n(flag, epoch, __ATOMIC_RELEASE);
#else
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
#endif
}
__global__ void __launch_bounds__(FCP_TPB) k_precond_apply_flags(SellView m, FlagView fv, const int32_t *__restrict__ llen, const double *__restrict__ d,
                                                                  const double *__restrict__ rhs, double *zk, const KrylovScalars *sc) {
  if (sc->done) return;
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int32_t ef = fv.epoch, eb = fv.epoch + 1;
  for (int64_t t = warp; t * 32 < fv.np; t += nwarps) {
    const int32_t i = fv.prow[t * 32 + lane];
    if (i >= 0) {
      const int64_t base = m.slptr[i >> 5] + (i & 31);
      const int32_t dpos = (m.rinfo[i] >> 16) & 0xffff;
      double z = rhs[i];
      for (int32_t k = 0; k < dpos; ++k) {
        const int64_t pos = base + (int64_t)k * 32;
        const int32_t j = m.ja[pos];
        sweep_wait(fv.ready + j, ef, sc);
        z = z - m.a[pos] * __ldcg(&zk[j]);
      }
      zk[i] = z * d[i];
      sweep_post(fv.ready + i, ef);
    }
    __syncwarp();
  }
  grid.sync();
  for (int64_t t = warp; t * 32 < fv.nbp; t += nwarps) {
    const int32_t i = fv.pbrow[t * 32 + lane];
    if (i >= 0) {
      const int64_t base = m.slptr[i >> 5] + (i & 31);
      const int32_t ri = m.rinfo[i];
      const int32_t dpos = (ri >> 16) & 0xffff;
      const int32_t len = llen ? llen[i] : (ri & 0xffff);
      const double di = d[i];
      double z = __ldcg(&zk[i]) / (di + FCP_SMALL);
      for (int32_t k = dpos + 1; k < len; ++k) {
        const int64_t pos = base + (int64_t)k * 32;
        const int32_t j = m.ja[pos];
        sweep_wait(fv.ready + j, eb, sc);
        z = z - m.a[pos] * __ldcg(&zk[j]);
      }
      zk[i] = z * di;
      sweep_post(fv.ready + i, eb);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Gauss-Seidel, linear_solvers.f90:96-201.  One sweep = the sequential loop :139-145
//     res(i) = rhs(i) - sum_k a(k) fi(ja(k)) ;  fi(i) = fi(i) + res(i)/(a(diag(i)) + small)
// in which row i sees the NEW values of the rows before it and the OLD values of itself and the rows after it.  Level scheduled over the
// lower triangle like the IC(0) forward sweep; because a row of an earlier level may have a HIGHER index than a row that still needs its
// old value, the new values go to a second array (xn) and the row sum reads xn for columns < i and fi for columns >= i -- in CSR (column)
// order, so every row rounds like the reference's.  k_gs_norms then publishes xn as fi and reduces sum|res| and sum|a_ii fi_i|.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FCP_TPB) k_gs_sweep(SellView m, LevelView lv, const double *__restrict__ rhs, const double *fi, double *xn,
                                                       double *__restrict__ res, double *__restrict__ adiag, const KrylovScalars *sc) {
  if (sc->done) return;
  cg::grid_group grid = cg::this_grid();
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int32_t L = 0; L < lv.nlevels; ++L) {
    const int32_t b = lv.lev_ptr[L], e = lv.lev_ptr[L + 1];
    for (int64_t q = b + gtid; q < e; q += gsz) {
      const int32_t i = lv.lev_rows[q];
      const int64_t base = m.slptr[i >> 5] + (i & 31);
      const int32_t ri = m.rinfo[i];
      const int32_t dpos = (ri >> 16) & 0xffff, len = ri & 0xffff;
      double r = rhs[i];
      for (int32_t k = 0; k < dpos; ++k) {
        const int64_t pos = base + (int64_t)k * 32;
        r = r - m.a[pos] * __ldcg(&xn[m.ja[pos]]);
      }
      for (int32_t k = dpos; k < len; ++k) {
        const int64_t pos = base + (int64_t)k * 32;
        r = r - m.a[pos] * fi[m.ja[pos]];
      }
      const double ad = m.a[base + (int64_t)dpos * 32];
      res[i] = r;
      adiag[i] = ad;
      xn[i] = fi[i] + r / (ad + FCP_SMALL);
    }
    grid.sync();
  }
}
__global__ void __launch_bounds__(FCP_TPB) k_gs_norms(int32_t n, double *__restrict__ fi, const double *__restrict__ xn, const double *__restrict__ res,
                                                       const double *__restrict__ adiag, const KrylovScalars *sc, RedArgs ra) {
  if (sc->done) return;
  const bool first = (sc->iters == 0);
  double s[2] = {0.0, 0.0};
  FCP_ROW_LOOP(r, n) {
    const double f = xn[r];
    fi[r] = f;
    s[0] = s[0] + fabs(res[r]);
    if (first) s[1] = s[1] + fabs(adiag[r] * f);
  }
  finish_reduce<2>(s, ra);
}

// ---------------------------------------------------------------------------------------------
// BiCGStab element kernels
// ---------------------------------------------------------------------------------------------
// reso = res ; pk = uk = 0 are set by the host (memset) ; this kernel: res = rhs - A fi, adiag, sums |res|, res*res
__global__ void __launch_bounds__(FCP_TPB) k_bicg_init(int32_t n, SellView m, const double *__restrict__ fi, const double *__restrict__ rhs,
                                                        double *__restrict__ res, double *__restrict__ reso, double *__restrict__ adiag, RedArgs ra) {
  double s[2] = {0.0, 0.0};
  FCP_ROW_LOOP(r, n) {
    const double rr = sell_row_sum<true>(m, fi, r, rhs[r]);
    res[r] = rr;
    reso[r] = rr;
    const int64_t base = m.slptr[r >> 5] + (r & 31);
    const int32_t dpos = (m.rinfo[r] >> 16) & 0xffff;
    adiag[r] = m.a[base + (int64_t)dpos * 32];
    s[0] = s[0] + fabs(rr);
    s[1] = s[1] + rr * rr;
  }
  finish_reduce<2>(s, ra);
}
// pk = res + om*(pk - alf*uk)     (:658)
__global__ void __launch_bounds__(FCP_TPB) k_bicg_pk(int32_t n, const double *__restrict__ res, const double *__restrict__ uk, double *__restrict__ pk,
                                                      const KrylovScalars *sc) {
  if (sc->done) return;
  const double om = sc->om, alf = sc->alf;   // alf of the previous iteration (1.0 at the start), :640,658
  FCP_ROW_LOOP(r, n) { pk[r] = res[r] + om * (pk[r] - alf * uk[r]); }
}
// fi += gam*zk ; res -= gam*uk    (:707-708)
__global__ void __launch_bounds__(FCP_TPB) k_bicg_half(int32_t n, double *__restrict__ fi, double *__restrict__ res, const double *__restrict__ zk,
                                                        const double *__restrict__ uk, const KrylovScalars *sc) {
  if (sc->done) return;
  const double gam = sc->gam;
  FCP_ROW_LOOP(r, n) {
    fi[r] = fi[r] + gam * zk[r];
    res[r] = res[r] - gam * uk[r];
  }
}
// fi += alf*zk ; res -= alf*vk ; sums |res|, res*reso, [|adiag fi|]    (:750-773)
__global__ void __launch_bounds__(FCP_TPB) k_bicg_update(int32_t n, double *__restrict__ fi, double *__restrict__ res, const double *__restrict__ zk,
                                                          const double *__restrict__ vk, const double *__restrict__ reso, const double *__restrict__ adiag,
                                                          const KrylovScalars *sc, RedArgs ra) {
  if (sc->done) return;
  const double alf = sc->alf;
  const bool first = (sc->iters == 0);
  double s[3] = {0.0, 0.0, 0.0};
  FCP_ROW_LOOP(r, n) {
    const double f = fi[r] + alf * zk[r];
    const double rr = res[r] - alf * vk[r];
    fi[r] = f;
    res[r] = rr;
    s[0] = s[0] + fabs(rr);
    s[1] = s[1] + rr * reso[r];
    if (first) s[2] = s[2] + fabs(adiag[r] * f);
  }
  finish_reduce<3>(s, ra);
}
// y = A x ; sums: y*v1 , y*y  or y*v1 only
// (k_spmv_dot above)

// ---------------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------------
struct Launcher {
  cudaStream_t st;
  FcpComm *comm;
  fcp_ctx *ctx;
  KrylovWS &ws;
  int n;
  int grid;
  const CommDev *cd;   // peer-memory path (nullptr: single GPU or NCCL path)
  RedArgs red(int epi) const { return RedArgs{ws.partials, ws.maxchunks, ws.counter, ws.sc, epi, comm ? (cd ? 2 : 0) : 1, cd, comm_chunk_info(comm)}; }
  // after a reducing kernel on the NCCL path: cross-rank sum of sc->red[0..ns) + epilogue kernel
  int post(int epi, int ns) const {
    if (!comm || cd) return FCP_OK;
    FCP_TRY(comm_allgather_sum(comm, ws.sc->red, ns, st));
    k_epilogue<<<1, 1, 0, st>>>(epi, ws.sc);
    FCP_LAUNCHED();
    FCP_CHECK_LAUNCH();
    return FCP_OK;
  }
  int halo(double *x) const {   // src-par/dpcg.f90:118  call exchange(pk)
    if (!comm) return FCP_OK;
    return comm_exchange(ctx, x, 1);
  }
  int halo_pk() const {         // exchange(pk): fused into k_cg_pk / the SpMV on the peer-memory path
    if (!comm || cd) return FCP_OK;
    return comm_exchange(ctx, ws.pk, 1);
  }
};

static int coop_grid(const void *kernel, int device, int *grid) {
  int nsm = 0, occ = 0;
  FCP_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  FCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, FCP_TPB, 0));
  if (occ < 1) { fcp_set_error("cooperative kernel does not fit on an SM"); return FCP_ECUDA; }
  *grid = nsm * occ;
  return FCP_OK;
}

static int launch_factor(bool ilu, SellPattern &p, const double *a, double *d, cudaStream_t st) {
  if (p.n == 0) return FCP_OK;
  int dev = 0, grid = 0;
  FCP_CUDA(cudaGetDevice(&dev));
  SellView m{p.slptr, p.rinfo, p.ja, a, p.llen};
  LevelView lv{p.lev_ptr, p.lev_rows, p.blev_ptr, p.blev_rows, p.nlevels, p.nblevels};
  const int32_t *tpos = p.tpos;
  void *args[] = {&m, &lv, &tpos, &d};
  const void *fn = ilu ? (const void *)k_factor_diag<true> : (const void *)k_factor_diag<false>;
  FCP_TRY(coop_grid(fn, dev, &grid));
#ifdef FCP_EMU
  (void)args;
  if (ilu) emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_factor_diag<true>, m, lv, tpos, d);
  else emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_factor_diag<false>, m, lv, tpos, d);
#else
  FCP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(FCP_TPB), args, 0, st));
#endif
  FCP_LAUNCHED();
  return FCP_OK;
}
static int launch_precond(SellPattern &p, const double *a, const double *d, const double *rhs, double *zk, const KrylovScalars *sc,
                          cudaStream_t st) {
  if (p.n == 0) return FCP_OK;
  static int grid = 0;
  if (!grid) {
    int dev = 0;
    FCP_CUDA(cudaGetDevice(&dev));
    FCP_TRY(coop_grid((const void *)k_precond_apply, dev, &grid));
  }
  SellView m{p.slptr, p.rinfo, p.ja, a, p.llen};
  LevelView lv{p.lev_ptr, p.lev_rows, p.blev_ptr, p.blev_rows, p.nlevels, p.nblevels};
  const int32_t *llen = p.llen;
  const char *sweep_env = getenv("FCP_SWEEP");       // read per call: tests and A/B timings switch it inside one process
  const bool use_flags = sweep_env && !strcmp(sweep_env, "flags");
  static bool announced = false;
  if (use_flags && !announced) {
    announced = true;
    fprintf(stderr, "libfcp_b200: FCP_SWEEP=flags: IC(0)/ILU(0) sweeps wait on per-row ready flags instead of a grid barrier per level\n");
  }
  if (use_flags) {
    static int fgrid = 0;
    if (!fgrid) {
      int dev = 0;
      FCP_CUDA(cudaGetDevice(&dev));
      FCP_TRY(coop_grid((const void *)k_precond_apply_flags, dev, &fgrid));
    }
    if (p.sweep_epoch > 2000000000) {          // epochs are int32: start over (once per ~10^9 applies)
      FCP_CUDA(cudaMemsetAsync(p.ready, 0, sizeof(int32_t) * (size_t)std::max(p.n, 1), st));
      p.sweep_epoch = 0;
    }
    FlagView fv{p.plev_rows, p.pblev_rows, p.nplev, p.npblev, p.ready, p.sweep_epoch + 1};
    p.sweep_epoch += 2;
    void *fargs[] = {&m, &fv, &llen, &d, &rhs, &zk, &sc};
#ifdef FCP_EMU
    (void)fargs;
    emu::launch_coop(emu::Cfg(fgrid, FCP_TPB, 0, st), k_precond_apply_flags, m, fv, llen, d, rhs, zk, sc);
#else
    FCP_CUDA(cudaLaunchCooperativeKernel((const void *)k_precond_apply_flags, dim3(fgrid), dim3(FCP_TPB), fargs, 0, st));
#endif
    FCP_LAUNCHED();
    return FCP_OK;
  }
  void *args[] = {&m, &lv, &llen, &d, &rhs, &zk, &sc};
#ifdef FCP_EMU
  (void)args;
  emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_precond_apply, m, lv, llen, d, rhs, zk, sc);
#else
  FCP_CUDA(cudaLaunchCooperativeKernel((const void *)k_precond_apply, dim3(grid), dim3(FCP_TPB), args, 0, st));
#endif
  FCP_LAUNCHED();
  return FCP_OK;
}

static int launch_gs_sweep(SellPattern &p, const double *a, const double *rhs, const double *fi, double *xn, double *res, double *adiag,
                           const KrylovScalars *sc, cudaStream_t st) {
  if (p.n == 0) return FCP_OK;
  static int grid = 0;
  if (!grid) {
    int dev = 0;
    FCP_CUDA(cudaGetDevice(&dev));
    FCP_TRY(coop_grid((const void *)k_gs_sweep, dev, &grid));
  }
  SellView m{p.slptr, p.rinfo, p.ja, a, p.llen};
  LevelView lv{p.lev_ptr, p.lev_rows, p.blev_ptr, p.blev_rows, p.nlevels, p.nblevels};
  void *args[] = {&m, &lv, &rhs, &fi, &xn, &res, &adiag, &sc};
#ifdef FCP_EMU
  (void)args;
  emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_gs_sweep, m, lv, rhs, fi, xn, res, adiag, sc);
#else
  FCP_CUDA(cudaLaunchCooperativeKernel((const void *)k_gs_sweep, dim3(grid), dim3(FCP_TPB), args, 0, st));
#endif
  FCP_LAUNCHED();
  return FCP_OK;
}

// opt-in dynamic shared memory limit of the TMA kernels: one process-wide value, only ever raised
static int tma_smem_limit(size_t smem) {
  static size_t configured = 0;
  if (smem <= configured) return FCP_OK;
  FCP_CUDA(cudaFuncSetAttribute(k_spmv_dot_tma<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  FCP_CUDA(cudaFuncSetAttribute(k_spmv_dot_tma<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  configured = smem;
  return FCP_OK;
}

// SpMV + dot dispatch: TMA-staged kernel when a tile (256 rows) of the pattern fits the shared-memory ring, else the
// load/use kernel.  FCP_SPMV=ldg forces the latter (A/B measurements).
template <int NS, bool SQ>
static int launch_spmv_dot(const SellPattern &p, const SellView &m, const double *x, double *y, const double *v1, const KrylovScalars *sc,
                           const RedArgs &ra, cudaStream_t st, int fused = 0, unsigned int seq_base = 0) {
  const int grid = fcp_nchunks(p.n);
  if (!grid) return FCP_OK;
  static int mode = -1;   // 0 ldg, 1 tma
  static int stages_env = 0;
  if (mode < 0) {
    const char *e = getenv("FCP_SPMV");
    mode = (e && !strcmp(e, "ldg")) ? 0 : (e && !strcmp(e, "tma")) ? 1 : 2;
    const char *g = getenv("FCP_SPMV_STAGES");
    stages_env = g ? atoi(g) : 0;
  }
  const size_t stage_bytes = (size_t)p.tile_cap * 4;
  const size_t extra = FCP_MAX_STAGES * sizeof(uint64_t) + 2 * (FCP_CHUNK / 32 + 1) * sizeof(int64_t) + 128;
  int nst = stages_env ? stages_env : (int)std::min<size_t>(FCP_MAX_STAGES, (size_t)(24 * 1024) / std::max<size_t>(stage_bytes, 1));
  nst = std::min(nst, FCP_MAX_STAGES);
  if (mode == 1 && p.tile_cap > 0 && nst >= 2 && !ra.cd) {
    const size_t smem = nst * stage_bytes + extra;
    FCP_TRY(tma_smem_limit(smem));
    static int nsm = 0;
    if (!nsm) {
      int dev = 0;
      FCP_CUDA(cudaGetDevice(&dev));
      FCP_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    }
    int occ = 0;
    FCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_dot_tma<NS, SQ>, FCP_TPB, smem));
    const int pgrid = std::min(grid, std::max(1, occ) * nsm);
    k_spmv_dot_tma<NS, SQ><<<pgrid, FCP_TPB, smem, st>>>(p.n, p.nslices, m, x, y, v1, sc, ra, p.tile_cap, nst);
    FCP_CHECK_LAUNCH();
  } else if (mode != 0) {
    if (p.tile_cap <= 8 * 256) k_spmv_dot_pipe<NS, SQ, 8><<<grid, FCP_TPB, 0, st>>>(p.n, m, x, y, v1, sc, ra, fused, seq_base);
    else k_spmv_dot_pipe<NS, SQ, 16><<<grid, FCP_TPB, 0, st>>>(p.n, m, x, y, v1, sc, ra, fused, seq_base);
  } else {
    k_spmv_dot<NS, SQ><<<grid, FCP_TPB, 0, st>>>(p.n, m, x, y, v1, sc, ra, fused, seq_base);
  }
  FCP_LAUNCHED();
  return FCP_OK;
}

// poll the device scalars; returns done flag
static int fetch_scalars(KrylovWS &ws, cudaStream_t st) {
  FCP_CUDA(cudaMemcpyAsync(ws.h_sc, ws.sc, sizeof(KrylovScalars), cudaMemcpyDeviceToHost, st));
  FCP_CUDA(cudaStreamSynchronize(st));
  return FCP_OK;
}

int krylov_solve(int solver, SellPattern &p, const double *a, double *fi, const double *rhs, KrylovWS &ws, int32_t itr_max,
                 double tol_abs, double tol_rel, fcp_report *rep, cudaStream_t st, FcpComm *comm, fcp_ctx *ctx) {
  const int32_t n = p.n;
  if (solver != FCP_SOLVER_DPCG && solver != FCP_SOLVER_ICCG && solver != FCP_SOLVER_BICGSTAB && solver != FCP_SOLVER_GAUSS_SEIDEL) {
    fcp_set_error("csrsolve: unknown solver id %d", solver);
    return FCP_EINVAL;
  }
  const CommDev *cd = comm_dev(comm);
  FCP_TRY(krylov_ws_alloc(ws, n, p.ncols));
  if (rep) { memset(rep, 0, sizeof(*rep)); rep->solver = solver; }
  const int grid = fcp_nchunks(n);
  if (cd && grid == 0) { fcp_set_error("a partition without cells cannot take part in the peer-memory path"); return FCP_EINVAL; }
  Launcher L{st, comm, ctx, ws, n, grid, cd};
  const int fw = cd ? 1 : 0;   // the SpMV after k_cg_pk takes its ghost columns from the fused halo push
  const unsigned int sb = comm_pk_base(comm);
  Profiler *prof = ctx ? &ctx->prof : nullptr;
  SellView m{p.slptr, p.rinfo, p.ja, a, p.llen};
  KrylovScalars init;
  memset(&init, 0, sizeof(init));
  init.tol_abs = tol_abs;
  init.tol_rel = tol_rel;
  init.itr_max = itr_max;
  *ws.h_sc = init;
  FCP_CUDA(cudaMemcpyAsync(ws.sc, ws.h_sc, sizeof(KrylovScalars), cudaMemcpyHostToDevice, st));
  FCP_CUDA(cudaStreamSynchronize(st));   // h_sc is reused as the read-back buffer
  if (n == 0 && !comm) return FCP_OK;
  const int BATCH = 16;

  if (solver == FCP_SOLVER_GAUSS_SEIDEL) {
    if (comm) { fcp_set_error("csrsolve: 'gauss-seidel' is a serial-tree solver (src-par has none); not available with a communicator"); return FCP_EINVAL; }
    FCP_TRY(sell_build_levels(p, st));
    if (itr_max <= 0) {                       // the DO loop :137 runs zero times: nothing is touched
      ws.h_sc->done = 1;
    } else {
      for (int it = 0; it < itr_max;) {
        for (int b = 0; b < BATCH && it < itr_max; ++b, ++it) {
          FCP_PROF(prof, FCP_K_PRECOND, st, FCP_TRY(launch_gs_sweep(p, a, rhs, fi, ws.pk, ws.res, ws.adiag, ws.sc, st)));
          if (grid) { k_gs_norms<<<grid, FCP_TPB, 0, st>>>(n, fi, ws.pk, ws.res, ws.adiag, ws.sc, L.red(EPI_GS)); FCP_LAUNCHED(); }
        }
        FCP_CHECK_LAUNCH();
        FCP_TRY(fetch_scalars(ws, st));
        if (ws.h_sc->done) break;
      }
    }
  } else if (solver == FCP_SOLVER_DPCG) {
    FCP_TRY(L.halo(fi));
    if (grid) { k_cg_init<true><<<grid, FCP_TPB, 0, st>>>(n, m, fi, rhs, ws.res, ws.adiag, ws.pk, L.red(EPI_INIT_CG)); FCP_LAUNCHED(); FCP_CHECK_LAUNCH(); }
    FCP_TRY(L.post(EPI_INIT_CG, 2));
    for (int it = 0; it < itr_max;) {
      for (int b = 0; b < BATCH && it < itr_max; ++b, ++it) {
        if (grid) FCP_PROF(prof, FCP_K_CG_PK, st, (k_cg_pk<true><<<grid, FCP_TPB, 0, st>>>(n, ws.res, ws.adiag, ws.zk, ws.pk, ws.sc, cd, comm_chunk_info(comm), sb), FCP_LAUNCHED()));
        FCP_TRY(L.halo_pk());
        if (grid) FCP_PROF(prof, FCP_K_SPMV_DOT, st, FCP_TRY((launch_spmv_dot<1, false>(p, m, ws.pk, ws.zk, ws.pk, ws.sc, L.red(EPI_PKAPK), st, fw, sb))));
        FCP_TRY(L.post(EPI_PKAPK, 1));
        if (grid) FCP_PROF(prof, FCP_K_CG_UPDATE, st, (k_cg_update<true><<<grid, FCP_TPB, 0, st>>>(n, fi, ws.res, ws.pk, ws.zk, ws.adiag, ws.sc, L.red(EPI_CG_UPDATE)), FCP_LAUNCHED()));
        FCP_TRY(L.post(EPI_CG_UPDATE, 3));
      }
      FCP_CHECK_LAUNCH();
      FCP_TRY(fetch_scalars(ws, st));
      if (cd) FCP_TRY(comm_check_error(ctx));   // a peer stopped responding: fail now instead of spinning through every batch
      if (ws.h_sc->done) break;
    }
  } else if (solver == FCP_SOLVER_ICCG) {
    FCP_TRY(sell_build_levels(p, st));
    FCP_TRY(krylov_ws_need(ws, &ws.d, (size_t)n));
    FCP_TRY(L.halo(fi));
    if (grid) { k_cg_init<false><<<grid, FCP_TPB, 0, st>>>(n, m, fi, rhs, ws.res, ws.adiag, ws.pk, L.red(EPI_INIT_CG)); FCP_LAUNCHED(); FCP_CHECK_LAUNCH(); }
    FCP_TRY(L.post(EPI_INIT_CG, 2));
    FCP_TRY(fetch_scalars(ws, st));
    if (!ws.h_sc->done) {
      FCP_TRY(launch_factor(false, p, a, ws.d, st));
      for (int it = 0; it < itr_max;) {
        for (int b = 0; b < BATCH && it < itr_max; ++b, ++it) {
          FCP_PROF(prof, FCP_K_PRECOND, st, FCP_TRY(launch_precond(p, a, ws.d, ws.res, ws.zk, ws.sc, st)));
          if (grid) { k_dot<<<grid, FCP_TPB, 0, st>>>(n, ws.res, ws.zk, ws.sc, L.red(EPI_SK)); FCP_LAUNCHED(); }
          FCP_TRY(L.post(EPI_SK, 1));
          if (grid) FCP_PROF(prof, FCP_K_CG_PK, st, (k_cg_pk<false><<<grid, FCP_TPB, 0, st>>>(n, ws.res, ws.adiag, ws.zk, ws.pk, ws.sc, cd, comm_chunk_info(comm), sb), FCP_LAUNCHED()));
          FCP_TRY(L.halo_pk());
          if (grid) FCP_PROF(prof, FCP_K_SPMV_DOT, st, FCP_TRY((launch_spmv_dot<1, false>(p, m, ws.pk, ws.zk, ws.pk, ws.sc, L.red(EPI_PKAPK), st, fw, sb))));
          FCP_TRY(L.post(EPI_PKAPK, 1));
          if (grid) FCP_PROF(prof, FCP_K_CG_UPDATE, st, (k_cg_update<false><<<grid, FCP_TPB, 0, st>>>(n, fi, ws.res, ws.pk, ws.zk, ws.adiag, ws.sc, L.red(EPI_CG_UPDATE)), FCP_LAUNCHED()));
          FCP_TRY(L.post(EPI_CG_UPDATE, 3));
        }
        FCP_CHECK_LAUNCH();
        FCP_TRY(fetch_scalars(ws, st));
        if (cd) FCP_TRY(comm_check_error(ctx));
        if (ws.h_sc->done) break;
      }
    }
  } else {
    FCP_TRY(sell_build_levels(p, st));
    FCP_TRY(sell_build_tpos(p, st));
    FCP_TRY(krylov_ws_need(ws, &ws.d, (size_t)n));
    FCP_TRY(krylov_ws_need(ws, &ws.reso, (size_t)n));
    FCP_TRY(krylov_ws_need(ws, &ws.uk, (size_t)n));
    FCP_TRY(krylov_ws_need(ws, &ws.vk, (size_t)n));
    FCP_CUDA(cudaMemsetAsync(ws.pk, 0, sizeof(double) * (size_t)p.ncols, st));
    FCP_CUDA(cudaMemsetAsync(ws.uk, 0, sizeof(double) * (size_t)n, st));
    FCP_TRY(L.halo(fi));
    if (grid) { k_bicg_init<<<grid, FCP_TPB, 0, st>>>(n, m, fi, rhs, ws.res, ws.reso, ws.adiag, L.red(EPI_INIT_BICG)); FCP_LAUNCHED(); FCP_CHECK_LAUNCH(); }
    FCP_TRY(L.post(EPI_INIT_BICG, 2));
    FCP_TRY(fetch_scalars(ws, st));
    if (!ws.h_sc->done) {
      FCP_TRY(launch_factor(true, p, a, ws.d, st));
      for (int it = 0; it < itr_max;) {
        for (int b = 0; b < BATCH && it < itr_max; ++b, ++it) {
          if (grid) { k_bicg_pk<<<grid, FCP_TPB, 0, st>>>(n, ws.res, ws.uk, ws.pk, ws.sc); FCP_LAUNCHED(); }
          FCP_PROF(prof, FCP_K_PRECOND, st, FCP_TRY(launch_precond(p, a, ws.d, ws.pk, ws.zk, ws.sc, st)));
          FCP_TRY(L.halo(ws.zk));
          if (grid) FCP_PROF(prof, FCP_K_SPMV_DOT, st, FCP_TRY((launch_spmv_dot<1, false>(p, m, ws.zk, ws.uk, ws.reso, ws.sc, L.red(EPI_UKRESO), st))));
          FCP_TRY(L.post(EPI_UKRESO, 1));
          if (grid) { k_bicg_half<<<grid, FCP_TPB, 0, st>>>(n, fi, ws.res, ws.zk, ws.uk, ws.sc); FCP_LAUNCHED(); }
          FCP_PROF(prof, FCP_K_PRECOND, st, FCP_TRY(launch_precond(p, a, ws.d, ws.res, ws.zk, ws.sc, st)));
          FCP_TRY(L.halo(ws.zk));
          if (grid) FCP_PROF(prof, FCP_K_SPMV_DOT, st, FCP_TRY((launch_spmv_dot<2, true>(p, m, ws.zk, ws.vk, ws.res, ws.sc, L.red(EPI_VK), st))));
          FCP_TRY(L.post(EPI_VK, 2));
          if (grid) { k_bicg_update<<<grid, FCP_TPB, 0, st>>>(n, fi, ws.res, ws.zk, ws.vk, ws.reso, ws.adiag, ws.sc, L.red(EPI_BICG_UPDATE)); FCP_LAUNCHED(); }
          FCP_TRY(L.post(EPI_BICG_UPDATE, 3));
        }
        FCP_CHECK_LAUNCH();
        FCP_TRY(fetch_scalars(ws, st));
        if (cd) FCP_TRY(comm_check_error(ctx));
        if (ws.h_sc->done) break;
      }
    }
  }
  FCP_TRY(fetch_scalars(ws, st));
  if (ws.h_sc->pad) { fcp_set_error("csrsolve: a barrier-free sweep (FCP_SWEEP=flags) gave up waiting for a row it depends on"); return FCP_ECUDA; }
  if (comm) FCP_TRY(L.halo(fi));   // src-par/dpcg.f90:183  call exchange(fi)
  if (cd) {
    comm_pk_advance(comm, ws.h_sc->iters);
    FCP_TRY(comm_check_error(ctx));
  }
  if (rep) {
    rep->res0 = ws.h_sc->res0;
    rep->resl = ws.h_sc->resl;
    rep->factor = ws.h_sc->factor;
    rep->resor = ws.h_sc->resor;
    rep->iters = ws.h_sc->iters;
  }
  return FCP_OK;
}
