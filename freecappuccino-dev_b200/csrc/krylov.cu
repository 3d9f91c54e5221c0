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
// Programmatic dependent launch (FCP_PDL, default on): the three kernels of a CG iteration are launched with the programmatic-stream-
// serialization attribute and begin with griddepcontrol.wait.  A kernel's CTAs may then be scheduled while the LAST wave of its predecessor is
// still running (the predecessor's CTAs signal launch_dependents as their first instruction); they sit at the wait until the predecessor has
// completed and its memory is visible, so launch latency, CTA ramp-up and the predecessor's last-CTA reduction epilogue overlap.  Without the
// launch attribute both instructions are no-ops.  Nothing about the arithmetic changes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() {
#ifndef FCP_EMU
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch_dependents() {
#ifndef FCP_EMU
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
static bool pdl_wanted() {
  const char *e = getenv("FCP_PDL");
  return !(e && !strcmp(e, "off"));
}
#ifdef FCP_EMU
#define FCP_LAUNCH_PDL(pdl, kernel, grid, st, ...) kernel<<<grid, FCP_TPB, 0, st>>>(__VA_ARGS__)
#else
template <class... KA, class... A>
static void launch_maybe_pdl(bool pdl, void (*k)(KA...), int grid, cudaStream_t st, A &&...a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(FCP_TPB);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k, KA(a)...);
}
#define FCP_LAUNCH_PDL(pdl, kernel, grid, st, ...) launch_maybe_pdl(pdl, kernel, grid, st, __VA_ARGS__)
#endif

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
  FCP_TRY(dev_alloc(&ws.adiag, (size_t)ncols));       // ghost slots: the residual-halo scheme keeps the diagonal of the cells across process faces
  ws.maxchunks = fcp_nchunks(n) + 1;
  FCP_TRY(dev_alloc(&ws.partials, (size_t)4 * ws.maxchunks));
  FCP_TRY(dev_alloc(&ws.counter, 1));
  FCP_CUDA(cudaMemset(ws.counter, 0, sizeof(unsigned int)));
  FCP_TRY(dev_alloc(&ws.bar, 4));
  FCP_TRY(dev_alloc(&ws.phase_ns, 8));
  FCP_CUDA(cudaMemset(ws.bar, 0, 4 * sizeof(unsigned int)));
  FCP_CUDA(cudaMemset(ws.phase_ns, 0, 8 * sizeof(unsigned long long)));
  FCP_CUDA(cudaGetDevice(&ws.ws_device));
  FCP_CUDA(cudaStreamSynchronize(0));   // cudaMemset is asynchronous; the solver streams do not order with the legacy stream
  FCP_CUDA(cudaMalloc((void **)&ws.sc, sizeof(KrylovScalars)));
  FCP_CUDA(cudaMallocHost((void **)&ws.h_sc, sizeof(KrylovScalars)));
  FCP_CUDA(cudaMallocHost((void **)&ws.h_poll, 4 * sizeof(int32_t)));
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
  cudaFree(ws.partials); cudaFree(ws.counter); cudaFree(ws.sc); cudaFree(ws.bar); cudaFree(ws.phase_ns);
  if (ws.h_sc) cudaFreeHost(ws.h_sc);
  if (ws.h_poll) cudaFreeHost(ws.h_poll);
  if (ws.ev[0]) cudaEventDestroy(ws.ev[0]);
  if (ws.ev[1]) cudaEventDestroy(ws.ev[1]);
  ws = KrylovWS();
}

// ---------------------------------------------------------------------------------------------
// L2 residency hints.  On a partition whose Krylov vectors fit the 126 MB L2 (8 ranks at 256^3: 5 x 16.8 MB) the vectors are re-read every
// iteration while the matrix (176 MB per rank) streams through once per iteration and would evict them: the matrix stream is read evict-first
// (__ldcs, as before) and, in the HINT variants of the three CG kernels, every vector access carries an L2 cache-policy operand -- evict_last for
// the vectors the host selected (krylov_l2_masks: by reuse per iteration, while they fit the budget), evict_normal for the others.  A hint never
// changes a value: bits and iteration counts are those of the plain variants.  FCP_L2=off disables, FCP_L2_MB sets the budget (default 96).
// ---------------------------------------------------------------------------------------------
struct L2Pol { unsigned long long last, norm; };
__device__ __forceinline__ L2Pol l2pol_make() {
  L2Pol q;
#ifdef FCP_EMU
  q.last = q.norm = 0ull;
#else
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(q.last));
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(q.norm));
#endif
  return q;
}
__device__ __forceinline__ unsigned long long l2pol_pick(const L2Pol &q, unsigned int mask, int bit) { return (mask >> bit) & 1u ? q.last : q.norm; }
// vector read through the non-coherent path (the kernel does not write this vector)
__device__ __forceinline__ double l2_ld_nc(const double *p, unsigned long long pol) {
#ifdef FCP_EMU
  (void)pol;
  return *p;
#else
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
#endif
}
// vector read of an array this kernel also writes (its own rows only)
__device__ __forceinline__ double l2_ld(const double *p, unsigned long long pol) {
#ifdef FCP_EMU
  (void)pol;
  return *p;
#else
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void l2_st(double *p, double v, unsigned long long pol) {
#ifdef FCP_EMU
  (void)pol;
  *p = v;
#else
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
#endif
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
template <bool SUB, bool NC = true>
__device__ __forceinline__ double sell_row_sum(const SellView &m, const double *x, int32_t r, double s) {
  const int64_t base = __ldg(&m.slptr[r >> 5]) + (r & 31);
  const int32_t len = __ldg(&m.rinfo[r]) & 0xffff;
  const double *__restrict__ ap = m.a + base;
  const int32_t *__restrict__ jp = m.ja + base;
#pragma unroll 4
  for (int32_t k = 0; k < len; ++k) {
    const double av = __ldcs(ap + (int64_t)k * 32);
    const int32_t c = __ldcs(jp + (int64_t)k * 32);
    const double t = av * (NC ? __ldg(x + c) : x[c]);
    s = SUB ? (s - t) : (s + t);
  }
  return s;
}

// Row sum for the rows of a chunk that owns process faces, on the peer-memory path: a ghost column (>= n) is not read
// from x but from this rank's LL slots, where the neighbour's k_cg_pk stored it during this very iteration; the word's
// embedded sequence number tells when it has arrived (src-par/dpcg.f90:118 `call exchange(pk)` + :129-143 halo term).
template <bool SUB, bool NC = true>
__device__ __forceinline__ double sell_row_sum_halo(const SellView &m, const double *x, int32_t r, double s, const CommDev *cd,
                                                    unsigned int seq) {
  const int64_t base = __ldg(&m.slptr[r >> 5]) + (r & 31);
  const int32_t len = __ldg(&m.rinfo[r]) & 0xffff;
  const int32_t n = cd->n;
  for (int32_t k = 0; k < len; ++k) {
    const double av = __ldcs(m.a + base + (int64_t)k * 32);
    const int32_t c = __ldcs(m.ja + base + (int64_t)k * 32);
    const double xv = c >= n ? p2p_ll_load(cd->ll + 2 * (size_t)__ldg(cd->ghost_ord + (c - n)), seq, cd->hdr) : (NC ? __ldg(x + c) : x[c]);
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
// One chunk of the direction update.  `info` = chunk index | halo bit 31.
template <bool JACOBI>
__device__ __forceinline__ void cg_pk_chunk(int32_t n, const double *res, const double *__restrict__ adiag, const double *zk, double *pk, double bet,
                                            int32_t info, const CommDev *cd, unsigned int seq) {
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
  for (int32_t j = j0 + (int32_t)threadIdx.x; j < j1; j += FCP_TPB) p2p_ll_store(cd->push_dst[j], pk[cd->push_cell[j]], seq);
}
template <bool JACOBI>
__global__ void __launch_bounds__(FCP_TPB) k_cg_pk(int32_t n, const double *__restrict__ res, const double *__restrict__ adiag,
                                                    const double *__restrict__ zk, double *pk, const KrylovScalars *sc, const CommDev *cd,
                                                    const int32_t *__restrict__ chunk_info, unsigned int seq_base) {
  pdl_launch_dependents();
  const int32_t info = chunk_info ? __ldg(chunk_info + blockIdx.x) : (int32_t)blockIdx.x;
  pdl_wait();
  if (sc->done) return;
  cg_pk_chunk<JACOBI>(n, res, adiag, zk, pk, sc->bet, info, cd, seq_base + (unsigned int)sc->iters + 1u);
}

// the same kernel with L2 cache-policy operands on every vector access (mask bits: 0 pk, 1 res / zk, 2 adiag); the ragged last chunk takes the plain code
template <bool JACOBI>
__global__ void __launch_bounds__(FCP_TPB) k_cg_pk_l2(int32_t n, const double *__restrict__ res, const double *__restrict__ adiag,
                                                       const double *__restrict__ zk, double *pk, const KrylovScalars *sc, const CommDev *cd,
                                                       const int32_t *__restrict__ chunk_info, unsigned int seq_base, unsigned int mask) {
  pdl_launch_dependents();
  const int32_t info = chunk_info ? __ldg(chunk_info + blockIdx.x) : (int32_t)blockIdx.x;
  pdl_wait();
  if (sc->done) return;
  const double bet = sc->bet;
  const int chunk = info & 0x7fffffff;
  const int64_t base = (int64_t)chunk * FCP_CHUNK + threadIdx.x;
  if (base + (FCP_IPT - 1) * FCP_TPB < n) {
    const L2Pol q = l2pol_make();
    const unsigned long long p_pk = l2pol_pick(q, mask, 0), p_z = l2pol_pick(q, mask, 1), p_ad = l2pol_pick(q, mask, 2);
    double z_[FCP_IPT], d_[FCP_IPT], p_[FCP_IPT];
#pragma unroll
    for (int j = 0; j < FCP_IPT; ++j) {
      const int64_t r = base + j * FCP_TPB;
      if (JACOBI) { z_[j] = l2_ld_nc(res + r, p_z); d_[j] = l2_ld_nc(adiag + r, p_ad); } else { z_[j] = l2_ld_nc(zk + r, p_z); d_[j] = 1.0; }
      p_[j] = l2_ld(pk + r, p_pk);
    }
#pragma unroll
    for (int j = 0; j < FCP_IPT; ++j) {
      const double z = JACOBI ? (z_[j] / d_[j]) : z_[j];
      l2_st(pk + base + j * FCP_TPB, z + bet * p_[j], p_pk);
    }
    if (info >= 0) return;
    const int32_t j0 = cd->chunk_ptr[chunk], j1 = cd->chunk_ptr[chunk + 1];
    __syncthreads();   // the chunk's pk values are written
    const unsigned int seq = seq_base + (unsigned int)sc->iters + 1u;
    for (int32_t j = j0 + (int32_t)threadIdx.x; j < j1; j += FCP_TPB) p2p_ll_store(cd->push_dst[j], l2_ld(pk + cd->push_cell[j], p_pk), seq);
    return;
  }
  cg_pk_chunk<JACOBI>(n, res, adiag, zk, pk, bet, info, cd, seq_base + (unsigned int)sc->iters + 1u);
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
// x / v1 loads: NC = true inside an ordinary kernel (the vectors are read-only for the kernel's lifetime: non-coherent path);
// NC = false inside the persistent solver kernel, where other CTAs rewrite them between grid barriers (plain loads: L1 is invalidated by
// the barrier's acquire fence, the non-coherent path is not)
// VM = 0: non-coherent path (__ldg); 1: coherent L1-cached load (persistent kernel); 2: non-coherent path with an L2 cache-policy operand
template <int VM>
__device__ __forceinline__ double ld_vec(const double *p, unsigned long long pol) {
  if (VM == 0) return __ldg(p);
  if (VM == 2) return l2_ld_nc(p, pol);
#ifdef FCP_EMU
  return *(const volatile double *)p;
#else
  double v;      // an L1-cached coherent load the compiler neither turns into the non-coherent path nor orders against the y stores
  asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
#endif
}

// One chunk of y = A x with the dot-product partial sums of the chunk in s[] (per thread).
template <int NS, bool SQ, int W, int VM>
__device__ __forceinline__ void spmv_dot_chunk(int32_t n, const SellView &m, const double *x, double *__restrict__ y, const double *v1, int32_t info, bool fused,
                                               const CommDev *cd, unsigned int seq, double (&s)[NS], unsigned long long px = 0ull, unsigned long long py = 0ull) {
  const int chunk = info & 0x7fffffff;
  const bool halo = fused && info < 0;
  const int64_t base = (int64_t)chunk * FCP_CHUNK + threadIdx.x;
  if (base + (FCP_IPT - 1) * FCP_TPB < n) {
    // chunks that own process faces (halo, CTA-uniform): the pipelined part covers the LOCAL entries of a row; its ghost
    // entries (stored last in the row, src-par/dpcg.f90:129-143) are added afterwards from the LL slots
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
      for (int k = 0; k < W; ++k) xv[k] = (k < len) ? ld_vec<VM>(x + c[k], px) : 0.0;
      const double vv = ld_vec<VM>(v1 + r, px);
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
      for (int32_t k = W; k < len; ++k) yr = yr + __ldcs(m.a + pos + (int64_t)k * 32) * ld_vec<VM>(x + __ldcs(m.ja + pos + (int64_t)k * 32), px);
      if (halo) {
        const int32_t full = __ldg(&m.rinfo[r]) & 0xffff;
        for (int32_t k = len; k < full; ++k) {
          const int32_t cg = __ldcs(m.ja + pos + (int64_t)k * 32);
          yr = yr + __ldcs(m.a + pos + (int64_t)k * 32) * p2p_ll_load(cd->ll + 2 * (size_t)__ldg(cd->ghost_ord + (cg - n)), seq, cd->hdr);
        }
      }
      if (VM == 2) l2_st(y + r, yr, py);
      else y[r] = yr;
      s[0] = s[0] + vv * yr;
      if (NS > 1 && SQ) s[NS - 1] = s[NS - 1] + yr * yr;
      pos = npos;
      len = nlen;
#pragma unroll
      for (int k = 0; k < W; ++k) c[k] = cn[k];
    }
  } else {
    // the ragged last chunk
    for (int j = 0; j < FCP_IPT; ++j) {
      const int64_t r64 = base + (int64_t)j * FCP_TPB;
      if (r64 >= n) continue;
      const int32_t r = (int32_t)r64;
      double yr;
      if (VM != 1) yr = halo ? sell_row_sum_halo<false>(m, x, r, 0.0, cd, seq) : sell_row_sum<false>(m, x, r, 0.0);
      else yr = halo ? sell_row_sum_halo<false, false>(m, x, r, 0.0, cd, seq) : sell_row_sum<false, false>(m, x, r, 0.0);
      y[r] = yr;
      s[0] = s[0] + v1[r] * yr;
      if (NS > 1 && SQ) s[NS - 1] = s[NS - 1] + yr * yr;
    }
  }
}
template <int NS, bool SQ, int W>
__global__ void __launch_bounds__(FCP_TPB, (W <= 8 ? FCP_PIPE_MINB : 2)) k_spmv_dot_pipe(int32_t n, SellView m, const double *__restrict__ x, double *__restrict__ y,
                                                            const double *__restrict__ v1, const KrylovScalars *sc, RedArgs ra, int fused,
                                                            unsigned int seq_base) {
  pdl_launch_dependents();
  const int32_t info = ra.chunk_info ? __ldg(ra.chunk_info + blockIdx.x) : (int32_t)blockIdx.x;
  pdl_wait();
  if (sc->done) return;
  double s[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.0;
  spmv_dot_chunk<NS, SQ, W, 0>(n, m, x, y, v1, info, fused != 0, ra.cd, seq_base + (unsigned int)sc->iters + 1u, s);
  finish_reduce_part<NS>(s, ra, info & 0x7fffffff, (int)gridDim.x);
}
// ... with L2 cache-policy operands on the vector accesses (mask bits: 0 x and v1, 1 y)
template <int NS, bool SQ, int W>
__global__ void __launch_bounds__(FCP_TPB, (W <= 8 ? FCP_PIPE_MINB : 2)) k_spmv_dot_pipe_l2(int32_t n, SellView m, const double *__restrict__ x, double *__restrict__ y,
                                                            const double *__restrict__ v1, const KrylovScalars *sc, RedArgs ra, int fused,
                                                            unsigned int seq_base, unsigned int mask) {
  pdl_launch_dependents();
  const int32_t info = ra.chunk_info ? __ldg(ra.chunk_info + blockIdx.x) : (int32_t)blockIdx.x;
  pdl_wait();
  if (sc->done) return;
  double s[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.0;
  const L2Pol q = l2pol_make();
  spmv_dot_chunk<NS, SQ, W, 2>(n, m, x, y, v1, info, fused != 0, ra.cd, seq_base + (unsigned int)sc->iters + 1u, s, l2pol_pick(q, mask, 0), l2pol_pick(q, mask, 1));
  finish_reduce_part<NS>(s, ra, info & 0x7fffffff, (int)gridDim.x);
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
__device__ __forceinline__ void cg_update_chunk(int32_t n, double *fi, double *res, const double *pk, const double *zk, const double *__restrict__ adiag,
                                                double alf, bool first, int chunk, double (&s)[3]) {
  const int64_t base = (int64_t)chunk * FCP_CHUNK + threadIdx.x;
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
    for (int j = 0; j < FCP_IPT; ++j) {
      const int64_t r = base + (int64_t)j * FCP_TPB;
      if (r >= n) continue;
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
}
template <bool JACOBI>
__global__ void __launch_bounds__(FCP_TPB) k_cg_update(int32_t n, double *__restrict__ fi, double *__restrict__ res, const double *__restrict__ pk,
                                                        const double *__restrict__ zk, const double *__restrict__ adiag,
                                                        const KrylovScalars *sc, RedArgs ra) {
  pdl_launch_dependents();
  pdl_wait();
  if (sc->done) return;
  double s[3] = {0.0, 0.0, 0.0};
  cg_update_chunk<JACOBI>(n, fi, res, pk, zk, adiag, sc->alf, sc->iters == 0, (int)blockIdx.x, s);
  finish_reduce<3>(s, ra);
}

// ... with L2 cache-policy operands (mask bits: 0 fi, 1 res, 2 pk, 3 zk, 4 adiag); ragged last chunk: the plain code
template <bool JACOBI>
__global__ void __launch_bounds__(FCP_TPB) k_cg_update_l2(int32_t n, double *fi, double *res, const double *__restrict__ pk, const double *__restrict__ zk,
                                                           const double *__restrict__ adiag, const KrylovScalars *sc, RedArgs ra, unsigned int mask) {
  pdl_launch_dependents();
  pdl_wait();
  if (sc->done) return;
  const double alf = sc->alf;
  const bool first = sc->iters == 0;
  double s[3] = {0.0, 0.0, 0.0};
  const int64_t base = (int64_t)blockIdx.x * FCP_CHUNK + threadIdx.x;
  if (base + (FCP_IPT - 1) * FCP_TPB < n) {
    const L2Pol q = l2pol_make();
    const unsigned long long p_fi = l2pol_pick(q, mask, 0), p_res = l2pol_pick(q, mask, 1), p_pk = l2pol_pick(q, mask, 2), p_zk = l2pol_pick(q, mask, 3),
                             p_ad = l2pol_pick(q, mask, 4);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      double f_[4], r_[4], p_[4], z_[4], d_[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t r = base + (h * 4 + j) * FCP_TPB;
        f_[j] = l2_ld(fi + r, p_fi); r_[j] = l2_ld(res + r, p_res); p_[j] = l2_ld_nc(pk + r, p_pk); z_[j] = l2_ld_nc(zk + r, p_zk);
        d_[j] = (JACOBI || first) ? l2_ld_nc(adiag + r, p_ad) : 1.0;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t r = base + (h * 4 + j) * FCP_TPB;
        const double f = f_[j] + alf * p_[j];
        const double rr = r_[j] - alf * z_[j];
        l2_st(fi + r, f, p_fi);
        l2_st(res + r, rr, p_res);
        s[0] = s[0] + fabs(rr);
        if (JACOBI) s[1] = s[1] + rr * (rr / d_[j]);
        if (first) s[2] = s[2] + fabs(d_[j] * f);
      }
    }
  } else {
    cg_update_chunk<JACOBI>(n, fi, res, pk, zk, adiag, alf, first, (int)blockIdx.x, s);
  }
  finish_reduce<3>(s, ra);
}

// ---------------------------------------------------------------------------------------------
// Residual halo (peer-memory path of DPCG, opt-in: FCP_HALO=res).  Measured on 8 B200s (profiles/r02_scaling.txt): 100.1 ms per step against 90.0 ms for the
// direction-vector push above -- the flagged loads of the ghost residuals cost k_cg_pk more (19.7 -> 26.8 us) than its remote stores did, and the SpMV does
// not get faster without its flagged loads (53.5 -> 53.2 us).  Kept with its parity tests as a measured alternative.  `call exchange(pk)` (src-par/dpcg.f90:118) needs the
// direction vector of the cells across every process face.  Instead of pushing pk from k_cg_pk -- whose completion then waits for its NVLink stores
// to be acknowledged, and whose values the SpMV has to fetch through flagged loads -- every rank carries the recurrence of its GHOST entries itself:
//     pk(ghost) = res(ghost) / a_ii(ghost) + bet * pk(ghost)
// is the very expression its owner evaluates, with the same operands (bet is global, the diagonal is exchanged once per solve, the residual arrives
// as LL words), hence the same bits.  The residual is pushed by k_cg_update, where the stores travel while the kernel's reduction and the
// cross-rank sum are under way; k_cg_pk has no remote stores and the SpMV reads ghost columns like any other column.
// Tags: the residual produced by iteration k (k = 0: k_cg_init) carries seq_base + k + 1 and is consumed by k_cg_pk of iteration k + 1.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FCP_TPB) k_push_ll(int32_t npro, const int32_t *__restrict__ push_cell, unsigned long long *const *__restrict__ push_dst,
                                                      const double *__restrict__ x, unsigned int seq) {
  for (int32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < npro; j += gridDim.x * blockDim.x) p2p_ll_store(push_dst[j], x[push_cell[j]], seq);
}
__global__ void __launch_bounds__(FCP_TPB) k_cg_pk_rh(int32_t n, const double *__restrict__ res, const double *__restrict__ adiag, double *pk,
                                                       const KrylovScalars *sc, const CommDev *cd, const int32_t *__restrict__ chunk_info,
                                                       unsigned int seq_base) {
  pdl_launch_dependents();
  const int32_t info = __ldg(chunk_info + blockIdx.x);
  pdl_wait();
  if (sc->done) return;
  const double bet = sc->bet;
  cg_pk_chunk<true>(n, res, adiag, nullptr, pk, bet, info & 0x7fffffff, nullptr, 0u);     // (no halo bit: no push)
  if (info >= 0) return;
  // ghost entries behind the process faces of this chunk
  const int chunk = info & 0x7fffffff;
  const int32_t j0 = cd->chunk_ptr[chunk], j1 = cd->chunk_ptr[chunk + 1];
  const unsigned int seq = seq_base + (unsigned int)sc->iters + 1u;
  for (int32_t j = j0 + (int32_t)threadIdx.x; j < j1; j += FCP_TPB) {
    const int32_t i = cd->push_ord[j], g = cd->slot[i];
    const double rg = p2p_ll_load(cd->ll + 2 * (size_t)i, seq, cd->hdr);
    pk[g] = rg / adiag[g] + bet * pk[g];
  }
}
__global__ void __launch_bounds__(FCP_TPB) k_cg_update_rh(int32_t n, double *__restrict__ fi, double *res, const double *__restrict__ pk,
                                                           const double *__restrict__ zk, const double *__restrict__ adiag, const KrylovScalars *sc,
                                                           RedArgs ra, unsigned int seq_base) {
  pdl_launch_dependents();
  const int32_t info = __ldg(ra.chunk_info + blockIdx.x);
  pdl_wait();
  if (sc->done) return;
  const int chunk = info & 0x7fffffff;
  double s[3] = {0.0, 0.0, 0.0};
  cg_update_chunk<true>(n, fi, res, pk, zk, adiag, sc->alf, sc->iters == 0, chunk, s);
  if (info < 0) {       // this chunk owns process faces: its fresh residuals go to the neighbours now, the reduction below hides the flight
    const CommDev *cd = ra.cd;
    const int32_t j0 = cd->chunk_ptr[chunk], j1 = cd->chunk_ptr[chunk + 1];
    __syncthreads();
    const unsigned int seq = seq_base + (unsigned int)sc->iters + 2u;
    for (int32_t j = j0 + (int32_t)threadIdx.x; j < j1; j += FCP_TPB) p2p_ll_store(cd->push_dst[j], res[cd->push_cell[j]], seq);
  }
  finish_reduce_part<3>(s, ra, chunk, (int)gridDim.x);
}

// ---------------------------------------------------------------------------------------------
// k_dpcg_persist: the WHOLE diagonal-PCG solve (linear_solvers.f90:280-358 / src-par/dpcg.f90:79-161) as ONE persistent cooperative kernel.
// A CTA owns the chunks q = blockIdx.x, blockIdx.x + gridDim.x, ... of the launch order (chunks with process faces first) in all three phases of an
// iteration -- {pk = zk + bet pk [+ halo push]} | {zk = A pk, pk.zk} | {fi, res update, sum|res|, next res.z} -- which run the same per-chunk code
// as the three kernels above and store the same per-chunk partial sums; the phases are separated by grid barriers built from one arrival
// counter and one generation word, and the barrier after a reducing phase is also the reduction: the LAST CTA to arrive adds the chunk
// partials in the fixed order of reduce.cuh, runs the cross-rank sum over the peer windows and the scalar epilogue (alpha / beta / convergence
// test), then releases the generation.  Bits and iteration counts are those of the three-kernel path; what disappears is three launches, three
// grid ramp-ups/tails and the host's poll per 16 iterations: the host launches once and reads the scalars once.
// Visibility: a CTA's stores are ordered before its arrival by __syncthreads + __threadfence (cumulative), the waiter's ld.acquire + fence
// invalidates its SM's L1, so the plain loads of the next phase see them; the vectors are therefore never read through the non-coherent path here.
// ---------------------------------------------------------------------------------------------
struct PersistArgs {
  int32_t n, nchunks;
  SellView m;
  double *fi, *res, *pk, *zk;
  const double *adiag;
  KrylovScalars *sc;
  RedArgs ra;                 // partials / stride / sc / cd / chunk_info (counter unused)
  unsigned int seq_base;
  int fused;                  // halo chunks read ghost columns from the LL slots
  unsigned int *bar;          // [0] arrivals (monotonic), [1] generation
  unsigned long long *phase_ns;   // nullptr, or [4]: accumulated ns of phase A, B, C as seen by CTA 0 (+ iterations) for the profiler
};
__device__ __forceinline__ unsigned int bar_ld_acquire(const unsigned int *p) {
#ifdef FCP_EMU
  emu::yield(); emu::os_yield();
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void bar_st_release(unsigned int *p, unsigned int v) {
#ifdef FCP_EMU
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
// arrival: returns the generation this barrier completes; *last = this CTA arrived last (all threads get both)
__device__ __forceinline__ unsigned int gbar_arrive(unsigned int *bar, bool *last) {
  __shared__ unsigned int ticket_s;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    ticket_s = atomicAdd(&bar[0], 1u);
  }
  __syncthreads();
  const unsigned int t = ticket_s;
  *last = (t % gridDim.x) == gridDim.x - 1u;
  return t / gridDim.x + 1u;
}
__device__ __forceinline__ void gbar_release(unsigned int *bar, unsigned int gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    bar_st_release(&bar[1], gen);
  }
}
__device__ __forceinline__ void gbar_wait(unsigned int *bar, unsigned int gen) {
  if (threadIdx.x == 0) {
    while ((int)(bar_ld_acquire(&bar[1]) - gen) < 0) {}
    __threadfence();
  }
  __syncthreads();
}
// plain barrier
__device__ __forceinline__ void gbar_sync(unsigned int *bar) {
  bool last;
  const unsigned int gen = gbar_arrive(bar, &last);
  if (last) gbar_release(bar, gen);
  else gbar_wait(bar, gen);
}
// barrier + reduction + epilogue: every CTA has stored its chunk partials
template <int NS>
__device__ __forceinline__ void gbar_reduce(unsigned int *bar, const RedArgs &ra, int nparts) {
  bool last;
  const unsigned int gen = gbar_arrive(bar, &last);
  if (last) {
    __threadfence();
    double acc[NS], total[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      double t = 0.0;
      for (int i = threadIdx.x; i < nparts; i += FCP_TPB) t = t + __ldcg(&ra.partials[(size_t)k * ra.stride + i]);
      acc[k] = t;
    }
    fcp_block_tree<NS>(acc, total);
    finish_epilogue<NS>(total, ra);
    if (threadIdx.x == 0 && ra.cd && *(volatile int *)&ra.cd->hdr->error) *(volatile int32_t *)&ra.sc->done = 1;   // a peer stopped answering: end the solve, the host reports it
    gbar_release(bar, gen);
  } else {
    gbar_wait(bar, gen);
  }
}
template <int NS>
__device__ __forceinline__ void store_chunk_partials(double (&s)[NS], const RedArgs &ra, int chunk) {
  double blk[NS];
  fcp_block_tree<NS>(s, blk);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) ra.partials[(size_t)k * ra.stride + chunk] = blk[k];
  }
}
// MINB = resident CTAs per SM the register allocation is made for (2: 128 registers, no spills; 3: 85; 4: 64 like the stand-alone SpMV kernel)
template <int W, int MINB>
__global__ void __launch_bounds__(FCP_TPB, MINB) k_dpcg_persist(PersistArgs g) {
  volatile KrylovScalars *vsc = g.sc;
  if (vsc->done) return;      // res0 < tol_abs or itr_max <= 0 (set by the init kernel's epilogue, a kernel boundary ago)
  const int32_t *__restrict__ order = g.ra.chunk_info;
  const bool timer = g.phase_ns && blockIdx.x == 0 && threadIdx.x == 0;
  unsigned long long t0 = timer ? p2p_now_ns() : 0ull;
  for (;;) {
    const unsigned int seq = g.seq_base + (unsigned int)vsc->iters + 1u;
    // ---- phase A: direction vector (+ halo push)
    const double bet = vsc->bet;
    for (int q = blockIdx.x; q < g.nchunks; q += gridDim.x) {
      const int32_t info = order ? __ldg(order + q) : (int32_t)q;
      cg_pk_chunk<true>(g.n, g.res, g.adiag, g.zk, g.pk, bet, info, g.ra.cd, seq);
    }
    gbar_sync(g.bar);
    unsigned long long t1 = 0ull;
    if (timer) { t1 = p2p_now_ns(); g.phase_ns[0] += t1 - t0; }
    // ---- phase B: zk = A pk, pk.zk
    for (int q = blockIdx.x; q < g.nchunks; q += gridDim.x) {
      const int32_t info = order ? __ldg(order + q) : (int32_t)q;
      double s[1] = {0.0};
      spmv_dot_chunk<1, false, W, 1>(g.n, g.m, g.pk, g.zk, g.pk, info, g.fused != 0, g.ra.cd, seq, s);
      store_chunk_partials<1>(s, g.ra, info & 0x7fffffff);
    }
    {
      RedArgs ra = g.ra;
      ra.epi = EPI_PKAPK;
      gbar_reduce<1>(g.bar, ra, g.nchunks);
    }
    unsigned long long t2 = 0ull;
    if (timer) { t2 = p2p_now_ns(); g.phase_ns[1] += t2 - t1; }
    // ---- phase C: solution / residual update, sum|res|, next res.z, first-iteration normalisation factor
    const double alf = vsc->alf;
    const bool first = vsc->iters == 0;
    for (int q = blockIdx.x; q < g.nchunks; q += gridDim.x) {
      const int32_t info = order ? __ldg(order + q) : (int32_t)q;
      double s[3] = {0.0, 0.0, 0.0};
      cg_update_chunk<true>(g.n, g.fi, g.res, g.pk, g.zk, g.adiag, alf, first, info & 0x7fffffff, s);
      store_chunk_partials<3>(s, g.ra, info & 0x7fffffff);
    }
    {
      RedArgs ra = g.ra;
      ra.epi = EPI_CG_UPDATE;
      gbar_reduce<3>(g.bar, ra, g.nchunks);
    }
    if (timer) { t0 = p2p_now_ns(); g.phase_ns[2] += t0 - t2; g.phase_ns[3] += 1ull; }
    if (vsc->done) break;
  }
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
// Flag-in-data sweeps (default, FCP_SWEEP=ll): the IC(0)/ILU(0) factor and the two triangular sweeps without ANY grid barrier and without
// acquire/release traffic.  A result travels as two 8-byte words {32 data bits | 32-bit sequence number} (p2p.cuh: the LL words of the halo
// protocol, here between the SMs of one GPU): an 8-byte store is atomic, so a word whose tag equals the sweep's sequence number carries valid data --
// the consumer needs no fence and the producer no release.  Rows are taken in level order, 32 consecutive positions (a tile) per warp, tiles dealt
// round-robin to the co-resident warps (cooperative launch); every level starts on a warp boundary, so a dependency always lies in an EARLIER tile:
// the lowest unfinished tile never waits and the sweep cannot deadlock.  Forward tiles are followed directly by the backward tiles (a backward row
// waits for its own forward value like for any other dependency): one kernel per apply, no barrier in between.  The triangles are read from
// the level-tile copy (TriTiles: full 256-/128-byte lines per warp although the rows of a level are scattered over the SELL slices).
// Per-row arithmetic and summation order are those of the sequential reference sweep (linear_solvers.f90:439-445, 458-475, 613-624): same bits.
// A wait gives up after ~2 s and raises sc->pad (krylov_solve turns it into an error) instead of hanging the GPU.
// ---------------------------------------------------------------------------------------------
struct TriView {
  const int64_t *tptr;
  const int32_t *tcol;
  const double *tval, *ttval;
  const int32_t *prow;
  double *dtile;
  int32_t ntiles;
};
static TriView tri_view(const TriTiles &t) { return TriView{t.tptr, t.tcol, t.tval, t.ttval, t.prow, t.dtile, t.ntiles}; }

// LL words between the SMs of ONE GPU: relaxed accesses at GPU scope.  (p2p.cuh's words travel between GPUs and are volatile = strong at SYSTEM
// scope; measured on a B200, system-scope polls and stores cost ~3.5 us per dependency hop inside a sweep.)
__device__ __forceinline__ void gpu_ll_load_words(const unsigned long long *src, unsigned long long &w0, unsigned long long &w1) {
#ifdef FCP_EMU
  p2p_ll_load_words(src, w0, w1);
#else
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void gpu_ll_store(unsigned long long *dst /* 16-byte aligned */, double v, unsigned int seq) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (bits & 0xffffffffull) | ((unsigned long long)seq << 32);
  const unsigned long long w1 = (bits >> 32) | ((unsigned long long)seq << 32);
#ifdef FCP_EMU
  p2p_ll_store_words(dst, w0, w1);
#else
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
#endif
}
__device__ __forceinline__ double sweep_ll_wait(const unsigned long long *src, unsigned int seq, const KrylovScalars *sc) {
  unsigned long long w0, w1;
#ifdef FCP_EMU
  (void)sc;
  for (;;) {
    gpu_ll_load_words(src, w0, w1);
    if ((unsigned int)(w0 >> 32) == seq && (unsigned int)(w1 >> 32) == seq) break;
    emu::yield(); emu::os_yield();
  }
#else
  unsigned int spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    gpu_ll_load_words(src, w0, w1);
    if ((unsigned int)(w0 >> 32) == seq && (unsigned int)(w1 >> 32) == seq) break;
    if ((++spins & 4095u) == 0u) {
      if (*(volatile const int32_t *)&sc->pad) break;
      const unsigned long long t = p2p_now_ns();
      if (!t0) t0 = t;
      else if (t - t0 > 2000000000ull) { *(volatile int32_t *)&const_cast<KrylovScalars *>(sc)->pad = 1; break; }
    }
  }
#endif
  return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}

// Sentinel wait.  Thousands of warps spinning on their own 32 x (up to 3) LL words saturate the L2 (measured on a B200: 69 % L2 throughput, all of
// it polls, 6 us per level).  So a warp first parks ONE lane on ONE word -- the last dependency of its last active row, the one most likely to
// arrive last -- and polls that with a pause in between; only when it has arrived do all lanes look at their own words (and spin for the few
// stragglers).  One sector per poll per warp instead of ~96.  All 32 lanes of the warp must call this.
__device__ __forceinline__ void sweep_ll_sentinel(const unsigned long long *src, int32_t col, unsigned int seq, const KrylovScalars *sc, unsigned int pause_ns) {
  const unsigned int have = __ballot_sync(0xffffffffu, col >= 0);
  if (have == 0u) return;
  const int lead = 31 - __clz((int)have);
  if ((int)(threadIdx.x & 31) == lead) {
#ifdef FCP_EMU
    (void)pause_ns;
    (void)sweep_ll_wait(src + 2 * (size_t)col, seq, sc);
#else
    unsigned long long w0, w1, t0 = 0;
    unsigned int spins = 0;
    for (;;) {
      gpu_ll_load_words(src + 2 * (size_t)col, w0, w1);
      if ((unsigned int)(w0 >> 32) == seq && (unsigned int)(w1 >> 32) == seq) break;
      if (pause_ns) __nanosleep(pause_ns);
      if ((++spins & 1023u) == 0u) {
        if (*(volatile const int32_t *)&sc->pad) break;
        const unsigned long long t = p2p_now_ns();
        if (!t0) t0 = t;
        else if (t - t0 > 2000000000ull) { *(volatile int32_t *)&const_cast<KrylovScalars *>(sc)->pad = 1; break; }
      }
    }
#endif
  }
  __syncwarp();
}

// values of this solve's matrix into the tile order (both triangles; the transposed entries for ILU(0))
__global__ void __launch_bounds__(FCP_TPB) k_tri_values(int64_t nent, const int32_t *__restrict__ src, const double *__restrict__ a, double *__restrict__ out) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nent; q += (int64_t)gridDim.x * blockDim.x) {
    const int32_t sp = __ldg(src + q);
    out[q] = sp >= 0 ? __ldg(a + sp) : 0.0;
  }
}
// d in the order of the backward tiles
__global__ void __launch_bounds__(FCP_TPB) k_tri_dtile(int32_t np, const int32_t *__restrict__ prow, const double *__restrict__ d, double *__restrict__ dtile) {
  for (int32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < np; q += gridDim.x * blockDim.x) {
    const int32_t i = __ldg(prow + q);
    dtile[q] = i >= 0 ? d[i] : 0.0;
  }
}

// d(i) = 1/(a_ii - sum_{k<diag} a_k^2 d(ja_k))               iccg     :439-445
// d(i) = 1/(a_ii - sum_{k<diag} a_k d(ja_k) a_T(k))          bicgstab :613-624
template <bool ILU>
__global__ void __launch_bounds__(FCP_TPB) k_factor_ll(TriView fw, const double *__restrict__ adiag, unsigned long long *dll, unsigned int seq, double *d,
                                                        const KrylovScalars *sc, unsigned int pause_ns) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp; t < fw.ntiles; t += nwarps) {
    const int32_t i = __ldg(fw.prow + t * 32 + lane);
    const int64_t b = __ldg(fw.tptr + t) + lane;
    const int32_t len = (int32_t)((__ldg(fw.tptr + t + 1) - __ldg(fw.tptr + t)) >> 5);
    {   // the row's last dependency (entries are packed from k = 0, columns ascending)
      int32_t clast = -1;
      if (i >= 0)
        for (int32_t k = len - 1; k >= 0 && clast < 0; --k) clast = __ldg(fw.tcol + b + (int64_t)k * 32);
      sweep_ll_sentinel(dll, clast, seq, sc, pause_ns);
    }
    if (i >= 0) {
      double di = adiag[i];
      for (int32_t k = 0; k < len; ++k) {
        const int32_t c = __ldg(fw.tcol + b + (int64_t)k * 32);
        if (c < 0) break;            // entries of a row are packed from k = 0
        const double ak = __ldg(fw.tval + b + (int64_t)k * 32);
        const double dj = sweep_ll_wait(dll + 2 * (size_t)c, seq, sc);
        if (ILU) di = di - ak * dj * __ldg(fw.ttval + b + (int64_t)k * 32);
        else di = di - ak * ak * dj;
      }
      const double dd = 1.0 / di;
      d[i] = dd;
      fw.dtile[t * 32 + lane] = dd;
      gpu_ll_store(dll + 2 * (size_t)i, dd, seq);
    }
  }
}

// zk = M^-1 rhs : forward sweep, zk/(d+small), backward sweep      (:458-475 ; quirk Q4 kept)
// The static data of a tile comes in two dependent stages -- A: row index, tile pointer, length; B: the entries, d and the right-hand side.
// PF = true software-pipelines them: while tile j is being processed (which is mostly waiting for dependencies) stage B of the warp's next tile and
// stage A of the one after are already in flight, so a tile that is due finds everything but its dependencies in registers.
struct TileA { int32_t i, len; int64_t b; };
template <int W> struct TileB { int32_t i, len; int64_t b; int32_t c[W]; double av[W], di, r0; };
__device__ __forceinline__ TileA ll_tile_a(const TriView &fw, const TriView &bw, int64_t tq, int64_t ntot, int lane) {
  TileA a;
  a.i = -1; a.len = 0; a.b = 0;
  if (tq < ntot) {
    const bool fwd = tq < fw.ntiles;
    const TriView &v = fwd ? fw : bw;
    const int64_t t = fwd ? tq : tq - fw.ntiles;
    a.i = __ldg(v.prow + t * 32 + lane);
    const int64_t b0 = __ldg(v.tptr + t);
    a.b = b0 + lane;
    a.len = (int32_t)((__ldg(v.tptr + t + 1) - b0) >> 5);
  }
  return a;
}
template <int W>
__device__ __forceinline__ TileB<W> ll_tile_b(const TriView &fw, const TriView &bw, const TileA &a, int64_t tq, int64_t ntot, int lane,
                                               const double *__restrict__ rhs) {
  TileB<W> q;
  q.i = a.i; q.len = a.len; q.b = a.b; q.di = 0.0; q.r0 = 0.0;
#pragma unroll
  for (int k = 0; k < W; ++k) { q.c[k] = -1; q.av[k] = 0.0; }
  if (tq < ntot) {
    const bool fwd = tq < fw.ntiles;
    const TriView &v = fwd ? fw : bw;
    const int64_t t = fwd ? tq : tq - fw.ntiles;
#pragma unroll
    for (int k = 0; k < W; ++k) q.c[k] = (a.i >= 0 && k < a.len) ? __ldg(v.tcol + a.b + (int64_t)k * 32) : -1;
#pragma unroll
    for (int k = 0; k < W; ++k) q.av[k] = (q.c[k] >= 0) ? __ldg(v.tval + a.b + (int64_t)k * 32) : 0.0;
    q.di = a.i >= 0 ? v.dtile[t * 32 + lane] : 0.0;
    q.r0 = (fwd && a.i >= 0) ? rhs[a.i] : 0.0;
  }
  return q;
}
template <int W>
__device__ __forceinline__ void ll_tile_process(const TriView &fw, const TriView &bw, const TileB<W> &q, int64_t tq, unsigned long long *zf,
                                                unsigned long long *zb, unsigned int seq, double *__restrict__ zk, const KrylovScalars *sc,
                                                unsigned int pause_ns) {
  const bool fwd = tq < fw.ntiles;
  const TriView &v = fwd ? fw : bw;
  const int32_t i = q.i;
  const unsigned long long *src = fwd ? zf : zb;
  // the first look at the dependencies: all loads independent of each other
  unsigned long long w0[W], w1[W];
#pragma unroll
  for (int k = 0; k < W; ++k)
    if (q.c[k] >= 0) gpu_ll_load_words(src + 2 * (size_t)q.c[k], w0[k], w1[k]);
  // pending words: bit k = dependency k of the register window, bit W = the row's own forward value (backward sweep)
  unsigned int pend = 0u;
  int32_t clast = -1;       // the last dependency inside the register window that has not arrived yet
#pragma unroll
  for (int k = 0; k < W; ++k)
    if (q.c[k] >= 0 && !((unsigned int)(w0[k] >> 32) == seq && (unsigned int)(w1[k] >> 32) == seq)) { pend |= 1u << k; clast = q.c[k]; }
  unsigned long long f0 = 0ull, f1 = 0ull;
  if (!fwd && i >= 0) {
    gpu_ll_load_words(zf + 2 * (size_t)i, f0, f1);
    if (!((unsigned int)(f0 >> 32) == seq && (unsigned int)(f1 >> 32) == seq)) pend |= 1u << W;
  }
  if (__any_sync(0xffffffffu, pend != 0u)) {          // (warp-uniform branch)
    // one lane parks on one word first (see sweep_ll_sentinel), then every lane polls ALL its pending words per round trip -- never one after the other
    const bool own_only = (pend >> W) != 0u && (pend & ((1u << W) - 1u)) == 0u;
    sweep_ll_sentinel(own_only ? zf : src, own_only ? i : clast, seq, sc, pause_ns);
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (pend) {
#pragma unroll
      for (int k = 0; k < W; ++k)
        if ((pend >> k) & 1u) gpu_ll_load_words(src + 2 * (size_t)q.c[k], w0[k], w1[k]);
      if ((pend >> W) & 1u) gpu_ll_load_words(zf + 2 * (size_t)i, f0, f1);
#pragma unroll
      for (int k = 0; k < W; ++k)
        if (((pend >> k) & 1u) && (unsigned int)(w0[k] >> 32) == seq && (unsigned int)(w1[k] >> 32) == seq) pend &= ~(1u << k);
      if (((pend >> W) & 1u) && (unsigned int)(f0 >> 32) == seq && (unsigned int)(f1 >> 32) == seq) pend &= ~(1u << W);
#ifdef FCP_EMU
      if (pend) { emu::yield(); emu::os_yield(); }
#else
      if (pend && (++spins & 4095u) == 0u) {
        if (*(volatile const int32_t *)&sc->pad) break;
        const unsigned long long t = p2p_now_ns();
        if (!t0) t0 = t;
        else if (t - t0 > 2000000000ull) { *(volatile int32_t *)&const_cast<KrylovScalars *>(sc)->pad = 1; break; }
      }
#endif
    }
  }
  if (i < 0) return;
  double z;
  if (fwd) z = q.r0;
  else z = __longlong_as_double((long long)((f0 & 0xffffffffull) | (f1 << 32))) / (q.di + FCP_SMALL);
#pragma unroll
  for (int k = 0; k < W; ++k) {
    if (q.c[k] < 0) continue;
    const double zj = __longlong_as_double((long long)((w0[k] & 0xffffffffull) | (w1[k] << 32)));
    z = z - q.av[k] * zj;
  }
  for (int32_t k = W; k < q.len; ++k) {
    const int32_t cc = __ldg(v.tcol + q.b + (int64_t)k * 32);
    if (cc < 0) break;
    z = z - __ldg(v.tval + q.b + (int64_t)k * 32) * sweep_ll_wait(src + 2 * (size_t)cc, seq, sc);
  }
  z = z * q.di;
  if (fwd) {
    gpu_ll_store(zf + 2 * (size_t)i, z, seq);
  } else {
    gpu_ll_store(zb + 2 * (size_t)i, z, seq);
    zk[i] = z;
  }
}
template <int W, bool PF>
__global__ void __launch_bounds__(FCP_TPB) k_precond_apply_ll(TriView fw, TriView bw, unsigned long long *zf, unsigned long long *zb, unsigned int seq,
                                                               const double *__restrict__ rhs, double *__restrict__ zk, const KrylovScalars *sc,
                                                               unsigned int pause_ns) {
  if (sc->done) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t ntot = (int64_t)fw.ntiles + bw.ntiles;
  if (!PF) {
    for (int64_t tq = warp; tq < ntot; tq += nwarps) {
      const TileA a = ll_tile_a(fw, bw, tq, ntot, lane);
      const TileB<W> q = ll_tile_b<W>(fw, bw, a, tq, ntot, lane, rhs);
      ll_tile_process<W>(fw, bw, q, tq, zf, zb, seq, zk, sc, pause_ns);
    }
    return;
  }
  TileB<W> cur = ll_tile_b<W>(fw, bw, ll_tile_a(fw, bw, warp, ntot, lane), warp, ntot, lane, rhs);
  TileA na = ll_tile_a(fw, bw, warp + nwarps, ntot, lane);
#pragma unroll 1
  for (int64_t tq = warp; tq < ntot; tq += nwarps) {
    const TileB<W> nb = ll_tile_b<W>(fw, bw, na, tq + nwarps, ntot, lane, rhs);      // in flight while this tile waits for its dependencies
    na = ll_tile_a(fw, bw, tq + 2 * nwarps, ntot, lane);
    ll_tile_process<W>(fw, bw, cur, tq, zf, zb, seq, zk, sc, pause_ns);
    cur = nb;
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

// FCP_SWEEP = ll (default: flag-in-data sweeps, no barrier) | barrier (one cooperative grid barrier per level) | flags (per-row ready flags with
// acquire/release, one barrier between the sweeps).  Read per call: tests and A/B timings switch it inside one process.
enum { SWEEP_LL = 0, SWEEP_BARRIER, SWEEP_FLAGS };
// the sweep switches are read ONCE PER SOLVE (krylov_solve -> sweep_cfg_refresh), not per preconditioner apply
struct SweepCfg { int mode = SWEEP_LL; bool pf = true; unsigned int pause_ns = 0; int ctas = 0; };
static thread_local SweepCfg g_sweep;
static void sweep_cfg_refresh() {
  SweepCfg c;
  const char *e = getenv("FCP_SWEEP");
  c.mode = (e && !strcmp(e, "barrier")) ? SWEEP_BARRIER : (e && !strcmp(e, "flags")) ? SWEEP_FLAGS : SWEEP_LL;
  e = getenv("FCP_SWEEP_PF");
  c.pf = !(e && !strcmp(e, "off"));            // software prefetch of the next tile's static data (default on)
  e = getenv("FCP_SWEEP_NS");
  c.pause_ns = e ? (unsigned int)std::max(0, atoi(e)) : 0u;
  e = getenv("FCP_SWEEP_CTAS");
  c.ctas = e ? atoi(e) : 0;
  g_sweep = c;
}
static int sweep_mode() { return g_sweep.mode; }
// cooperative grid of the LL kernels: all CTAs co-resident; FCP_SWEEP_CTAS caps the CTAs per SM (fewer resident warps = fewer warps polling)
static int sweep_grid(const void *fn, KrylovWS &ws, int slot, int *grid) {
  int &cap = ws.persist_grid[slot];
  if (!cap) FCP_TRY(coop_grid(fn, ws.ws_device, &cap));
  *grid = cap;
  if (g_sweep.ctas > 0) {
    const int v = g_sweep.ctas;
    int nsm = 0;
    FCP_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ws.ws_device));
    if (v > 0) *grid = std::min(cap, v * std::max(nsm, 1));
  }
  return FCP_OK;
}
// pause of the sentinel lane between two polls (FCP_SWEEP_NS, default 0: it polls back to back -- one sector per poll per warp does not load the L2)
static unsigned int sweep_pause_ns() { return g_sweep.pause_ns; }
static unsigned int next_ll_epoch(SellPattern &p, cudaStream_t st) {
  if (p.ll_epoch >= 0xfffffff0u) {          // 32-bit tags: start over (once per ~4e9 sweeps)
    for (auto *q : p.zll) cudaMemsetAsync(q, 0, sizeof(unsigned long long) * 2 * (size_t)std::max(p.n, 1), st);
    p.ll_epoch = 0;
  }
  return ++p.ll_epoch;
}

static int launch_factor(bool ilu, SellPattern &p, const double *a, double *d, KrylovWS &ws, cudaStream_t st) {
  if (p.n == 0) return FCP_OK;
  if (sweep_mode() == SWEEP_LL) {
    FCP_TRY(sell_build_tiles(p, ilu, st));
    for (int dir = 0; dir < 2; ++dir) {          // this solve's matrix values in tile order
      const TriTiles &t = p.tri[dir];
      if (!t.nent) continue;
      const int g = (int)std::min<int64_t>((t.nent + FCP_TPB - 1) / FCP_TPB, 148 * 16);
      k_tri_values<<<g, FCP_TPB, 0, st>>>(t.nent, t.tsrc, a, t.tval);
      FCP_LAUNCHED();
      if (dir == 0 && ilu) { k_tri_values<<<g, FCP_TPB, 0, st>>>(t.nent, t.ttsrc, a, t.ttval); FCP_LAUNCHED(); }
    }
    TriView fw = tri_view(p.tri[0]);
    const double *adiag = ws.adiag;
    unsigned long long *dll = p.zll[2];
    unsigned int seq = next_ll_epoch(p, st);
    const KrylovScalars *sc = ws.sc;
    const void *fn = ilu ? (const void *)k_factor_ll<true> : (const void *)k_factor_ll<false>;
    int grid = 0;
    FCP_TRY(sweep_grid(fn, ws, ilu ? 5 : 4, &grid));
    unsigned int pause = sweep_pause_ns();
    void *args[] = {&fw, &adiag, &dll, &seq, &d, &sc, &pause};
#ifdef FCP_EMU
    (void)args;
    if (ilu) emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_factor_ll<true>, fw, adiag, dll, seq, d, sc, pause);
    else emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_factor_ll<false>, fw, adiag, dll, seq, d, sc, pause);
#else
    FCP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(FCP_TPB), args, 0, st));
#endif
    FCP_LAUNCHED();
    const TriTiles &bt = p.tri[1];
    k_tri_dtile<<<std::max(1, std::min((bt.ntiles * 32 + FCP_TPB - 1) / FCP_TPB, 148 * 16)), FCP_TPB, 0, st>>>(bt.ntiles * 32, bt.prow, d, bt.dtile);
    FCP_LAUNCHED();
    FCP_CHECK_LAUNCH();
    return FCP_OK;
  }
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
static int launch_precond(SellPattern &p, const double *a, const double *d, const double *rhs, double *zk, const KrylovScalars *sc, KrylovWS &ws,
                          cudaStream_t st) {
  if (p.n == 0) return FCP_OK;
  const int mode = sweep_mode();
  if (mode == SWEEP_LL) {
    TriView fw = tri_view(p.tri[0]), bw = tri_view(p.tri[1]);
    unsigned long long *zf = p.zll[0], *zb = p.zll[1];
    unsigned int seq = next_ll_epoch(p, st);
    const bool wide = std::max(p.tri[0].maxlen, p.tri[1].maxlen) > 4;
    const bool pf = g_sweep.pf;
    const void *fn = wide ? (pf ? (const void *)k_precond_apply_ll<8, true> : (const void *)k_precond_apply_ll<8, false>)
                          : (pf ? (const void *)k_precond_apply_ll<4, true> : (const void *)k_precond_apply_ll<4, false>);
    int grid = 0;
    FCP_TRY(sweep_grid(fn, ws, (wide ? 7 : 6), &grid));
    if (pf) { int g2 = 0; FCP_TRY(coop_grid(fn, ws.ws_device, &g2)); grid = std::min(grid, g2); }
    unsigned int pause = sweep_pause_ns();
    void *args[] = {&fw, &bw, &zf, &zb, &seq, &rhs, &zk, &sc, &pause};
#ifdef FCP_EMU
    (void)args;
    if (wide && pf) emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_precond_apply_ll<8, true>, fw, bw, zf, zb, seq, rhs, zk, sc, pause);
    else if (wide) emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_precond_apply_ll<8, false>, fw, bw, zf, zb, seq, rhs, zk, sc, pause);
    else if (pf) emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_precond_apply_ll<4, true>, fw, bw, zf, zb, seq, rhs, zk, sc, pause);
    else emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_precond_apply_ll<4, false>, fw, bw, zf, zb, seq, rhs, zk, sc, pause);
#else
    FCP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(FCP_TPB), args, 0, st));
#endif
    FCP_LAUNCHED();
    return FCP_OK;
  }
  int &grid = ws.persist_grid[8];         // (per workspace = per device; round 1 cached this per process)
  if (!grid) FCP_TRY(coop_grid((const void *)k_precond_apply, ws.ws_device, &grid));
  SellView m{p.slptr, p.rinfo, p.ja, a, p.llen};
  LevelView lv{p.lev_ptr, p.lev_rows, p.blev_ptr, p.blev_rows, p.nlevels, p.nblevels};
  const int32_t *llen = p.llen;
  const bool use_flags = mode == SWEEP_FLAGS;
  static bool announced = false;
  if (use_flags && !announced) {
    announced = true;
    fprintf(stderr, "libfcp_b200: FCP_SWEEP=flags: IC(0)/ILU(0) sweeps wait on per-row ready flags instead of a grid barrier per level\n");
  }
  if (use_flags) {
    int &fgrid = ws.persist_grid[9];
    if (!fgrid) FCP_TRY(coop_grid((const void *)k_precond_apply_flags, ws.ws_device, &fgrid));
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
                           const KrylovScalars *sc, KrylovWS &ws, cudaStream_t st) {
  if (p.n == 0) return FCP_OK;
  int &grid = ws.persist_grid[10];
  if (!grid) FCP_TRY(coop_grid((const void *)k_gs_sweep, ws.ws_device, &grid));
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
                           const RedArgs &ra, cudaStream_t st, int fused = 0, unsigned int seq_base = 0, unsigned int l2mask = 0, bool pdl = false) {
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
  } else if (mode != 0 && l2mask) {
    if (p.tile_cap <= 8 * 256) FCP_LAUNCH_PDL(pdl, (k_spmv_dot_pipe_l2<NS, SQ, 8>), grid, st, p.n, m, x, y, v1, sc, ra, fused, seq_base, l2mask);
    else FCP_LAUNCH_PDL(pdl, (k_spmv_dot_pipe_l2<NS, SQ, 16>), grid, st, p.n, m, x, y, v1, sc, ra, fused, seq_base, l2mask);
  } else if (mode != 0) {
    if (p.tile_cap <= 8 * 256) FCP_LAUNCH_PDL(pdl, (k_spmv_dot_pipe<NS, SQ, 8>), grid, st, p.n, m, x, y, v1, sc, ra, fused, seq_base);
    else FCP_LAUNCH_PDL(pdl, (k_spmv_dot_pipe<NS, SQ, 16>), grid, st, p.n, m, x, y, v1, sc, ra, fused, seq_base);
  } else {
    k_spmv_dot<NS, SQ><<<grid, FCP_TPB, 0, st>>>(p.n, m, x, y, v1, sc, ra, fused, seq_base);
  }
  FCP_LAUNCHED();
  return FCP_OK;
}

// FCP_DPCG=persist: the persistent cooperative kernel; default: the three-kernel iteration with pipelined polling.  (Measured on a B200,
// profiles/r02_dpcg_ab.txt: the persistent kernel gives the same bits and iteration count but its SpMV phase is slower than the stand-alone kernel, so
// it stays opt-in.)  Read per call so that tests and A/B timings can switch inside one process.
static bool dpcg_persist_wanted() {
  const char *e = getenv("FCP_DPCG");
  return e && !strcmp(e, "persist");
}
static int launch_dpcg_persist(const SellPattern &p, const SellView &m, double *fi, KrylovWS &ws, const RedArgs &ra, int fused, unsigned int seq_base,
                               Profiler *prof, cudaStream_t st) {
  const int nchunks = fcp_nchunks(p.n);
  const bool wide = p.tile_cap > 8 * 256;
  int minb = 3;
  if (const char *e = getenv("FCP_PERSIST_MINB")) { const int v = atoi(e); if (v >= 2 && v <= 4) minb = v; }   // measurements only
  const int variant = wide ? 0 : minb - 1;      // 0: W = 16 (long rows, 2 CTAs/SM); 1..3: W = 8 with 2, 3, 4 CTAs/SM
  const void *fn = variant == 0 ? (const void *)k_dpcg_persist<16, 2> : variant == 1 ? (const void *)k_dpcg_persist<8, 2>
                 : variant == 2 ? (const void *)k_dpcg_persist<8, 3> : (const void *)k_dpcg_persist<8, 4>;
  int &cap = ws.persist_grid[variant];
  if (!cap) FCP_TRY(coop_grid(fn, ws.ws_device, &cap));
  int grid = std::min(cap, nchunks);
  if (const char *e = getenv("FCP_PERSIST_GRID")) { const int v = atoi(e); if (v > 0) grid = std::min(grid, v); }   // measurements only
  const bool timing = prof && prof->on;
  FCP_CUDA(cudaMemsetAsync(ws.bar, 0, 4 * sizeof(unsigned int), st));
  if (timing) FCP_CUDA(cudaMemsetAsync(ws.phase_ns, 0, 8 * sizeof(unsigned long long), st));
  PersistArgs g{p.n, nchunks, m, fi, ws.res, ws.pk, ws.zk, ws.adiag, ws.sc, ra, seq_base, fused, ws.bar, timing ? ws.phase_ns : nullptr};
  void *args[] = {&g};
  size_t tok = timing ? prof->begin(FCP_K_KRYLOV_PERSIST, st) : 0;
#ifdef FCP_EMU
  (void)args; (void)fn;
  if (variant == 0) emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_dpcg_persist<16, 2>, g);
  else if (variant == 1) emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_dpcg_persist<8, 2>, g);
  else if (variant == 2) emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_dpcg_persist<8, 3>, g);
  else emu::launch_coop(emu::Cfg(grid, FCP_TPB, 0, st), k_dpcg_persist<8, 4>, g);
#else
  FCP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(FCP_TPB), args, 0, st));
#endif
  if (timing) prof->end(tok, st);
  FCP_LAUNCHED();
  if (timing) {          // book the phases under the classes of the kernels they replace
    unsigned long long h[8];
    FCP_CUDA(cudaMemcpyAsync(h, ws.phase_ns, sizeof(h), cudaMemcpyDeviceToHost, st));
    FCP_CUDA(cudaStreamSynchronize(st));
    const int cls[3] = {FCP_K_CG_PK, FCP_K_SPMV_DOT, FCP_K_CG_UPDATE};
    for (int k = 0; k < 3; ++k) { prof->total_ms[cls[k]] += 1e-6 * (double)h[k]; prof->launches[cls[k]] += (int64_t)h[3]; }
  }
  return FCP_OK;
}

// Which Krylov vectors get the evict_last L2 policy: in the order of their reuse per CG iteration (pk: 3 reads + 1 write; res: 2 + 1; zk: 1 + 1;
// fi: 1 + 1; adiag: 2 reads) while the running total stays inside the budget (FCP_L2_MB, default 96 of the 126 MB).  Nothing fits at 256^3 on one GPU
// (134 MB per vector): the plain kernels run there.
struct L2Masks { unsigned int pk = 0, spmv = 0, upd = 0; };
static L2Masks krylov_l2_masks(int32_t n, int32_t ncols) {
  L2Masks m;
  const char *e = getenv("FCP_L2");
  // Measured on a B200 (profiles/r02_l2_hints.txt): with the default L2 configuration the evict_last operands do NOT keep 84 MB of vectors
  // resident against the 176 MB matrix stream -- every kernel still runs at HBM speed and the hinted loads cost 5-10 % -- so the hints are opt-in
  // (FCP_L2=auto | all); FCP_L2=carve additionally sets the persisting-L2 carve-out (cudaLimitPersistingL2CacheSize) to the budget.
  if (!e || !strcmp(e, "off")) return m;
  double budget = 96e6;
  if (const char *b = getenv("FCP_L2_MB")) budget = 1e6 * atof(b);
  if (e && !strcmp(e, "all")) budget = 1e30;
  if (e && !strcmp(e, "carve")) {
    static double carved = 0.0;
    if (carved != budget) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)budget); carved = budget; }
  }
  double used = 0.0;
  auto fits = [&](double bytes) { if (used + bytes > budget) return false; used += bytes; return true; };
  const bool v_pk = fits(8.0 * ncols), v_res = fits(8.0 * n), v_zk = fits(8.0 * ncols), v_fi = fits(8.0 * ncols), v_ad = fits(8.0 * n);
  m.pk = (v_pk ? 1u : 0u) | ((v_res || v_zk) ? 2u : 0u) | (v_ad ? 4u : 0u);     // bit 1: the z source (res for dpcg, zk for iccg)
  m.spmv = (v_pk ? 1u : 0u) | (v_zk ? 2u : 0u);
  m.upd = (v_fi ? 1u : 0u) | (v_res ? 2u : 0u) | (v_pk ? 4u : 0u) | (v_zk ? 8u : 0u) | (v_ad ? 16u : 0u);
  return m;
}

// Convergence polling that never drains the stream: after batch b is enqueued its `done` flag is copied to a pinned slot and an event recorded;
// the host then waits for the event of batch b-1 -- batch b is already queued behind it, so the GPU runs on while the host looks at the flag.
// (Round 1 synchronised the stream after every batch of 16 iterations: 9 % of an 8-GPU step was idle time.)  When batch b-1 converged, the
// kernels of batch b return at once on the device-side flag; the final fetch_scalars() waits for them.
struct BatchPoll {
  KrylovWS &ws;
  cudaStream_t st;
  const int *err_flag;      // peer-memory path: the window's error word (a wait that timed out), polled the same way
  int posted = 0;
  int post() {
    FCP_CUDA(cudaMemcpyAsync(&ws.h_poll[posted & 1], (const char *)ws.sc + offsetof(KrylovScalars, done), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    ws.h_poll[2 + (posted & 1)] = 0;
    if (err_flag) FCP_CUDA(cudaMemcpyAsync(&ws.h_poll[2 + (posted & 1)], err_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FCP_CUDA(cudaEventRecord(ws.ev[posted & 1], st));
    ++posted;
    return FCP_OK;
  }
  int previous_done(bool *done) {
    *done = false;
    if (posted < 2) return FCP_OK;
    FCP_CUDA(cudaEventSynchronize(ws.ev[(posted - 2) & 1]));
    *done = ws.h_poll[(posted - 2) & 1] != 0;
    if (ws.h_poll[2 + ((posted - 2) & 1)]) { fcp_set_error("peer-memory protocol timeout: a rank stopped responding"); return FCP_ENCCL; }
    return FCP_OK;
  }
};

// read the device scalars (synchronises the stream)
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
  sweep_cfg_refresh();
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
  BatchPoll poll{ws, st, comm_error_flag(comm)};
  L2Masks l2 = krylov_l2_masks(n, p.ncols);
  if (cd) l2 = L2Masks();       // (the hinted kernels exist for the single-GPU / pk-halo forms only)
  const bool pdl = pdl_wanted() && (!comm || cd) && !(prof && prof->on);      // (an event between two kernels breaks the programmatic dependency anyway)

  if (solver == FCP_SOLVER_GAUSS_SEIDEL) {
    if (comm) { fcp_set_error("csrsolve: 'gauss-seidel' is a serial-tree solver (src-par has none); not available with a communicator"); return FCP_EINVAL; }
    FCP_TRY(sell_build_levels(p, st));
    if (itr_max <= 0) {                       // the DO loop :137 runs zero times: nothing is touched
      ws.h_sc->done = 1;
    } else {
      for (int it = 0; it < itr_max;) {
        for (int b = 0; b < BATCH && it < itr_max; ++b, ++it) {
          FCP_PROF(prof, FCP_K_PRECOND, st, FCP_TRY(launch_gs_sweep(p, a, rhs, fi, ws.pk, ws.res, ws.adiag, ws.sc, ws, st)));
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
    const char *halo_env = getenv("FCP_HALO");
    const bool res_halo = cd && halo_env && !strcmp(halo_env, "res") && !dpcg_persist_wanted();     // opt-in: measured slower on 8 B200s (see the comment at k_push_ll)
    if (res_halo) {
      // the diagonal of the cells across the process faces, zero ghost directions, and the initial residual as LL words
      FCP_TRY(comm_exchange(ctx, ws.adiag, 1));
      if (p.ncols > n) FCP_CUDA(cudaMemsetAsync(ws.pk + n, 0, sizeof(double) * (size_t)(p.ncols - n), st));
      if (cd && ctx->npro) {
        const CommDev *hd = comm_dev_host(comm);
        k_push_ll<<<std::max(1, std::min((ctx->npro + FCP_TPB - 1) / FCP_TPB, 148)), FCP_TPB, 0, st>>>(ctx->npro, hd->push_cell, hd->push_dst, ws.res, sb + 1u);
        FCP_LAUNCHED();
      }
    }
    if (grid && (!comm || cd) && dpcg_persist_wanted()) {
      // the whole iteration loop on the device: one cooperative launch, one read of the scalars (below)
      FCP_TRY(launch_dpcg_persist(p, m, fi, ws, L.red(EPI_NONE), fw, sb, prof, st));
    } else
    for (int it = 0; it < itr_max;) {
      for (int b = 0; b < BATCH && it < itr_max; ++b, ++it) {
        if (grid && l2.pk) FCP_PROF(prof, FCP_K_CG_PK, st, (k_cg_pk_l2<true><<<grid, FCP_TPB, 0, st>>>(n, ws.res, ws.adiag, ws.zk, ws.pk, ws.sc, cd, comm_chunk_info(comm), sb, l2.pk), FCP_LAUNCHED()));
        else if (grid && res_halo) FCP_PROF(prof, FCP_K_CG_PK, st, (FCP_LAUNCH_PDL(pdl, k_cg_pk_rh, grid, st, n, ws.res, ws.adiag, ws.pk, ws.sc, cd, comm_chunk_info(comm), sb), FCP_LAUNCHED()));
        else if (grid) FCP_PROF(prof, FCP_K_CG_PK, st, (FCP_LAUNCH_PDL(pdl, k_cg_pk<true>, grid, st, n, ws.res, ws.adiag, ws.zk, ws.pk, ws.sc, cd, comm_chunk_info(comm), sb), FCP_LAUNCHED()));
        FCP_TRY(L.halo_pk());
        if (grid) FCP_PROF(prof, FCP_K_SPMV_DOT, st, FCP_TRY((launch_spmv_dot<1, false>(p, m, ws.pk, ws.zk, ws.pk, ws.sc, L.red(EPI_PKAPK), st, res_halo ? 0 : fw, sb, l2.spmv, pdl))));
        FCP_TRY(L.post(EPI_PKAPK, 1));
        if (grid && l2.upd) FCP_PROF(prof, FCP_K_CG_UPDATE, st, (k_cg_update_l2<true><<<grid, FCP_TPB, 0, st>>>(n, fi, ws.res, ws.pk, ws.zk, ws.adiag, ws.sc, L.red(EPI_CG_UPDATE), l2.upd), FCP_LAUNCHED()));
        else if (grid && res_halo) FCP_PROF(prof, FCP_K_CG_UPDATE, st, (FCP_LAUNCH_PDL(pdl, k_cg_update_rh, grid, st, n, fi, ws.res, ws.pk, ws.zk, ws.adiag, ws.sc, L.red(EPI_CG_UPDATE), sb), FCP_LAUNCHED()));
        else if (grid) FCP_PROF(prof, FCP_K_CG_UPDATE, st, (FCP_LAUNCH_PDL(pdl, k_cg_update<true>, grid, st, n, fi, ws.res, ws.pk, ws.zk, ws.adiag, ws.sc, L.red(EPI_CG_UPDATE)), FCP_LAUNCHED()));
        FCP_TRY(L.post(EPI_CG_UPDATE, 3));
      }
      FCP_CHECK_LAUNCH();
      FCP_TRY(poll.post());
      bool conv = false;
      FCP_TRY(poll.previous_done(&conv));      // (also fails when a peer stopped responding, instead of spinning through every batch)
      if (conv) break;
    }
  } else if (solver == FCP_SOLVER_ICCG) {
    FCP_TRY(sell_build_levels(p, st));
    FCP_TRY(krylov_ws_need(ws, &ws.d, (size_t)n));
    FCP_TRY(L.halo(fi));
    if (grid) { k_cg_init<false><<<grid, FCP_TPB, 0, st>>>(n, m, fi, rhs, ws.res, ws.adiag, ws.pk, L.red(EPI_INIT_CG)); FCP_LAUNCHED(); FCP_CHECK_LAUNCH(); }
    FCP_TRY(L.post(EPI_INIT_CG, 2));
    FCP_TRY(fetch_scalars(ws, st));
    if (!ws.h_sc->done) {
      FCP_TRY(launch_factor(false, p, a, ws.d, ws, st));
      for (int it = 0; it < itr_max;) {
        for (int b = 0; b < BATCH && it < itr_max; ++b, ++it) {
          FCP_PROF(prof, FCP_K_PRECOND, st, FCP_TRY(launch_precond(p, a, ws.d, ws.res, ws.zk, ws.sc, ws, st)));
          if (grid) { k_dot<<<grid, FCP_TPB, 0, st>>>(n, ws.res, ws.zk, ws.sc, L.red(EPI_SK)); FCP_LAUNCHED(); }
          FCP_TRY(L.post(EPI_SK, 1));
          if (grid && l2.pk) FCP_PROF(prof, FCP_K_CG_PK, st, (k_cg_pk_l2<false><<<grid, FCP_TPB, 0, st>>>(n, ws.res, ws.adiag, ws.zk, ws.pk, ws.sc, cd, comm_chunk_info(comm), sb, l2.pk), FCP_LAUNCHED()));
          else if (grid) FCP_PROF(prof, FCP_K_CG_PK, st, (k_cg_pk<false><<<grid, FCP_TPB, 0, st>>>(n, ws.res, ws.adiag, ws.zk, ws.pk, ws.sc, cd, comm_chunk_info(comm), sb), FCP_LAUNCHED()));
          FCP_TRY(L.halo_pk());
          if (grid) FCP_PROF(prof, FCP_K_SPMV_DOT, st, FCP_TRY((launch_spmv_dot<1, false>(p, m, ws.pk, ws.zk, ws.pk, ws.sc, L.red(EPI_PKAPK), st, fw, sb, l2.spmv))));
          FCP_TRY(L.post(EPI_PKAPK, 1));
          if (grid && l2.upd) FCP_PROF(prof, FCP_K_CG_UPDATE, st, (k_cg_update_l2<false><<<grid, FCP_TPB, 0, st>>>(n, fi, ws.res, ws.pk, ws.zk, ws.adiag, ws.sc, L.red(EPI_CG_UPDATE), l2.upd), FCP_LAUNCHED()));
          else if (grid) FCP_PROF(prof, FCP_K_CG_UPDATE, st, (k_cg_update<false><<<grid, FCP_TPB, 0, st>>>(n, fi, ws.res, ws.pk, ws.zk, ws.adiag, ws.sc, L.red(EPI_CG_UPDATE)), FCP_LAUNCHED()));
          FCP_TRY(L.post(EPI_CG_UPDATE, 3));
        }
        FCP_CHECK_LAUNCH();
        FCP_TRY(poll.post());
        bool conv = false;
        FCP_TRY(poll.previous_done(&conv));
        if (conv) break;
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
      FCP_TRY(launch_factor(true, p, a, ws.d, ws, st));
      for (int it = 0; it < itr_max;) {
        for (int b = 0; b < BATCH && it < itr_max; ++b, ++it) {
          if (grid) { k_bicg_pk<<<grid, FCP_TPB, 0, st>>>(n, ws.res, ws.uk, ws.pk, ws.sc); FCP_LAUNCHED(); }
          FCP_PROF(prof, FCP_K_PRECOND, st, FCP_TRY(launch_precond(p, a, ws.d, ws.pk, ws.zk, ws.sc, ws, st)));
          FCP_TRY(L.halo(ws.zk));
          if (grid) FCP_PROF(prof, FCP_K_SPMV_DOT, st, FCP_TRY((launch_spmv_dot<1, false>(p, m, ws.zk, ws.uk, ws.reso, ws.sc, L.red(EPI_UKRESO), st))));
          FCP_TRY(L.post(EPI_UKRESO, 1));
          if (grid) { k_bicg_half<<<grid, FCP_TPB, 0, st>>>(n, fi, ws.res, ws.zk, ws.uk, ws.sc); FCP_LAUNCHED(); }
          FCP_PROF(prof, FCP_K_PRECOND, st, FCP_TRY(launch_precond(p, a, ws.d, ws.res, ws.zk, ws.sc, ws, st)));
          FCP_TRY(L.halo(ws.zk));
          if (grid) FCP_PROF(prof, FCP_K_SPMV_DOT, st, FCP_TRY((launch_spmv_dot<2, true>(p, m, ws.zk, ws.vk, ws.res, ws.sc, L.red(EPI_VK), st))));
          FCP_TRY(L.post(EPI_VK, 2));
          if (grid) { k_bicg_update<<<grid, FCP_TPB, 0, st>>>(n, fi, ws.res, ws.zk, ws.vk, ws.reso, ws.adiag, ws.sc, L.red(EPI_BICG_UPDATE)); FCP_LAUNCHED(); }
          FCP_TRY(L.post(EPI_BICG_UPDATE, 3));
        }
        FCP_CHECK_LAUNCH();
        FCP_TRY(poll.post());
        bool conv = false;
        FCP_TRY(poll.previous_done(&conv));
        if (conv) break;
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
