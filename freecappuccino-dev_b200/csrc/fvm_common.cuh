// fvm_common.cuh -- mesh view and the cell-centric gather macros shared by fvm.cu and fvm_ext.cu (see fvm.cu header)
#pragma once
#include <cstdlib>
#include <cstring>
#include "fcp_internal.h"

struct MeshView {
  int32_t n, F, B;
  const int64_t *slptr;
  const int32_t *len, *ent, *other, *slot;
  const unsigned long long *kinds;   // compact face kinds per cell (FaceLists::kinds) when the launcher selects the compact lists, else nullptr
  const double *arx, *ary, *arz, *xf, *yf, *zf, *facint, *Df;
  const double *xc, *yc, *zc, *vol;
  const int32_t *owner, *neigh;
  const int64_t *a_slptr;    // matrix SELL slice pointers
  const int32_t *a_rinfo;
  const int32_t *a_ja;       // matrix SELL column indices (0-based)
  const int32_t *a_llen;     // local (non-halo) entries per row, nullptr when there are no halo columns
  const int32_t *per_cell, *per_face, *per_slot;   // periodic pairs per boundary face (fcp_internal.h), nullptr without periodic patches
  const double *per_df;
  const int32_t *proc_flip;  // per boundary face: 1 = process face owned by the peer's cell in the unpartitioned mesh (fcp_set_process_orientation), or nullptr
};
static inline MeshView fcp_mesh_view(const fcp_ctx *c) {
  MeshView m;
  m.n = c->n; m.F = c->F; m.B = c->B;
  m.slptr = c->fl.slptr; m.len = c->fl.len; m.ent = c->fl.ent; m.other = c->fl.other; m.slot = c->fl.slot; m.kinds = nullptr;
  m.arx = c->arx; m.ary = c->ary; m.arz = c->arz; m.xf = c->xf; m.yf = c->yf; m.zf = c->zf;
  m.facint = c->facint; m.Df = c->Df; m.xc = c->xc; m.yc = c->yc; m.zc = c->zc; m.vol = c->vol;
  m.owner = c->owner; m.neigh = c->neigh;
  m.per_cell = c->per_cell; m.per_face = c->per_face; m.per_slot = c->per_slot; m.per_df = c->per_df; m.proc_flip = c->d_proc_flip;
  m.a_slptr = c->pat.slptr; m.a_rinfo = c->pat.rinfo; m.a_ja = c->pat.ja; m.a_llen = c->pat.llen;
  return m;
}

#define FCP_CELL_LOOP(c, n)                                                                       \
  for (int j__ = 0; j__ < FCP_IPT; ++j__)                                                         \
    for (int32_t c = (int32_t)((int64_t)blockIdx.x * FCP_CHUNK + j__ * FCP_TPB + threadIdx.x), once__ = 1; \
         once__ && c < (n); once__ = 0)

// walk the faces of cell c: e = signed entry, o = index across the face, sl = matrix slot (>=0 two-sided face,
// -1-bctype for a physical boundary face), f = 0-based face index
#define FCP_FACE_LOOP(m, c)                                                        \
  const int64_t fbase__ = (m).slptr[(c) >> 5] + ((c) & 31);                        \
  const int32_t flen__ = (m).len[c];                                               \
  for (int32_t q__ = 0; q__ < flen__; ++q__)
#define FCP_FACE_FETCH(m)                                                          \
  const int32_t e = __ldcs((m).ent + fbase__ + (int64_t)q__ * 32);                 \
  const int32_t o = __ldcs((m).other + fbase__ + (int64_t)q__ * 32);               \
  const int32_t sl = __ldcs((m).slot + fbase__ + (int64_t)q__ * 32);               \
  const int32_t f = (e > 0 ? e : -e) - 1;                                          \
  (void)o; (void)sl; (void)f

// Batched walk: the list entries of W faces are loaded first (3W independent, coalesced loads), then every gather of
// those W faces is issued before the first use, so a cell costs two memory round trips instead of two per face.  The
// arithmetic still runs face by face in ascending face order.  Entries past the end of the list read as e = 0 (no face).
#define FCP_FACE_BATCHES(m, c, W)                                                   \
  const int64_t fbase__ = (m).slptr[(c) >> 5] + ((c) & 31);                        \
  const int32_t flen__ = (m).len[c];                                               \
  for (int32_t q0__ = 0; q0__ < flen__; q0__ += (W))
#define FCP_BATCH_LISTS(m, W, e, o, sl)                                            \
  int32_t e[W], o[W], sl[W];                                                       \
  _Pragma("unroll") for (int k__ = 0; k__ < (W); ++k__) {                          \
    const bool on__ = q0__ + k__ < flen__;                                         \
    const int64_t pos__ = fbase__ + (int64_t)(q0__ + k__) * 32;                    \
    e[k__] = on__ ? __ldcs((m).ent + pos__) : 0;                                   \
    o[k__] = on__ ? __ldcs((m).other + pos__) : 0;                                 \
    sl[k__] = on__ ? __ldcs((m).slot + pos__) : -1;                                \
  }

// ---------------------------------------------------------------------------------------------
// Face-list staging.  A thread walks 8 cells (stride 256); every cell costs two DEPENDENT memory round trips: its face list (ent, other, slot),
// then the gathers the list addresses.  ListStage moves the first round trip off the critical path: while cell j is being processed the list of
// cell j+1 is copied global -> shared memory by cp.async (no registers, no scoreboard entry), so that cell j+1 starts with its list already on
// chip and costs ONE round trip.  Two stages of [1 + 3 W][256] int32 (37 KB for W = 6); a thread only ever reads the slots it filled itself, so
// cp.async.wait_group is the only synchronisation.  Entries past the end of the cell's slice read as e = 0 (no face), like FCP_BATCH_LISTS.
// ---------------------------------------------------------------------------------------------
#ifdef FCP_EMU
__device__ __forceinline__ void fcp_cp_async4(int32_t *dst_smem, const int32_t *src) { *dst_smem = *src; }
__device__ __forceinline__ void fcp_cp_async8(double *dst_smem, const double *src) { *dst_smem = *src; }
__device__ __forceinline__ void fcp_cp_async_commit() {}
template <int N> __device__ __forceinline__ void fcp_cp_async_wait() {}
#else
__device__ __forceinline__ void fcp_cp_async4(int32_t *dst_smem, const int32_t *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void fcp_cp_async8(double *dst_smem, const double *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void fcp_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void fcp_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif
// The slice pointers of the (up to) FCP_IPT cells a thread walks, fetched ONCE per warp: the cells of iteration j lie in slice
// blockIdx.x * (FCP_CHUNK / 32) + j * (FCP_TPB / 32) + warp, so lane j holds slptr[that slice] and lane j + FCP_IPT slptr[that slice + 1]; an
// iteration takes them with two shuffles instead of a dependent load (an L2 round trip per cell in front of every list fetch).
struct StageSlices {
  int64_t v;
  __device__ __forceinline__ void init(const MeshView &m) {
    static_assert(2 * FCP_IPT <= 32, "one lane per slice bound");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t s = (int64_t)blockIdx.x * (FCP_CHUNK / 32) + (int64_t)(lane % FCP_IPT) * (FCP_TPB / 32) + warp + (lane / FCP_IPT);
    const int64_t nsl = ((int64_t)m.n + 31) >> 5;
    v = (lane < 2 * FCP_IPT && s <= nsl) ? __ldg(m.slptr + s) : 0;
  }
  // every lane of the warp must call this (shuffles); j = iteration of the cell loop
  __device__ __forceinline__ void get(int j, int64_t &b0, int64_t &b1) const {
    b0 = __shfl_sync(0xffffffffu, v, j);
    b1 = __shfl_sync(0xffffffffu, v, j + FCP_IPT);
  }
};
__device__ __forceinline__ void fcp_prefetch_l2(const void *p) {
#ifndef FCP_EMU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
// Compact lists (m.kinds != nullptr; kernels that need no matrix slot): `slot` only tells a gradient kernel whether a face is two-sided and, if not,
// the patch type, and `len` is a fourth stream -- both are folded into ONE 8-byte word per cell (FaceLists::kinds: nibble k = 0 for a two-sided face,
// 1 + bctype otherwise; top byte = the cell's face count, 255 = more than 14 faces: such a cell reads the plain lists).  A hexahedron then reads
// 56 bytes of list instead of 76.  The word's halves take the place of the `len` row and of the first `slot` row of the stage.
template <int W, int NS = 2>
struct ListStage {
  int32_t v[NS][1 + 3 * W][FCP_TPB];
  // issue the copies of cell c's list into stage st (c >= n: nothing to fetch, the stage reads as an empty list); b0, b1 = slice bounds (StageSlices)
  __device__ __forceinline__ void fetch(const MeshView &m, int64_t c64, int st, int64_t b0, int64_t b1) {
    const int t = threadIdx.x;
    const bool cl = m.kinds != nullptr;
    if (c64 < m.n) {
      const int32_t c = (int32_t)c64;
      const int32_t width = (int32_t)((b1 - b0) >> 5);
      const int64_t fbase = b0 + (c & 31);
      if (cl) {
        const int32_t *kw = reinterpret_cast<const int32_t *>(m.kinds + c);
        fcp_cp_async4(&v[st][0][t], kw);
        fcp_cp_async4(&v[st][1 + 2 * W][t], kw + 1);
      } else {
        fcp_cp_async4(&v[st][0][t], m.len + c);
      }
#pragma unroll
      for (int k = 0; k < W; ++k) {
        if (k < width) {
          const int64_t pos = fbase + (int64_t)k * 32;
          fcp_cp_async4(&v[st][1 + k][t], m.ent + pos);
          fcp_cp_async4(&v[st][1 + W + k][t], m.other + pos);
          if (!cl) fcp_cp_async4(&v[st][1 + 2 * W + k][t], m.slot + pos);
        } else {
          v[st][1 + k][t] = 0; v[st][1 + W + k][t] = 0;
          if (!cl) v[st][1 + 2 * W + k][t] = -1;
        }
      }
    } else {
      v[st][0][t] = 0;
      if (cl) v[st][1 + 2 * W][t] = 0;
    }
    fcp_cp_async_commit();
  }
  // plain lists
  __device__ __forceinline__ int32_t len(int st) const { return v[st][0][threadIdx.x]; }
  __device__ __forceinline__ int32_t ent(int st, int k) const { return v[st][1 + k][threadIdx.x]; }
  __device__ __forceinline__ int32_t oth(int st, int k) const { return v[st][1 + W + k][threadIdx.x]; }
  // plain or compact lists (cl = m.kinds != nullptr); len: 255 = "read m.len[c] and the plain lists"
  __device__ __forceinline__ int32_t len(int st, bool cl) const {
    return cl ? (int32_t)((uint32_t)v[st][1 + 2 * W][threadIdx.x] >> 24) : v[st][0][threadIdx.x];
  }
  __device__ __forceinline__ int32_t slot(int st, int k, bool cl) const {
    if (!cl) return v[st][1 + 2 * W + k][threadIdx.x];
    const uint32_t h = (uint32_t)(k < 8 ? v[st][0][threadIdx.x] : v[st][1 + 2 * W][threadIdx.x]);
    return -(int32_t)((h >> (4 * (k & 7))) & 15u);       // 0 = two-sided (>= 0), -1 - bctype otherwise: what the gradient kernels test
  }
  __device__ __forceinline__ void read(int st, int32_t (&e)[W], int32_t (&o)[W], int32_t (&sl)[W], bool cl = false) const {
    const int32_t flen = len(st, cl);
#pragma unroll
    for (int k = 0; k < W; ++k) {
      const bool on = k < flen;
      e[k] = on ? ent(st, k) : 0;
      o[k] = on ? oth(st, k) : 0;
      sl[k] = on ? slot(st, k, cl) : -1;
    }
  }
};
// a stage in DYNAMIC shared memory (ListStage<10> is 62 KB, over the 48 KB static limit): launch with sizeof(ListStage<WS>) bytes
#ifdef FCP_EMU
#define FCP_STAGE_DYN_N(WS, NS, name) ListStage<WS, NS> &name = *reinterpret_cast<ListStage<WS, NS> *>(emu::dyn_smem())
#else
#define FCP_STAGE_DYN_N(WS, NS, name)                             \
  extern __shared__ __align__(16) unsigned char stage_raw__[];    \
  ListStage<WS, NS> &name = *reinterpret_cast<ListStage<WS, NS> *>(stage_raw__)
#endif
#define FCP_STAGE_DYN(WS, name) FCP_STAGE_DYN_N(WS, 2, name)
template <int WS, int NS = 2, class K>
static int fcp_stage_smem(K kernel, size_t *bytes) {
  *bytes = sizeof(ListStage<WS, NS>);
  if (*bytes > 48 * 1024) FCP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*bytes));
  return FCP_OK;
}

// the cell loop of FCP_CELL_LOOP with staged lists: `c` is the cell, `st` the stage that holds its list
__device__ __forceinline__ int64_t fcp_chunk_cell(int j) { return (int64_t)blockIdx.x * FCP_CHUNK + (int64_t)j * FCP_TPB + threadIdx.x; }
// NS stages: the list of cell j + NS - 1 is in flight while cell j is processed; `stn` = the stage of cell j + 1, complete when NS >= 3 (what a
// prefetch of that cell's gathers needs), the trailing arguments = a statement executed once per iteration after the wait (before the c < n test)
#define FCP_STAGED_LOOP_BEGIN_N(NS, stage, m, n, c, st, stn, ...)                                                     \
  StageSlices slices__; slices__.init(m);                                                                             \
  _Pragma("unroll") for (int s__ = 0; s__ < (NS) - 1; ++s__) {                                                        \
    int64_t b0__, b1__; slices__.get(s__, b0__, b1__);                                                                \
    (stage).fetch((m), fcp_chunk_cell(s__), s__, b0__, b1__);                                                         \
  }                                                                                                                   \
  for (int j__ = 0, st = 0; j__ < FCP_IPT; ++j__, st = (st + 1 == (NS) ? 0 : st + 1)) {                               \
    const int stn = st + 1 == (NS) ? 0 : st + 1; (void)stn;                                                           \
    if (j__ + (NS) - 1 < FCP_IPT) {                                                                                   \
      int64_t b0__, b1__; slices__.get(j__ + (NS) - 1, b0__, b1__);                                                   \
      (stage).fetch((m), fcp_chunk_cell(j__ + (NS) - 1), (st + (NS) - 1) % (NS), b0__, b1__);                         \
      fcp_cp_async_wait<1>();                                                                                         \
    } else fcp_cp_async_wait<0>();                                                                                    \
    __VA_ARGS__;                                                                                                      \
    if (fcp_chunk_cell(j__) >= (n)) continue;                                                                         \
    const int32_t c = (int32_t)fcp_chunk_cell(j__);
#define FCP_STAGED_LOOP_BEGIN(stage, m, n, c, st) FCP_STAGED_LOOP_BEGIN_N(2, stage, m, n, c, st, stn__, (void)0)
#define FCP_STAGED_LOOP_END }

// FCP_FACE_BATCHES / FCP_BATCH_LISTS with the list taken from the stage when the cell fits it (<= WS faces), else from global memory as before
#define FCP_FACE_BATCHES_STAGED(stage, st, m, c, W)                                 \
  const bool cl__ = (m).kinds != nullptr;                                          \
  const int32_t slen__ = (stage).len(st, cl__);                                    \
  const int32_t flen__ = (cl__ && slen__ == 255) ? (m).len[c] : slen__;            \
  const bool staged__ = flen__ <= (int32_t)(sizeof((stage).v[0]) / sizeof((stage).v[0][0]) - 1) / 3;   \
  const int64_t fbase__ = staged__ ? 0 : (m).slptr[(c) >> 5] + ((c) & 31);         \
  for (int32_t q0__ = 0; q0__ < flen__; q0__ += (W))
#define FCP_BATCH_LISTS_STAGED(stage, st, WS, m, W, e, o, sl)                      \
  int32_t e[W], o[W], sl[W];                                                       \
  _Pragma("unroll") for (int k__ = 0; k__ < (W); ++k__) {                          \
    const bool on__ = q0__ + k__ < flen__;                                         \
    if (staged__) {                                                                \
      const int idx__ = on__ ? q0__ + k__ : 0;                                     \
      e[k__] = on__ ? (stage).ent(st, idx__) : 0;                                  \
      o[k__] = on__ ? (stage).oth(st, idx__) : 0;                                  \
      sl[k__] = on__ ? (stage).slot(st, idx__, cl__) : -1;                         \
    } else {                                                                       \
      const int64_t pos__ = fbase__ + (int64_t)(q0__ + k__) * 32;                  \
      e[k__] = on__ ? __ldcs((m).ent + pos__) : 0;                                 \
      o[k__] = on__ ? __ldcs((m).other + pos__) : 0;                               \
      sl[k__] = on__ ? __ldcs((m).slot + pos__) : -1;                              \
    }                                                                              \
  }

// Switches of the face kernels, read at every launch (a handful of launches per step): FCP_FACE_OCC = 2 | 3 CTAs per SM asked of the compiler,
// FCP_FACE_PF = 0 | 1 | 2 L2 prefetch of the next cell's operands (see k_grad_gauss), FCP_FACE_CL = 0 | 1 compact lists in the gradient kernels
// (ListStage), FCP_FACE_OG = 0 | 1 face geometry from the owner-ordered arrays (fcp_ctx::og).  The environment overrides the defaults, one value for every kernel or a comma-separated value per kernel (tools/face_ab.py); the defaults
// are per kernel, set from that measurement.
struct FaceVariant { int occ, pf, cl, og; };
enum { FCP_FK_GRAD_GAUSS = 0, FCP_FK_GRAD_LSQ, FCP_FK_GRADP, FCP_FK_ASSEMBLE, FCP_FK_COUNT };
static inline FaceVariant fcp_face_variant(int kernel) {
  static const FaceVariant defaults[FCP_FK_COUNT] = {
      /* k_grad_gauss     */ {2, 0, 1, 1},     // (also k_grad_gauss_fvx)
      /* k_grad_lsq       */ {2, 0, 0, 0},
      /* k_gradp          */ {2, 0, 1, 1},
      /* k_assemble_pcorr */ {2, 0, 0, 0},
  };
  FaceVariant v = defaults[kernel];
  // "1" = every kernel, "1,0,2,1" = per kernel in the order of the enum
  auto pick = [kernel](const char *e, int fallback) {
    if (!e || !*e) return fallback;
    if (!strchr(e, ',')) return atoi(e);
    for (int k = 0; k < kernel; ++k) { e = strchr(e, ','); if (!e) return fallback; ++e; }
    return atoi(e);
  };
  v.occ = pick(getenv("FCP_FACE_OCC"), v.occ) >= 3 ? 3 : 2;
  v.pf = pick(getenv("FCP_FACE_PF"), v.pf);
  if (v.pf < 0 || v.pf > 2) v.pf = 0;
  v.cl = pick(getenv("FCP_FACE_CL"), v.cl) != 0;
  v.og = pick(getenv("FCP_FACE_OG"), v.og) != 0;
  return v;
}

// compact_ok: the kernel needs no matrix slot; og_ok: the kernel uses a face index only to address arx, ary, arz, facint, xf, yf, zf
static inline int fcp_apply_face_variant(fcp_ctx *ctx, const FaceVariant &fv, MeshView &m, bool compact_ok, bool og_ok) {
  if (fv.cl && compact_ok) m.kinds = ctx->fl.kinds;
  if (fv.og && og_ok) {
    FCP_TRY(fvm_ensure_og(ctx));
    const int64_t g = ctx->og_n;
    m.ent = ctx->fl.gent;
    m.arx = ctx->og; m.ary = ctx->og + g; m.arz = ctx->og + 2 * g; m.facint = ctx->og + 3 * g;
    m.xf = ctx->og + 4 * g; m.yf = ctx->og + 5 * g; m.zf = ctx->og + 6 * g;
  }
  return FCP_OK;
}

__device__ __forceinline__ int64_t diag_pos(const MeshView &m, int32_t c) {
  return m.a_slptr[c >> 5] + (c & 31) + (int64_t)((m.a_rinfo[c] >> 16) & 0xffff) * 32;
}


#define FCP_GRID(n) (fcp_nchunks(n) > 0 ? fcp_nchunks(n) : 1), FCP_TPB, 0, ctx->stream   // never an empty grid (invalid configuration); the kernels guard c < n
