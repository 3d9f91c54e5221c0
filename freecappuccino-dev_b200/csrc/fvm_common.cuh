// fvm_common.cuh -- mesh view and the cell-centric gather macros shared by fvm.cu and fvm_ext.cu (see fvm.cu header)
#pragma once
#include "fcp_internal.h"

struct MeshView {
  int32_t n, F, B;
  const int64_t *slptr;
  const int32_t *len, *ent, *other, *slot;
  const double *arx, *ary, *arz, *xf, *yf, *zf, *facint, *Df;
  const double *xc, *yc, *zc, *vol;
  const int32_t *owner, *neigh;
  const int64_t *a_slptr;    // matrix SELL slice pointers
  const int32_t *a_rinfo;
  const int32_t *a_ja;       // matrix SELL column indices (0-based)
  const int32_t *a_llen;     // local (non-halo) entries per row, nullptr when there are no halo columns
  const int32_t *per_cell, *per_face, *per_slot;   // periodic pairs per boundary face (fcp_internal.h), nullptr without periodic patches
  const double *per_df;
};
static inline MeshView fcp_mesh_view(const fcp_ctx *c) {
  MeshView m;
  m.n = c->n; m.F = c->F; m.B = c->B;
  m.slptr = c->fl.slptr; m.len = c->fl.len; m.ent = c->fl.ent; m.other = c->fl.other; m.slot = c->fl.slot;
  m.arx = c->arx; m.ary = c->ary; m.arz = c->arz; m.xf = c->xf; m.yf = c->yf; m.zf = c->zf;
  m.facint = c->facint; m.Df = c->Df; m.xc = c->xc; m.yc = c->yc; m.zc = c->zc; m.vol = c->vol;
  m.owner = c->owner; m.neigh = c->neigh;
  m.per_cell = c->per_cell; m.per_face = c->per_face; m.per_slot = c->per_slot; m.per_df = c->per_df;
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
template <int W>
struct ListStage {
  int32_t v[2][1 + 3 * W][FCP_TPB];
  // issue the copies of cell c's list into stage st (c >= n: nothing to fetch, the stage reads as an empty list)
  __device__ __forceinline__ void fetch(const MeshView &m, int64_t c64, int st) {
    const int t = threadIdx.x;
    if (c64 < m.n) {
      const int32_t c = (int32_t)c64;
      const int64_t b0 = __ldg(&m.slptr[c >> 5]);
      const int32_t width = (int32_t)((__ldg(&m.slptr[(c >> 5) + 1]) - b0) >> 5);
      const int64_t fbase = b0 + (c & 31);
      fcp_cp_async4(&v[st][0][t], m.len + c);
#pragma unroll
      for (int k = 0; k < W; ++k) {
        if (k < width) {
          const int64_t pos = fbase + (int64_t)k * 32;
          fcp_cp_async4(&v[st][1 + k][t], m.ent + pos);
          fcp_cp_async4(&v[st][1 + W + k][t], m.other + pos);
          fcp_cp_async4(&v[st][1 + 2 * W + k][t], m.slot + pos);
        } else {
          v[st][1 + k][t] = 0; v[st][1 + W + k][t] = 0; v[st][1 + 2 * W + k][t] = -1;
        }
      }
    } else {
      v[st][0][t] = 0;
    }
    fcp_cp_async_commit();
  }
  __device__ __forceinline__ int32_t len(int st) const { return v[st][0][threadIdx.x]; }
  __device__ __forceinline__ void read(int st, int32_t (&e)[W], int32_t (&o)[W], int32_t (&sl)[W]) const {
    const int t = threadIdx.x;
    const int32_t flen = v[st][0][t];
#pragma unroll
    for (int k = 0; k < W; ++k) {
      const bool on = k < flen;
      e[k] = on ? v[st][1 + k][t] : 0;
      o[k] = on ? v[st][1 + W + k][t] : 0;
      sl[k] = on ? v[st][1 + 2 * W + k][t] : -1;
    }
  }
};
// a stage in DYNAMIC shared memory (ListStage<10> is 62 KB, over the 48 KB static limit): launch with sizeof(ListStage<WS>) bytes
#ifdef FCP_EMU
#define FCP_STAGE_DYN(WS, name) ListStage<WS> &name = *reinterpret_cast<ListStage<WS> *>(emu::dyn_smem())
#else
#define FCP_STAGE_DYN(WS, name)                                   \
  extern __shared__ __align__(16) unsigned char stage_raw__[];    \
  ListStage<WS> &name = *reinterpret_cast<ListStage<WS> *>(stage_raw__)
#endif
template <int WS, class K>
static int fcp_stage_smem(K kernel, size_t *bytes) {
  *bytes = sizeof(ListStage<WS>);
  if (*bytes > 48 * 1024) FCP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*bytes));
  return FCP_OK;
}

// the cell loop of FCP_CELL_LOOP with staged lists: `c` is the cell, `st` the stage that holds its list
__device__ __forceinline__ int64_t fcp_chunk_cell(int j) { return (int64_t)blockIdx.x * FCP_CHUNK + (int64_t)j * FCP_TPB + threadIdx.x; }
#define FCP_STAGED_LOOP_BEGIN(stage, m, n, c, st)                                                                    \
  (stage).fetch((m), fcp_chunk_cell(0), 0);                                                                           \
  for (int j__ = 0, st = 0; j__ < FCP_IPT; ++j__, st ^= 1) {                                                          \
    if (j__ + 1 < FCP_IPT) { (stage).fetch((m), fcp_chunk_cell(j__ + 1), st ^ 1); fcp_cp_async_wait<1>(); }          \
    else fcp_cp_async_wait<0>();                                                                                      \
    if (fcp_chunk_cell(j__) >= (n)) continue;                                                                         \
    const int32_t c = (int32_t)fcp_chunk_cell(j__);
#define FCP_STAGED_LOOP_END }

// FCP_FACE_BATCHES / FCP_BATCH_LISTS with the list taken from the stage when the cell fits it (<= WS faces), else from global memory as before
#define FCP_FACE_BATCHES_STAGED(stage, st, m, c, W)                                 \
  const int32_t flen__ = (stage).len(st);                                          \
  const bool staged__ = flen__ <= (int32_t)(sizeof((stage).v[0]) / sizeof((stage).v[0][0]) - 1) / 3;   \
  const int64_t fbase__ = staged__ ? 0 : (m).slptr[(c) >> 5] + ((c) & 31);         \
  for (int32_t q0__ = 0; q0__ < flen__; q0__ += (W))
#define FCP_BATCH_LISTS_STAGED(stage, st, WS, m, W, e, o, sl)                      \
  int32_t e[W], o[W], sl[W];                                                       \
  _Pragma("unroll") for (int k__ = 0; k__ < (W); ++k__) {                          \
    const bool on__ = q0__ + k__ < flen__;                                         \
    if (staged__) {                                                                \
      const int idx__ = on__ ? q0__ + k__ : 0;                                     \
      e[k__] = on__ ? (stage).v[st][1 + idx__][threadIdx.x] : 0;                   \
      o[k__] = on__ ? (stage).v[st][1 + (WS) + idx__][threadIdx.x] : 0;            \
      sl[k__] = on__ ? (stage).v[st][1 + 2 * (WS) + idx__][threadIdx.x] : -1;      \
    } else {                                                                       \
      const int64_t pos__ = fbase__ + (int64_t)(q0__ + k__) * 32;                  \
      e[k__] = on__ ? __ldcs((m).ent + pos__) : 0;                                 \
      o[k__] = on__ ? __ldcs((m).other + pos__) : 0;                               \
      sl[k__] = on__ ? __ldcs((m).slot + pos__) : -1;                              \
    }                                                                              \
  }

__device__ __forceinline__ int64_t diag_pos(const MeshView &m, int32_t c) {
  return m.a_slptr[c >> 5] + (c & 31) + (int64_t)((m.a_rinfo[c] >> 16) & 0xffff) * 32;
}


#define FCP_GRID(n) (fcp_nchunks(n) > 0 ? fcp_nchunks(n) : 1), FCP_TPB, 0, ctx->stream   // never an empty grid (invalid configuration); the kernels guard c < n
