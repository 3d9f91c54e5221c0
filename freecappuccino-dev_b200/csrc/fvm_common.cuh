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

__device__ __forceinline__ int64_t diag_pos(const MeshView &m, int32_t c) {
  return m.a_slptr[c >> 5] + (c & 31) + (int64_t)((m.a_rinfo[c] >> 16) & 0xffff) * 32;
}


#define FCP_GRID(n) (fcp_nchunks(n) > 0 ? fcp_nchunks(n) : 1), FCP_TPB, 0, ctx->stream   // never an empty grid (invalid configuration); the kernels guard c < n
