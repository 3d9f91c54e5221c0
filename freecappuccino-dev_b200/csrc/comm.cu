// comm.cu -- multi-GPU layer in the src-par layout: one rank = one mesh partition = one GPU.
//   exchange(phi)      src-par/exchange.f90:3-129     -> pack kernel + one NCCL group of send/recv + unpack kernel
//   global_sum(x)      src-par/global_sum_mpi.f90:4-37 -> all-gather of the per-rank partial sums + rank-ordered add
//                                                         (deterministic, identical on every rank)
// NCCL is resolved with dlopen at fcp_comm_init: a single-GPU host program needs no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include "fcp_internal.h"
#include "p2p.cuh"

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.handle) return FCP_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
#ifdef FCP_EMU   // tests/emu: the blocking inter-process stand-in linked into the emulation library itself (test infrastructure only)
  (void)names;
  Dl_info self;
  if (dladdr((void *)&g_nccl, &self) && self.dli_fname) h = dlopen(self.dli_fname, RTLD_NOW);
#else
  for (const char *nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
#endif
  if (!h) { fcp_set_error("cannot load libnccl.so.2: %s", dlerror()); return FCP_ENCCL; }
#define SYM(field, name)                                                             \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                        \
  if (!g_nccl.field) { fcp_set_error("libnccl: missing symbol %s", name); return FCP_ENCCL; }
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(AllGather, "ncclAllGather");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.handle = h;
  return FCP_OK;
}
#define FCP_NCCL(call)                                                                               \
  do {                                                                                               \
    ncclResult_t r__ = (call);                                                                       \
    if (r__ != ncclSuccess) {                                                                        \
      fcp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__));       \
      return FCP_ENCCL;                                                                              \
    }                                                                                                \
  } while (0)

struct FcpComm {
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  int32_t npro = 0;
  std::vector<int> peer;            // per process patch (patch order): rank on the other side
  std::vector<int32_t> off, cnt;    // per process patch: offset / count in faces inside the halo buffers
  int32_t *d_cell = nullptr;        // [npro] owner cell of each process face (my_mpi_module bufind)
  int32_t *d_slot = nullptr;        // [npro] ghost slot (numCells + boundary-face offset)
  double *sendbuf = nullptr, *recvbuf = nullptr;   // [3*npro]
  double *gather = nullptr;         // [4*nranks]
  double *d_scalar = nullptr;       // [4]
  // ---- peer-memory path (CUDA IPC over NVLink) ----
  bool p2p = false;
  void *win = nullptr;              // own window: WinHeader | staging [2][3*npro] | halo vector [numTotal]
  std::vector<void *> peer_win;     // by rank (own = win)
  CommDev h_dev;                    // host copy of the device descriptor
  CommDev *d_dev = nullptr;
  int32_t *d_frank = nullptr, *d_rord = nullptr, *d_chunk_ptr = nullptr, *d_push_cell = nullptr, *d_push_ord = nullptr, *d_ghost_ord = nullptr, *d_order = nullptr;
  unsigned long long **d_push_dst = nullptr;
  unsigned long long xseq = 0;      // sequence number of the generic halo exchanges (identical on all ranks)
  unsigned int pk_base = 0;         // sequence base of the fused direction-vector pushes (advanced after every solve)
};
int comm_nranks(const FcpComm *c) { return c ? c->nranks : 1; }
const CommDev *comm_dev(const FcpComm *c) { return (c && c->p2p) ? c->d_dev : nullptr; }
const CommDev *comm_dev_host(const FcpComm *c) { return (c && c->p2p) ? &c->h_dev : nullptr; }
const int32_t *comm_chunk_info(const FcpComm *c) { return (c && c->p2p) ? c->d_order : nullptr; }
unsigned int comm_pk_base(const FcpComm *c) { return c ? c->pk_base : 0u; }
void comm_pk_advance(FcpComm *c, int32_t iters) { if (c) c->pk_base += (unsigned int)iters + 2u; }   // the residual-halo scheme tags one push past the last iteration
void comm_free(FcpComm *c) {
  if (!c) return;
  for (size_t r = 0; r < c->peer_win.size(); ++r)
    if ((int)r != c->rank && c->peer_win[r]) cudaIpcCloseMemHandle(c->peer_win[r]);
  cudaFree(c->win); cudaFree(c->d_dev); cudaFree(c->d_frank); cudaFree(c->d_rord); cudaFree(c->d_chunk_ptr); cudaFree(c->d_push_cell); cudaFree(c->d_push_ord);
  cudaFree(c->d_ghost_ord); cudaFree(c->d_order); cudaFree(c->d_push_dst);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  cudaFree(c->d_cell); cudaFree(c->d_slot); cudaFree(c->sendbuf); cudaFree(c->recvbuf); cudaFree(c->gather); cudaFree(c->d_scalar);
  delete c;
}

__global__ void k_halo_pack(int32_t npro, int ncomp, const int32_t *__restrict__ cell, const double *__restrict__ phi, double *__restrict__ buf) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npro) return;
  const int64_t c = cell[i];
  for (int k = 0; k < ncomp; ++k) buf[(int64_t)ncomp * i + k] = phi[ncomp * c + k];   // exchange.f90:48-66
}
__global__ void k_halo_unpack(int32_t npro, int ncomp, const int32_t *__restrict__ slot, const double *__restrict__ buf, double *__restrict__ phi) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npro) return;
  const int64_t s = slot[i];
  for (int k = 0; k < ncomp; ++k) phi[ncomp * s + k] = buf[(int64_t)ncomp * i + k];   // exchange.f90:110-127
}

// ---- peer-memory exchange: ONE push kernel (pack + NVLink stores into the peer's staging + flag) and ONE pull kernel
// (spin on the local flags + unpack into the ghost slots).  Staging is double buffered by sequence parity: a rank can be
// at most one exchange ahead of a neighbour.
__global__ void __launch_bounds__(256) k_halo_push(const CommDev *__restrict__ cd, int ncomp, const double *__restrict__ phi, unsigned long long seq) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int par = (int)(seq & 1ull);
  if (i < cd->npro) {
    const int r = cd->frank[i];
    double *dst = cd->peer_stage[r] + (long long)par * cd->peer_stride[r] + (long long)ncomp * cd->rord[i];
    const int64_t c = cd->cell[i];
    for (int k = 0; k < ncomp; ++k) dst[k] = phi[ncomp * c + k];   // exchange.f90:48-66 + the MPI_Sendrecv of :81-99
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(&cd->hdr->push_ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) cd->hdr->push_ticket = 0u;
  __threadfence_system();
  if (threadIdx.x < cd->nnb) p2p_st_release(&cd->peer_hdr[cd->nb_rank[threadIdx.x]]->xflag[cd->rank], seq);
}
__global__ void __launch_bounds__(256) k_halo_pull(const CommDev *__restrict__ cd, int ncomp, double *__restrict__ phi, unsigned long long seq) {
  if (threadIdx.x < cd->nnb) p2p_wait(&cd->hdr->xflag[cd->nb_rank[threadIdx.x]], seq, cd->hdr);
  __syncthreads();
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cd->npro) return;
  const double *src = cd->stage + (long long)(seq & 1ull) * cd->stride + (long long)ncomp * i;
  const int64_t s = cd->slot[i];
  for (int k = 0; k < ncomp; ++k) phi[ncomp * s + k] = p2p_ld_data(src + k);     // exchange.f90:110-127
}

static int comm_exchange_mode(fcp_ctx *ctx, double *field, int ncomp, bool p2p) {
  FcpComm *c = ctx->comm;
  if (!c || c->npro == 0) return FCP_OK;
  cudaStream_t st = ctx->stream;
  const int grid = (c->npro + 255) / 256;
  if (p2p) {
    const unsigned long long seq = ++c->xseq;
    size_t tok = ctx->prof.begin(FCP_K_HALO, st);
    k_halo_push<<<grid, 256, 0, st>>>(c->d_dev, ncomp, field, seq);
    k_halo_pull<<<grid, 256, 0, st>>>(c->d_dev, ncomp, field, seq);
    ctx->prof.end(tok, st);
    FCP_LAUNCHED(); FCP_LAUNCHED();
    FCP_CHECK_LAUNCH();
    return FCP_OK;
  }
  k_halo_pack<<<grid, 256, 0, st>>>(c->npro, ncomp, c->d_cell, field, c->sendbuf);
  FCP_LAUNCHED();
  FCP_NCCL(g_nccl.GroupStart());
  for (size_t j = 0; j < c->peer.size(); ++j) {
    FCP_NCCL(g_nccl.Send(c->sendbuf + (size_t)ncomp * c->off[j], (size_t)ncomp * c->cnt[j], ncclDouble, c->peer[j], c->comm, st));
    FCP_NCCL(g_nccl.Recv(c->recvbuf + (size_t)ncomp * c->off[j], (size_t)ncomp * c->cnt[j], ncclDouble, c->peer[j], c->comm, st));
  }
  FCP_NCCL(g_nccl.GroupEnd());
  k_halo_unpack<<<grid, 256, 0, st>>>(c->npro, ncomp, c->d_slot, c->recvbuf, field);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
int comm_exchange(fcp_ctx *ctx, double *field, int ncomp) {
  FcpComm *c = ctx->comm;
  return comm_exchange_mode(ctx, field, ncomp, c && c->p2p);
}

// vals[k] <- sum over ranks (rank 0 first) of vals[k]; every rank gets the same bits
__global__ void k_rank_ordered_sum(int nranks, int count, const double *__restrict__ gathered, double *__restrict__ vals) {
  int k = threadIdx.x;
  if (k >= count) return;
  double s = gathered[k];
  for (int r = 1; r < nranks; ++r) s = s + gathered[r * count + k];
  vals[k] = s;
}
int comm_allgather_sum(FcpComm *c, double *d_vals, int count, cudaStream_t st) {
  if (!c || c->nranks == 1) return FCP_OK;
  FCP_NCCL(g_nccl.AllGather(d_vals, c->gather, (size_t)count, ncclDouble, c->comm, st));
  k_rank_ordered_sum<<<1, 32, 0, st>>>(c->nranks, count, c->gather, d_vals);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

int comm_allreduce_minmax(FcpComm *c, double *d_mm, cudaStream_t st) {   // global_min / global_max, src-par/global_sum_mpi.f90
  if (!c || c->nranks == 1) return FCP_OK;
  FCP_NCCL(g_nccl.AllReduce(d_mm, d_mm, 1, ncclDouble, ncclMin, c->comm, st));
  FCP_NCCL(g_nccl.AllReduce(d_mm + 1, d_mm + 1, 1, ncclDouble, ncclMax, c->comm, st));
  return FCP_OK;
}

// geometry of process faces once the ghost cell centres are known: facint like geometry.f90:581-606 (variant 2),
// Df like :648-664, with the local cell as P and the ghost cell as N (src-par/geometry.f90:826-871 fpro)
__global__ void k_process_face_geom(int32_t npro, const int32_t *__restrict__ pface, const int32_t *__restrict__ cell, const int32_t *__restrict__ slot,
                                    const double *xc, const double *yc, const double *zc, const double *xf, const double *yf, const double *zf,
                                    const double *arx, const double *ary, const double *arz, double *facint, double *Df) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npro) return;
  const int32_t f = pface[i], p = cell[i], q = slot[i];
  double xpn = xf[f] - xc[p], ypn = yf[f] - yc[p], zpn = zf[f] - zc[p];
  const double djp = sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
  xpn = xf[f] - xc[q]; ypn = yf[f] - yc[q]; zpn = zf[f] - zc[q];
  const double djn = sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
  facint[f] = djp / (djp + djn);
  xpn = xc[q] - xc[p]; ypn = yc[q] - yc[p]; zpn = zc[q] - zc[p];
  const double are = arx[f] * arx[f] + ary[f] * ary[f] + arz[f] * arz[f];
  Df[f] = are / (arx[f] * xpn + ary[f] * ypn + arz[f] * zpn);
}

// ---- halo plan: everything the communication layer derives from the patch table, on the HOST and without touching a
// device, so that it can be checked on CPU for any number of neighbours (fcp_comm_plan; tests/test_comm_plan.py).
struct HaloPlan {
  std::vector<int> peer;                 // per process patch (patch order): rank on the other side
  std::vector<int32_t> off, cnt;         // per process patch: offset / count of its faces in the halo buffers
  std::vector<int32_t> cell, slot;       // per process face (patch order): owner cell, ghost slot (0-based field indices)
  std::vector<int32_t> frank;            // per process face: peer rank
  std::vector<int32_t> chunk_ptr;        // [nchunks+1] process faces grouped by the 2048-row chunk of their owner cell
  std::vector<int32_t> chunk_face;       // [npro] face ordinals in chunk order
  std::vector<int32_t> order;            // [nchunks] launch order: chunk | bit 31 when it owns process faces (those first)
  std::vector<int32_t> ghost_ord;        // [B] boundary face -> process-face ordinal (-1 for physical patches)
};
static int halo_plan_build(HaloPlan &pl, int32_t n, int32_t F, int32_t B, int32_t nb, const int32_t *bctype, const int32_t *nfaces,
                           const int32_t *startFace, const int32_t *owner0 /* [F+B] 0-based */, const int32_t *peer_rank, int rank, int nranks) {
  pl = HaloPlan();
  for (int32_t ib = 0; ib < nb; ++ib) {
    if (bctype[ib] != FCP_BC_PROCESS) continue;
    if (!peer_rank || peer_rank[ib] < 0 || peer_rank[ib] >= nranks || peer_rank[ib] == rank) {
      fcp_set_error("fcp_comm_init: process patch %d has no valid peer rank", ib);
      return FCP_EINVAL;
    }
    pl.peer.push_back(peer_rank[ib]);
    pl.off.push_back((int32_t)pl.slot.size());       // faces of a patch are contiguous in both buffers, patches in patch order
    pl.cnt.push_back(nfaces[ib]);
    for (int32_t i = 0; i < nfaces[ib]; ++i) {
      const int32_t f = startFace[ib] + i;
      pl.slot.push_back(n + (f - F));
      pl.cell.push_back(owner0[f]);
      pl.frank.push_back(peer_rank[ib]);
    }
  }
  const int32_t npro = (int32_t)pl.slot.size();
  const int nch = std::max(fcp_nchunks(n), 1);
  pl.chunk_ptr.assign(nch + 1, 0);
  pl.chunk_face.assign(std::max(npro, 1), 0);
  for (int32_t i = 0; i < npro; ++i) pl.chunk_ptr[pl.cell[i] / FCP_CHUNK + 1]++;
  for (int k = 0; k < nch; ++k) pl.chunk_ptr[k + 1] += pl.chunk_ptr[k];
  {
    std::vector<int32_t> fill(pl.chunk_ptr.begin(), pl.chunk_ptr.end() - 1);
    for (int32_t i = 0; i < npro; ++i) pl.chunk_face[fill[pl.cell[i] / FCP_CHUNK]++] = i;
  }
  for (int k = 0; k < nch; ++k) if (pl.chunk_ptr[k + 1] > pl.chunk_ptr[k]) pl.order.push_back(k | (int32_t)0x80000000);
  for (int k = 0; k < nch; ++k) if (pl.chunk_ptr[k + 1] == pl.chunk_ptr[k]) pl.order.push_back(k);
  pl.ghost_ord.assign(std::max(B, 1), -1);
  for (int32_t i = 0; i < npro; ++i) pl.ghost_ord[pl.slot[i] - n] = i;
  return FCP_OK;
}

extern "C" int fcp_comm_plan(const fcp_mesh_desc *md, const int32_t *peer_rank, int rank, int nranks, int32_t *npatch, int32_t *patch_peer,
                             int32_t *patch_off, int32_t *patch_cnt, int32_t *cell, int32_t *slot, int32_t *chunk_ptr, int32_t *chunk_face,
                             int32_t *chunk_order, int32_t *ghost_ord) {
  if (!md) return FCP_EINVAL;
  const int32_t n = md->numCells, F = md->numInnerFaces, B = md->numBoundaryFaces;
  std::vector<int32_t> owner0((size_t)F + B);
  for (size_t f = 0; f < owner0.size(); ++f) owner0[f] = md->owner[f] - 1;
  HaloPlan pl;
  FCP_TRY(halo_plan_build(pl, n, F, B, md->numBoundaries, md->bctype, md->nfaces, md->startFace, owner0.data(), peer_rank, rank, nranks));
  if (npatch) *npatch = (int32_t)pl.peer.size();
  for (size_t j = 0; j < pl.peer.size(); ++j) {
    if (patch_peer) patch_peer[j] = pl.peer[j];
    if (patch_off) patch_off[j] = pl.off[j];
    if (patch_cnt) patch_cnt[j] = pl.cnt[j];
  }
  if (cell) std::copy(pl.cell.begin(), pl.cell.end(), cell);
  if (slot) std::copy(pl.slot.begin(), pl.slot.end(), slot);
  if (chunk_ptr) std::copy(pl.chunk_ptr.begin(), pl.chunk_ptr.end(), chunk_ptr);
  if (chunk_face && !pl.slot.empty()) std::copy(pl.chunk_face.begin(), pl.chunk_face.begin() + pl.slot.size(), chunk_face);
  if (chunk_order) std::copy(pl.order.begin(), pl.order.end(), chunk_order);
  if (ghost_ord && B > 0) std::copy(pl.ghost_ord.begin(), pl.ghost_ord.begin() + B, ghost_ord);
  return FCP_OK;
}

// ---- peer-memory set-up ------------------------------------------------------------------------------------------
struct WinRecord {               // what every rank publishes about its window (all-gathered through NCCL)
  cudaIpcMemHandle_t handle;     // 64 bytes
  long long off_stage, stride, off_ll, npro;
  int ok, pad;
};
__global__ void k_i32_to_f64(int32_t n, const int32_t *__restrict__ src, double *__restrict__ dst, int32_t add_index) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)(src ? src[i] : 0) + (add_index ? (double)i : 0.0);
}
__global__ void k_f64_to_i32(int32_t n, const double *__restrict__ src, int32_t *__restrict__ dst) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (int32_t)src[i];
}
// per-face integers of the matching face on the peer, through the NCCL send/recv path (set-up only)
static int nccl_swap_face_ints(fcp_ctx *ctx, FcpComm *c, const int32_t *d_src /* nullptr: the face ordinal */, int32_t *d_dst) {
  cudaStream_t st = ctx->stream;
  const int grid = (c->npro + 255) / 256;
  if (c->npro == 0) return FCP_OK;
  k_i32_to_f64<<<grid, 256, 0, st>>>(c->npro, d_src, c->sendbuf, d_src ? 0 : 1);
  FCP_NCCL(g_nccl.GroupStart());
  for (size_t j = 0; j < c->peer.size(); ++j) {
    FCP_NCCL(g_nccl.Send(c->sendbuf + c->off[j], (size_t)c->cnt[j], ncclDouble, c->peer[j], c->comm, st));
    FCP_NCCL(g_nccl.Recv(c->recvbuf + c->off[j], (size_t)c->cnt[j], ncclDouble, c->peer[j], c->comm, st));
  }
  FCP_NCCL(g_nccl.GroupEnd());
  k_f64_to_i32<<<grid, 256, 0, st>>>(c->npro, c->recvbuf, d_dst);
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

static int p2p_setup(fcp_ctx *ctx, FcpComm *c, const HaloPlan &pl) {
  const char *env = getenv("FCP_COMM");
  const bool want = !(env && !strcmp(env, "nccl")) && c->nranks <= FCP_MAXR && c->nranks > 1;
  cudaStream_t st = ctx->stream;
  // window: header | staging [2][3*npro] doubles | LL slots [npro][2] words
  const size_t hdr_bytes = (sizeof(WinHeader) + 255) / 256 * 256;
  const long long stride = 3ll * std::max(c->npro, 1);
  const size_t stage_bytes = ((size_t)2 * stride * sizeof(double) + 255) / 256 * 256;
  const size_t ll_bytes = (size_t)std::max(c->npro, 1) * 16;
  const size_t win_bytes = hdr_bytes + stage_bytes + ll_bytes;
  WinRecord mine;
  memset(&mine, 0, sizeof(mine));
  mine.off_stage = (long long)hdr_bytes; mine.stride = stride; mine.off_ll = (long long)(hdr_bytes + stage_bytes); mine.npro = c->npro;
  mine.ok = 0;
  if (want) {
    if (cudaMalloc(&c->win, win_bytes) == cudaSuccess && cudaMemset(c->win, 0, win_bytes) == cudaSuccess &&
        cudaDeviceSynchronize() == cudaSuccess && cudaIpcGetMemHandle(&mine.handle, c->win) == cudaSuccess)
      mine.ok = 1;
    else cudaGetLastError();
  }
  // all-gather the records (NCCL doubles as the barrier that orders every rank's memset before any peer store)
  static_assert(sizeof(WinRecord) % 8 == 0, "WinRecord is gathered as 8-byte words");
  const size_t words = sizeof(WinRecord) / 8;
  double *d_rec = nullptr;
  FCP_TRY(dev_alloc(&d_rec, words * (c->nranks + 1)));
  FCP_CUDA(cudaMemcpyAsync(d_rec + words * c->nranks, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  FCP_NCCL(g_nccl.AllGather(d_rec + words * c->nranks, d_rec, words, ncclDouble, c->comm, st));
  std::vector<WinRecord> recs(c->nranks);
  FCP_CUDA(cudaMemcpyAsync(recs.data(), d_rec, sizeof(WinRecord) * c->nranks, cudaMemcpyDeviceToHost, st));
  FCP_CUDA(cudaStreamSynchronize(st));
  bool all_ok = want;
  for (auto &r : recs) all_ok = all_ok && r.ok;
  c->peer_win.assign(c->nranks, nullptr);
  int opened = 1;
  if (all_ok) {
    for (int r = 0; r < c->nranks; ++r) {
      if (r == c->rank) { c->peer_win[r] = c->win; continue; }
      if (cudaIpcOpenMemHandle(&c->peer_win[r], recs[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        c->peer_win[r] = nullptr;
        opened = 0;
      }
    }
  } else opened = 0;
  // every rank must take the same path: min over ranks of `opened`
  {
    double v = (double)opened;
    FCP_CUDA(cudaMemcpyAsync(c->d_scalar, &v, sizeof(double), cudaMemcpyHostToDevice, st));
    FCP_NCCL(g_nccl.AllReduce(c->d_scalar, c->d_scalar, 1, ncclDouble, ncclMin, c->comm, st));
    FCP_CUDA(cudaMemcpyAsync(&v, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, st));
    FCP_CUDA(cudaStreamSynchronize(st));
    opened = (int)v;
  }
  cudaFree(d_rec);
  if (!opened) {
    for (int r = 0; r < c->nranks; ++r)
      if (r != c->rank && c->peer_win[r]) { cudaIpcCloseMemHandle(c->peer_win[r]); c->peer_win[r] = nullptr; }
    if (env && !strcmp(env, "p2p")) { fcp_set_error("FCP_COMM=p2p requested but CUDA IPC peer mapping is not available"); return FCP_ENCCL; }
    c->p2p = false;
    return FCP_OK;   // NCCL send/recv + all-gather path
  }
  // per-face data of the matching face on the peer
  { std::vector<int32_t> fr(pl.frank); if (fr.empty()) fr.push_back(0); FCP_TRY(dev_upload(&c->d_frank, fr.data(), fr.size())); }
  FCP_TRY(dev_alloc(&c->d_rord, (size_t)std::max(c->npro, 1)));
  FCP_TRY(nccl_swap_face_ints(ctx, c, nullptr, c->d_rord));
  std::vector<int32_t> rord(std::max(c->npro, 1), 0);
  FCP_CUDA(cudaMemcpyAsync(rord.data(), c->d_rord, sizeof(int32_t) * (size_t)c->npro, cudaMemcpyDeviceToHost, st));
  FCP_CUDA(cudaStreamSynchronize(st));
  // fused-push lists in chunk order: owner cell and the address of the face's LL slot in the peer's window
  std::vector<int32_t> pcell(std::max(c->npro, 1), 0);
  std::vector<unsigned long long *> pdst(std::max(c->npro, 1), nullptr);
  for (int32_t j = 0; j < c->npro; ++j) {
    const int32_t i = pl.chunk_face[j];
    pcell[j] = pl.cell[i];
    pdst[j] = (unsigned long long *)((char *)c->peer_win[pl.frank[i]] + recs[pl.frank[i]].off_ll) + 2 * (size_t)rord[i];
  }
  FCP_TRY(dev_upload(&c->d_chunk_ptr, pl.chunk_ptr.data(), pl.chunk_ptr.size()));
  FCP_TRY(dev_upload(&c->d_push_cell, pcell.data(), pcell.size()));
  FCP_TRY(dev_upload(&c->d_push_ord, pl.chunk_face.data(), pl.chunk_face.size()));
  FCP_TRY(dev_upload(&c->d_push_dst, pdst.data(), pdst.size()));
  FCP_TRY(dev_upload(&c->d_order, pl.order.data(), pl.order.size()));
  FCP_TRY(dev_upload(&c->d_ghost_ord, pl.ghost_ord.data(), pl.ghost_ord.size()));
  CommDev &d = c->h_dev;
  memset(&d, 0, sizeof(d));
  d.rank = c->rank; d.nranks = c->nranks;
  std::vector<int> nb(c->peer.begin(), c->peer.end());
  std::sort(nb.begin(), nb.end());
  nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
  d.nnb = (int)nb.size();
  for (int k = 0; k < d.nnb; ++k) d.nb_rank[k] = nb[k];
  for (int r = 0; r < c->nranks; ++r) {
    char *base = (char *)c->peer_win[r];
    d.peer_hdr[r] = (WinHeader *)base;
    d.peer_stage[r] = (double *)(base + recs[r].off_stage);
    d.peer_stride[r] = recs[r].stride;
  }
  d.hdr = d.peer_hdr[c->rank]; d.stage = d.peer_stage[c->rank]; d.stride = stride;
  d.ll = (unsigned long long *)((char *)c->win + mine.off_ll);
  d.npro = c->npro; d.n = ctx->n; d.cell = c->d_cell; d.slot = c->d_slot; d.frank = c->d_frank; d.rord = c->d_rord;
  d.ghost_ord = c->d_ghost_ord; d.chunk_ptr = c->d_chunk_ptr; d.push_cell = c->d_push_cell; d.push_ord = c->d_push_ord; d.push_dst = c->d_push_dst; d.order = c->d_order;
  {
    const char *e = getenv("FCP_P2P_POLL");
    const unsigned int gpu_scope = (e && !strcmp(e, "gpu")) ? 1u : 0u;
    FCP_CUDA(cudaMemcpyAsync(&d.hdr->poll_gpu_scope, &gpu_scope, sizeof(gpu_scope), cudaMemcpyHostToDevice, st));
  }
  FCP_CUDA(cudaMalloc((void **)&c->d_dev, sizeof(CommDev)));
  FCP_CUDA(cudaMemcpyAsync(c->d_dev, &d, sizeof(CommDev), cudaMemcpyHostToDevice, st));
  FCP_CUDA(cudaStreamSynchronize(st));
  c->p2p = true;
  return FCP_OK;
}

static int p2p_set_timeout(FcpComm *c, unsigned long long ns, cudaStream_t st) {
  FCP_CUDA(cudaMemcpyAsync(&c->h_dev.hdr->timeout_ns, &ns, sizeof(ns), cudaMemcpyHostToDevice, st));
  FCP_CUDA(cudaStreamSynchronize(st));
  return FCP_OK;
}
__global__ void k_selfcheck_fill(int32_t n, int32_t nT, int rank, double *__restrict__ t) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nT) t[i] = i < n ? (double)rank * 16777216.0 + (double)i : -1.0;
}
// Start-up self-check of the peer-memory layout: the same cell-tagged field is exchanged once through the NVLink windows
// and once through NCCL send/recv; every rank compares its ghost slots bit for bit and the ranks agree (min) on the
// outcome.  On any difference or time-out ALL ranks drop to the NCCL path (unless FCP_COMM=p2p insists): a wrong layout
// must never turn into a silent wrong answer or a hang.
static int p2p_selfcheck(fcp_ctx *ctx, FcpComm *c) {
  cudaStream_t st = ctx->stream;
  const int32_t nT = ctx->nT, n = ctx->n, B = ctx->B;
  double *t1 = nullptr, *t2 = nullptr;
  FCP_TRY(dev_alloc(&t1, (size_t)std::max(nT, 1)));
  FCP_TRY(dev_alloc(&t2, (size_t)std::max(nT, 1)));
  if (nT > 0) {
    k_selfcheck_fill<<<(nT + 255) / 256, 256, 0, st>>>(n, nT, c->rank, t1);
    FCP_CUDA(cudaMemcpyAsync(t2, t1, sizeof(double) * (size_t)nT, cudaMemcpyDeviceToDevice, st));
  }
  FCP_TRY(p2p_set_timeout(c, 2000000000ull, st));
  FCP_TRY(comm_exchange_mode(ctx, t1, 1, true));
  FCP_TRY(comm_exchange_mode(ctx, t2, 1, false));
  std::vector<double> g1(std::max(B, 1)), g2(std::max(B, 1));
  if (B > 0) {
    FCP_CUDA(cudaMemcpyAsync(g1.data(), t1 + n, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, st));
    FCP_CUDA(cudaMemcpyAsync(g2.data(), t2 + n, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, st));
  }
  int err = 0;
  FCP_CUDA(cudaMemcpyAsync(&err, &c->h_dev.hdr->error, sizeof(int), cudaMemcpyDeviceToHost, st));
  FCP_CUDA(cudaStreamSynchronize(st));
  double ok = (!err && (B == 0 || memcmp(g1.data(), g2.data(), sizeof(double) * (size_t)B) == 0)) ? 1.0 : 0.0;
  FCP_CUDA(cudaMemcpyAsync(c->d_scalar, &ok, sizeof(double), cudaMemcpyHostToDevice, st));
  FCP_NCCL(g_nccl.AllReduce(c->d_scalar, c->d_scalar, 1, ncclDouble, ncclMin, c->comm, st));
  FCP_CUDA(cudaMemcpyAsync(&ok, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, st));
  FCP_CUDA(cudaStreamSynchronize(st));
  cudaFree(t1); cudaFree(t2);
  FCP_TRY(p2p_set_timeout(c, 20000000000ull, st));
  if (ok < 0.5) {
    const char *env = getenv("FCP_COMM");
    if (env && !strcmp(env, "p2p")) { fcp_set_error("peer-memory self-check failed (ghost values differ from the NCCL exchange or a peer timed out)"); return FCP_ENCCL; }
    fprintf(stderr, "libfcp_b200: rank %d: peer-memory self-check failed, using the NCCL path\n", c->rank);
    const int zero = 0;
    FCP_CUDA(cudaMemcpyAsync(&c->h_dev.hdr->error, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    FCP_CUDA(cudaStreamSynchronize(st));
    c->p2p = false;
  }
  return FCP_OK;
}

// 1 when the peer-memory (CUDA IPC / NVLink) path is active, 0 when the context uses NCCL send/recv, -1 without a communicator
extern "C" int fcp_comm_mode(const fcp_ctx *ctx) {
  if (!ctx || !ctx->comm) return -1;
  return ctx->comm->p2p ? 1 : 0;
}
// raised by a kernel that gave up waiting for a peer (p2p.cuh: p2p_wait)
const int *comm_error_flag(const FcpComm *c) { return (c && c->p2p) ? &c->h_dev.hdr->error : nullptr; }   // device address
int comm_check_error(fcp_ctx *ctx) {
  FcpComm *c = ctx->comm;
  if (!c || !c->p2p) return FCP_OK;
  int err = 0;
  FCP_CUDA(cudaMemcpyAsync(&err, &c->h_dev.hdr->error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FCP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (err) { fcp_set_error("peer-memory protocol timeout: a rank stopped responding"); return FCP_ENCCL; }
  return FCP_OK;
}

__global__ void k_set_process_facint(int32_t npro, const int32_t *__restrict__ pface, const double *__restrict__ fpro, double *__restrict__ facint) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npro) facint[pface[i]] = fpro[i];
}
extern "C" int fcp_set_process_facint(fcp_ctx *ctx, const double *fpro, int32_t count) {
  if (!ctx || (!fpro && count > 0)) return FCP_EINVAL;
  if (!ctx->comm) { fcp_set_error("fcp_set_process_facint: call fcp_comm_init first (it fills the process-face geometry this call overrides)"); return FCP_ESTATE; }
  if (count != ctx->npro) { fcp_set_error("fcp_set_process_facint: %d values for %d process faces", count, ctx->npro); return FCP_EINVAL; }
  if (count == 0) return FCP_OK;
  FCP_CUDA(cudaSetDevice(ctx->device));
  double *d = nullptr;
  FCP_TRY(dev_upload(&d, fpro, (size_t)count));
  k_set_process_facint<<<(count + 255) / 256, 256, 0, ctx->stream>>>(count, ctx->d_procface, d, ctx->facint);
  ctx->og_valid = false;
  FCP_LAUNCHED();
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) { fcp_set_error("fcp_set_process_facint: %s", cudaGetErrorString(e)); return FCP_ECUDA; }
  return FCP_OK;
}

extern "C" int fcp_set_process_orientation(fcp_ctx *ctx, const int32_t *flipped, int32_t count) {
  if (!ctx || (!flipped && count > 0)) return FCP_EINVAL;
  if (!ctx->comm) { fcp_set_error("fcp_set_process_orientation: call fcp_comm_init first"); return FCP_ESTATE; }
  if (count != ctx->npro) { fcp_set_error("fcp_set_process_orientation: %d values for %d process faces", count, ctx->npro); return FCP_EINVAL; }
  FCP_CUDA(cudaSetDevice(ctx->device));
  std::vector<int32_t> h((size_t)std::max(ctx->B, 1), 0);
  for (int32_t i = 0; i < count; ++i) h[ctx->h_procface[i] - ctx->F] = flipped[i] ? 1 : 0;
  if (ctx->d_proc_flip) { cudaFree(ctx->d_proc_flip); ctx->d_proc_flip = nullptr; }
  FCP_TRY(dev_upload(&ctx->d_proc_flip, h.data(), h.size()));
  return FCP_OK;
}

extern "C" int fcp_comm_unique_id(void *id128) {
  if (!id128) return FCP_EINVAL;
  FCP_TRY(nccl_load());
  ncclUniqueId id;
  FCP_NCCL(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return FCP_OK;
}

extern "C" int fcp_comm_init(fcp_ctx *ctx, int rank, int nranks, const void *id128, const int32_t *peer_rank) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { fcp_set_error("fcp_comm_init: bad argument"); return FCP_EINVAL; }
  if (ctx->comm) { fcp_set_error("fcp_comm_init: communicator already initialised"); return FCP_ESTATE; }
  FCP_CUDA(cudaSetDevice(ctx->device));
  FCP_TRY(nccl_load());
  FcpComm *c = new FcpComm();
  c->rank = rank;
  c->nranks = nranks;
  c->npro = ctx->npro;
  HaloPlan pl;
  {
    std::vector<int32_t> owner(ctx->nF);
    FCP_CUDA(cudaMemcpy(owner.data(), ctx->owner, sizeof(int32_t) * (size_t)ctx->nF, cudaMemcpyDeviceToHost));
    int rc = halo_plan_build(pl, ctx->n, ctx->F, ctx->B, ctx->nb, ctx->bctype.data(), ctx->nfaces.data(), ctx->startFace.data(), owner.data(),
                             peer_rank, rank, nranks);
    if (rc != FCP_OK) { delete c; return rc; }
  }
  c->peer = pl.peer; c->off = pl.off; c->cnt = pl.cnt;
  std::vector<int32_t> cell = pl.cell, slot = pl.slot;
  if ((int32_t)cell.size() != ctx->npro) { fcp_set_error("fcp_comm_init: internal process-face count mismatch"); delete c; return FCP_EINVAL; }
  if (cell.empty()) { cell.push_back(0); slot.push_back(0); }
  FCP_TRY(dev_upload(&c->d_cell, cell.data(), cell.size()));
  FCP_TRY(dev_upload(&c->d_slot, slot.data(), slot.size()));
  FCP_TRY(dev_alloc(&c->sendbuf, (size_t)3 * std::max(ctx->npro, 1)));
  FCP_TRY(dev_alloc(&c->recvbuf, (size_t)3 * std::max(ctx->npro, 1)));
  FCP_TRY(dev_alloc(&c->gather, (size_t)4 * nranks));
  FCP_TRY(dev_alloc(&c->d_scalar, 4));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  FCP_NCCL(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
  ctx->comm = c;
  FCP_TRY(p2p_setup(ctx, c, pl));
  if (c->p2p) FCP_TRY(p2p_selfcheck(ctx, c));
  // ghost copies of the cell-centre data (src-par/geometry.f90:769-773) and the process-face geometry
  FCP_TRY(comm_exchange(ctx, ctx->xc, 1));
  FCP_TRY(comm_exchange(ctx, ctx->yc, 1));
  FCP_TRY(comm_exchange(ctx, ctx->zc, 1));
  FCP_TRY(comm_exchange(ctx, ctx->vol, 1));
  if (ctx->npro) {
    k_process_face_geom<<<(ctx->npro + 255) / 256, 256, 0, ctx->stream>>>(ctx->npro, ctx->d_procface, c->d_cell, c->d_slot, ctx->xc, ctx->yc, ctx->zc,
                                                                          ctx->xf, ctx->yf, ctx->zf, ctx->arx, ctx->ary, ctx->arz, ctx->facint, ctx->Df);
    ctx->og_valid = false;
    FCP_LAUNCHED();
    FCP_CHECK_LAUNCH();
  }
  // Patch-type flags that gate collectives (the ppref broadcast, the global outlet sum of adjustMassFlow) and the ppref = 0 rule must be the
  // same on every rank: a real src-par decomposition omits a patch on the ranks where it has no faces.  Global flag = max over ranks.
  {
    double h_flags[3] = {ctx->has_pressure_patch ? 1.0 : 0.0, ctx->has_outlet ? 1.0 : 0.0, ctx->has_inout ? 1.0 : 0.0};
    FCP_CUDA(cudaMemcpyAsync(c->d_scalar, h_flags, sizeof(h_flags), cudaMemcpyHostToDevice, ctx->stream));
    FCP_TRY(comm_allgather_sum(c, c->d_scalar, 3, ctx->stream));
    FCP_CUDA(cudaMemcpyAsync(h_flags, c->d_scalar, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream));
    FCP_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->g_pressure_patch = h_flags[0] > 0.0;
    ctx->g_outlet = h_flags[1] > 0.0;
    ctx->g_inout = h_flags[2] > 0.0;
  }
  FCP_CUDA(cudaStreamSynchronize(ctx->stream));
  return FCP_OK;
}

extern "C" int fcp_exchange(fcp_ctx *ctx, int field) {
  if (!ctx) return FCP_EINVAL;
  if (field < 0 || (field >= FCP_F_FLMASS && field <= FCP_F_H) || field >= FCP_F_COUNT /* every other id is a cell field */) { fcp_set_error("fcp_exchange: field %d is not a cell field", field); return FCP_EINVAL; }
  void *p = nullptr;
  FCP_TRY(fcp_field_devptr(ctx, field, &p, nullptr));
  return comm_exchange(ctx, (double *)p, fcp_is_gradient_field(field) ? 3 : 1);
}

static int global_reduce(fcp_ctx *ctx, double *value, int op /*0 sum 1 max 2 min*/) {
  if (!ctx || !value) return FCP_EINVAL;
  FcpComm *c = ctx->comm;
  if (!c || c->nranks == 1) return FCP_OK;
  FCP_CUDA(cudaMemcpyAsync(c->d_scalar, value, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (op == 0) FCP_TRY(comm_allgather_sum(c, c->d_scalar, 1, ctx->stream));
  else FCP_NCCL(g_nccl.AllReduce(c->d_scalar, c->d_scalar, 1, ncclDouble, op == 1 ? ncclMax : ncclMin, c->comm, ctx->stream));
  FCP_CUDA(cudaMemcpyAsync(value, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FCP_CUDA(cudaStreamSynchronize(ctx->stream));
  return FCP_OK;
}
extern "C" int fcp_global_sum(fcp_ctx *ctx, double *value) { return global_reduce(ctx, value, 0); }
extern "C" int fcp_global_max(fcp_ctx *ctx, double *value) { return global_reduce(ctx, value, 1); }
extern "C" int fcp_global_min(fcp_ctx *ctx, double *value) { return global_reduce(ctx, value, 2); }
// integer sum through the rank-ordered double sum: exact while the total stays below 2^53 (src-par uses it for nnz and cell counts)
extern "C" int fcp_global_isum(fcp_ctx *ctx, int64_t *value) {
  if (!ctx || !value) return FCP_EINVAL;
  if (*value > (int64_t)1 << 52 || *value < -((int64_t)1 << 52)) { fcp_set_error("fcp_global_isum: |value| exceeds 2^52"); return FCP_EINVAL; }
  double v = (double)*value;
  FCP_TRY(global_reduce(ctx, &v, 0));
  *value = (int64_t)v;
  return FCP_OK;
}
