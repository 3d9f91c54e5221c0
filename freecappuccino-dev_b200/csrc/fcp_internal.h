// fcp_internal.h -- internal structures of libfcp_b200.so (not part of the C-ABI, see include/fcp.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/fcp.h"

// ---------------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------------
void fcp_set_error(const char *fmt, ...);
#include <atomic>
extern std::atomic<int64_t> g_fcp_launches;   // number of kernels launched by this library (fcp_launch_count)

#define FCP_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      fcp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));    \
      return FCP_ECUDA;                                                                        \
    }                                                                                          \
  } while (0)
#define FCP_TRY(call)                 \
  do {                                \
    int r__ = (call);                 \
    if (r__ != FCP_OK) return r__;    \
  } while (0)
#define FCP_LAUNCHED() (++g_fcp_launches)
#define FCP_CHECK_LAUNCH() FCP_CUDA(cudaGetLastError())

// parameters.f90:6  `small = 1e-20` is a default-real literal promoted to double
#define FCP_SMALL ((double)1e-20f)

// ---------------------------------------------------------------------------------------------
// The reduction / work decomposition every vector kernel uses (mirrored by oracle/orc.cpp:orc_sum_tree):
// a CTA of 256 threads owns a CHUNK of 2048 consecutive items; thread t handles items t, t+256, ... in that
// order; warp xor-butterfly; the 8 warp sums are added in warp order; the chunk partials are reduced the same way
// by the last CTA to finish.  Fixed tree => run-to-run bitwise identical sums and iteration counts.
// ---------------------------------------------------------------------------------------------
#define FCP_TPB 256
#define FCP_IPT 8
#define FCP_CHUNK (FCP_TPB * FCP_IPT)
static inline int fcp_nchunks(int64_t n) { return (int)((n + FCP_CHUNK - 1) / FCP_CHUNK); }

// ---------------------------------------------------------------------------------------------
// SELL-32 storage.  Rows are grouped in slices of 32 (one warp); within a slice entry j of lane l is at
// slptr[slice] + 32*j + l, so that a warp reads one 256-byte line of values (128 bytes of indices) per column step.
// The order of the entries inside a row is the CSR order (columns ascending, diagonal embedded, halo columns
// >= n last), so row sums round exactly like the reference's sequential CSR loops.
// ---------------------------------------------------------------------------------------------
// one triangle of the matrix in level-tile order (see SellPattern::tri)
struct TriTiles {
  int32_t ntiles = 0, maxlen = 0;
  int64_t nent = 0;               // 32 * sum of the tile lengths
  int64_t *tptr = nullptr;        // [ntiles+1]
  int32_t *tcol = nullptr;        // [nent] column (0-based) or -1
  int32_t *tsrc = nullptr;        // [nent] SELL position of the entry in a() or -1
  int32_t *ttsrc = nullptr;       // [nent] forward only, built with tpos: SELL position of the transposed entry (ILU(0)) or -1
  const int32_t *prow = nullptr;  // = plev_rows / pblev_rows
  double *tval = nullptr;         // [nent] values of this solve's matrix
  double *ttval = nullptr;        // [nent] transposed values (ILU(0) factor), forward only
  double *dtile = nullptr;        // [32 ntiles] the factor diagonal d in tile order
};
struct SellPattern {
  int32_t n = 0;        // rows (numCells)
  int32_t ncols = 0;    // columns (numTotal when halo columns exist)
  int64_t nnz = 0;      // local CSR entries (host-visible a(nnz))
  int64_t nnz_ext = 0;  // nnz + halo entries
  int64_t nnzp = 0;     // padded SELL entries
  int32_t nslices = 0;
  int32_t tile_cap = 0;  // max number of padded entries in a tile of 8 slices (256 rows): stage size of the TMA SpMV
  // host copies (kept: csr pattern queries, conversions)
  std::vector<int32_t> h_ia, h_ja, h_diag;     // local CSR, 1-based (the reference's ia, ja, diag)
  // device
  int64_t *slptr = nullptr;   // [nslices+1]
  int32_t *rinfo = nullptr;   // [n] len | dpos<<16   (len counts halo entries; dpos = offset of the diagonal in the row)
  int32_t *ja = nullptr;      // [nnzp] 0-based column; padding = own row
  int32_t *ia0 = nullptr;     // [n+1] 0-based local CSR row pointer (for CSR<->SELL value conversion)
  int32_t *llen = nullptr;    // [n] number of LOCAL entries (== len when there are no halo columns); nullptr if identical
  // level schedule of the lower-triangular part (IC(0)/ILU(0) sweeps), built lazily
  bool levels_built = false;
  int32_t nlevels = 0;
  int32_t *lev_ptr = nullptr;   // [nlevels+1] offsets into lev_rows
  int32_t *lev_rows = nullptr;  // [n] rows ordered by level (ascending row index inside a level)
  std::vector<int32_t> h_lev_ptr;
  int32_t nblevels = 0;         // same for the backward sweep (dependencies = entries after the diagonal)
  int32_t *blev_ptr = nullptr, *blev_rows = nullptr;
  std::vector<int32_t> h_blev_ptr;
  int32_t *tpos = nullptr;      // [nnzp] SELL position of the transposed entry (ILU(0) of BiCGStab), lazily
  // barrier-free sweeps (FCP_SWEEP=flags): the same level order with every level padded to whole warps (-1 = no row), one ready flag per row
  int32_t *plev_rows = nullptr, *pblev_rows = nullptr;
  int32_t nplev = 0, npblev = 0;   // padded lengths (multiples of 32)
  int32_t *ready = nullptr;        // [n] epoch of the last sweep that finished the row
  int32_t sweep_epoch = 0;         // host counter: a preconditioner apply uses epoch+1 (forward) and epoch+2 (backward)
  // flag-in-data sweeps (default): the strictly lower / strictly upper triangle re-stored per TILE of 32 consecutive positions of the padded
  // level order (tile t, entry k, lane l at ttptr[t] + 32 k + l), so that a warp reads its rows' entries as full lines although the rows of a
  // level are scattered over the SELL slices; entries keep the CSR order of their row.  Structure once per pattern, values once per solve.
  TriTiles tri[2];                 // [0] forward (entries before the diagonal), [1] backward (local entries after it)
  unsigned long long *zll[3] = {nullptr, nullptr, nullptr};   // [2n] LL words each: forward result, backward result, factor diagonal
  unsigned int ll_epoch = 0;       // sequence number of the last sweep / factor that used the LL arrays
};

// cell -> faces gather lists, SELL-32 as well (ent, other, slot share the layout)
struct FaceLists {
  int64_t nnzp = 0;
  int64_t *slptr = nullptr;   // [nslices+1]
  int32_t *len = nullptr;     // [n]
  int32_t *ent = nullptr;     // +(f+1): cell is the face's owner (P side), -(f+1): neighbour (N side); 0 padding
  int32_t *other = nullptr;   // field index of the value across the face (cell, ghost slot or boundary slot), 0-based
  int32_t *slot = nullptr;    // SELL position of a(cell,other) or -1 - bctype for physical boundary faces
  int32_t *gent = nullptr;    // like ent, but +-(g+1) with g = the face's position in the OWNER-ORDERED geometry arrays (fcp_ctx::og): see fvm_ensure_og
  unsigned long long *kinds = nullptr;   // [n] compact form of len + the sign of slot: nibble k = 0 (two-sided face) or 1 + bctype, top byte = len (255: > 14 faces)
};

// Krylov workspace (linear_solvers.f90:33  res,reso,pk,zk,d,uk,vk) + device scalars
struct KrylovScalars {
  double red[4];      // raw sums of the kernel that just finished (after the cross-rank sum in multi-GPU runs)
  double sk, s0, pkapk, resl, res0, factor, resor;
  double bet, alf, om, gam, beto, ukreso, vkres, vkvk;
  double tol_abs, tol_rel;
  int32_t iters, done, itr_max, pad;
};
struct KrylovWS {
  int32_t n = 0, ncols = 0;
  double *res = nullptr, *pk = nullptr, *zk = nullptr, *adiag = nullptr, *d = nullptr;
  double *reso = nullptr, *uk = nullptr, *vk = nullptr, *tmp = nullptr;
  KrylovScalars *sc = nullptr;        // device
  KrylovScalars *h_sc = nullptr;      // pinned host mirror
  int32_t *h_poll = nullptr;          // pinned [4]: the `done` flag ([0,1]) and the peer-memory error word ([2,3]) after each of the two batches in flight
  double *partials = nullptr;         // [4 * maxchunks]
  unsigned int *counter = nullptr;    // last-block ticket
  unsigned int *bar = nullptr;        // [4] grid barrier of the persistent solver kernels: arrivals, generation
  unsigned long long *phase_ns = nullptr;   // [8] per-phase times of the persistent kernels (CTA 0's view), profiler only
  int persist_grid[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // cooperative grid size per cooperative kernel variant on ws_device (0: not queried yet)
  int ws_device = -1;
  int maxchunks = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr};
};

struct FcpComm;   // comm.cu

// ---------------------------------------------------------------------------------------------
// Peer-memory communication (comm.cu).  Every rank owns one device "window" (header + halo staging) that all peers of
// the node map through CUDA IPC.  Inside a Krylov iteration nothing but the compute kernels runs: the kernel that
// produces the direction vector stores the values its neighbours need straight into their windows over NVLink, and
// the last CTA of every reducing kernel exchanges the partial sums the same way.  Both use flag-in-data words
// (p2p.cuh: a 32-bit sequence number rides in every 8-byte word), so there is no fence, no flag and no host round trip.
// The generic exchange(phi) outside the iteration uses a staged push/pull with a sequence flag.
// ---------------------------------------------------------------------------------------------
#define FCP_MAXR 16
struct WinHeader {
  unsigned long long xflag[FCP_MAXR];        // written by peer r: sequence number of its last complete generic halo push
  unsigned long long rll[2][FCP_MAXR][8];    // written by peer r: LL words of its <= 4 partial sums, two slots by sequence parity
  // local only
  unsigned long long red_seq;
  unsigned long long timeout_ns;             // how long a spin may last before it raises `error` (20 s; 2 s during the start-up self-check)
  unsigned int push_ticket;
  unsigned int poll_gpu_scope;               // 1: polls of this rank's OWN window use relaxed GPU-scope loads (FCP_P2P_POLL=gpu) instead of system-scope volatile ones
  int error, pad1;
};
struct CommDev {
  int rank, nranks, nnb;                   // nnb: number of distinct neighbour ranks
  int nb_rank[FCP_MAXR];
  WinHeader *hdr;                          // own header
  WinHeader *peer_hdr[FCP_MAXR];           // by rank (self included)
  double *peer_stage[FCP_MAXR];            // by rank: that rank's staging area [2][stride] (generic exchange)
  long long peer_stride[FCP_MAXR];         // 3 * npro of that rank
  double *stage;                           // own staging
  long long stride;
  unsigned long long *ll;                  // own LL slots of the fused direction-vector halo: [npro][2]
  int32_t npro, n;
  const int32_t *cell, *slot;              // per process face (patch order): owner cell, ghost slot
  const int32_t *frank, *rord;             // peer rank and the face's ordinal on the peer
  const int32_t *ghost_ord;                // [numBoundaryFaces] boundary face -> process-face ordinal (-1 for physical patches)
  // fused push, faces grouped by the 2048-row chunk of their owner cell
  const int32_t *chunk_ptr;                // [nchunks+1]
  const int32_t *push_cell;                // [npro] owner cell, chunk order
  const int32_t *push_ord;                 // [npro] process-face ordinal (patch order: index into cell / slot / ll), chunk order
  unsigned long long *const *push_dst;     // [npro] address of the face's LL slot in the peer's window, chunk order
  const int32_t *order;                    // [nchunks] launch order: chunk index | bit 31 when the chunk owns process faces (those first)
};

// per-kernel-class CUDA-event timing (fcp_profile_*): an event pair around every launch of a class
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Rec { int k; size_t e0, e1; };
  std::vector<Rec> recs;
  double total_ms[FCP_K_COUNT] = {0};
  int64_t launches[FCP_K_COUNT] = {0};
  size_t begin(int k, cudaStream_t st) {
    if (!on) return 0;
    while (pool.size() < used + 2) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    recs.push_back(Rec{k, used, used + 1});
    cudaEventRecord(pool[used], st);
    used += 2;
    return recs.size();
  }
  void end(size_t tok, cudaStream_t st) { if (on && tok) cudaEventRecord(pool[recs[tok - 1].e1], st); }
  void resolve() {
    for (auto &r : recs) {
      float ms = 0.f;
      cudaEventSynchronize(pool[r.e1]);
      if (cudaEventElapsedTime(&ms, pool[r.e0], pool[r.e1]) == cudaSuccess) { total_ms[r.k] += ms; launches[r.k] += 1; }
    }
    recs.clear();
    used = 0;
  }
  void reset() { resolve(); for (int k = 0; k < FCP_K_COUNT; ++k) { total_ms[k] = 0; launches[k] = 0; } }
  ~Profiler() { for (auto e : pool) cudaEventDestroy(e); }
};
#define FCP_PROF(prof, k, st, stmt)                      \
  do {                                                   \
    Profiler *p__ = (prof);                              \
    size_t tok__ = p__ ? p__->begin(k, st) : 0;          \
    stmt;                                                \
    if (p__) p__->end(tok__, st);                        \
  } while (0)

struct fcp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int32_t n = 0, F = 0, B = 0, nT = 0, nF = 0, nb = 0, npro = 0;
  std::vector<int32_t> bctype, nfaces, startFace;   // host patch table (startFace 0-based)
  std::vector<int32_t> h_kPN, h_kNP;                // 1-based host CSR positions (icell_jcell / jcell_icell)
  // mesh on device
  int32_t *owner = nullptr, *neigh = nullptr;       // 0-based
  double *arx = nullptr, *ary = nullptr, *arz = nullptr, *xf = nullptr, *yf = nullptr, *zf = nullptr;  // [nF]
  double *facint = nullptr, *Df = nullptr;          // [F] (+ process faces in global orientation when partitioned: [nF])
  double *xc = nullptr, *yc = nullptr, *zc = nullptr, *vol = nullptr;   // [nT]
  int32_t *bftype = nullptr;                        // [B] patch type of each boundary face
  int32_t *kPN = nullptr, *kNP = nullptr;           // [F] SELL positions of a(P,N), a(N,P)
  SellPattern pat;
  FaceLists fl;
  // owner-ordered face geometry: face f sits at gpos[f] = og_slptr[owner >> 5] + 32 * (rank of f among its owner's faces) + (owner & 31), i.e. in the SELL
  // layout of the OWNER cells.  Lanes that walk consecutive cells then read consecutive addresses for a face's area vector / factor / centre (their own
  // faces and, on any banded numbering, the faces owned by their neighbours) instead of every third element of a face-ordered array.
  int32_t *d_gpos = nullptr;      // [nF]
  int64_t og_n = 0;               // padded length of one owner-ordered array
  double *og = nullptr;           // [7][og_n]: arx, ary, arz, facint, xf, yf, zf; built on demand (fvm_ensure_og)
  bool og_valid = false;          // reset when the process-face geometry / factors change (fcp_comm_init, fcp_set_process_facint)
  double *field[FCP_F_COUNT] = {nullptr};
  double *Dmat[4] = {nullptr, nullptr, nullptr, nullptr};   // LSQ matrices per method (index FCP_GRAD_*); QR: [18][n]
  int32_t max_cell_faces = 0;                       // longest cell->face list (the QR gradient holds at most 6)
  double *d_mmpart = nullptr;                       // limiter: per-chunk {min,max} partials + the final pair
  double *d_sum = nullptr;                          // [8] device scalar slots (calcp_piso pavg; constant_mass_flow_forcing sums)
  KrylovWS ws;
  FcpComm *comm = nullptr;
  Profiler prof;
  double *flushbuf = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  bool has_pressure_patch = false, has_outlet = false, has_inout = false;   // from the LOCAL patch table
  bool g_pressure_patch = false, g_outlet = false, g_inout = false;         // the same over ALL ranks (fcp_comm_init); == local without a communicator
  int uvw_smem_configured = 0;                      // bit v: the dynamic shared-memory limit of k_uvw_assemble variant v has been raised on this context's device
  int flux_variant = 0, flux_grad_method = 0;       // fcp_set_flux_variant: 1 = the MPI tree's facefluxmass on inner faces (quirk Q10), gradients by this method
  int32_t nout = 0;
  int32_t *d_oface = nullptr;                       // outlet faces in patch order (adjustMassFlow)
  double *d_flowo = nullptr;                        // [4] outlet mass flow: local sum, then the sum over all ranks
  double *d_csr_stage = nullptr;                    // [nnz] CSR-order staging for uploads / downloads of a(nnz) (allocated on first use)
  int32_t *d_aprpos = nullptr;                      // [npro] SELL position of the halo entry of each process face
  int32_t *d_procface = nullptr;                    // [npro] 0-based face index of each process face (patch order)
  int32_t *d_proc_flip = nullptr;                   // [B] per boundary face: 1 = a process face whose owner in the unpartitioned mesh is the PEER's cell (fcp_set_process_orientation)
  std::vector<int32_t> h_procface;
  double *d_ppref = nullptr;                        // [4] broadcast slot for pp(pRefCell)
  // periodic pairs (sparse_matrix.f90:141-171): per BOUNDARY face, for the faces of a periodic patch and of its twin patch
  int32_t nper = 0;                                 // numPeriodic: periodic faces, each pair counted once
  int32_t *per_cell = nullptr;                      // [B] the cell across the pair (0-based), -1 for every other boundary face
  int32_t *per_face = nullptr;                      // [B] the other face of the pair (0-based face index)
  int32_t *per_slot = nullptr;                      // [B] SELL position of a(owner of this face, per_cell)
  double *per_df = nullptr;                         // [B] the "Df(i)" of the pair (quirk Q21: the Df of inner face number i = the pair's ordinal in its patch)
};

struct fcp_solver {
  int device = 0;
  cudaStream_t stream = nullptr;
  SellPattern pat;
  double *a = nullptr;        // SELL values
  double *a_csr = nullptr;    // staging (device, nnz)
  double *fi = nullptr, *rhs = nullptr;
  KrylovWS ws;
};

// ---- fvm.cu -------------------------------------------------------------------------------------
struct CorrectArgs {          // velocity / pressure correction fused into gradp_and_sources(pp), calcp_simple.f90:416-429
  double *u, *v, *w, *pres;
  const double *apu, *apv, *apw;
  double urfp;
  const double *ppref_src;    // device address of pp(pRefCell) or nullptr -> 0 (pressure patches present)
};
struct AsmArgs {
  const double *den, *u, *v, *w, *p, *dPdxi, *apu;
  const double *apv, *apw;    // read on periodic faces only (facefluxmass2_periodic)
  double *pp;                 // boundary values of pp on pressure patches are zeroed
  double *ub, *vb, *wb;       // same arrays as u,v,w (boundary slots written on pressure patches)
  double *a, *su, *flmass;
  const double *gU = nullptr, *gV = nullptr, *gW = nullptr;   // non-null: inner faces use the MPI tree's facefluxmass (quirk Q10) with these velocity gradients
};
int fvm_ensure_og(fcp_ctx *ctx);
int fvm_grad_gauss(fcp_ctx *ctx, const double *u, double *g);
int fvm_lsq_matrix(fcp_ctx *ctx, bool weighted, double *D);
int fvm_grad_lsq(fcp_ctx *ctx, bool weighted, const double *D, const double *phi, double *g, int row2_reference);
int fvm_laplacian(fcp_ctx *ctx, const double *mu, const double *phi, double *a, double *su);
int fvm_gradp(fcp_ctx *ctx, int scheme, double *p, const double *apu, double *su, double *sv, double *sw, double *dPdxi, double *gtmp,
              const CorrectArgs *correct);
int fvm_assemble_pcorr(fcp_ctx *ctx, const AsmArgs &g, bool piso = false);
int fvm_adjust_mass_flow(fcp_ctx *ctx, int32_t nout, const int32_t *d_oface, const double *den, double *u, double *v, double *w,
                         double *flmass, double flomas);
int fvm_correct_flux(fcp_ctx *ctx, const double *a, const double *pp, double *flmass);
int fvm_correct_flux_periodic(fcp_ctx *ctx, const double *a, const double *x, double *flmass);
int fvm_correct_pressure_bnd(fcp_ctx *ctx, const double *den, const double *apu, const double *pp, double *u, double *v, double *w,
                             double *flmass);
int fvm_nonorth(fcp_ctx *ctx, const double *den, const double *apu, const double *dPdxi, double *su, double *flmass);

// ---- fvm_ext.cu ---------------------------------------------------------------------------------
int fvm_slope_limiter(fcp_ctx *ctx, int limiter, const double *phi, double *g);
int fvm_lsq_qr_matrix(fcp_ctx *ctx, double *D);
int fvm_grad_lsq_qr(fcp_ctx *ctx, const double *D, const double *phi, double *g);
int fvm_piso_hbya(fcp_ctx *ctx, const double *h, const double *rU, const double *rV, const double *rW, const double *apu, const double *apv,
                  const double *apw, double *u, double *v, double *w, double *su, double *sv, double *sw);
int fvm_sum(fcp_ctx *ctx, const double *x, double *d_out);
int fvm_piso_pupdate(fcp_ctx *ctx, double ncells_global, double urfp, const double *d_sum, const double *pp, double *p);
int fvm_cmf_forcing(fcp_ctx *ctx, double magUbar, const double *apu, double *u, double *d_sums);
int fvm_update_boundary(fcp_ctx *ctx, double *phi);
int fvm_neg_vol(fcp_ctx *ctx, double *q);
int fvm_wall_distance_finish(fcp_ctx *ctx, int stage, double *phi, const double *g, double *wd);
int fvm_piso_fluxmc(fcp_ctx *ctx, const double *den, const double *apu, const double *dPdxi, double *su);

// ---- fvm_uvw.cu ---------------------------------------------------------------------------------
struct UvwArgs {
  const double *u, *v, *w, *vis, *visw, *den, *flmass;
  const double *dUdxi, *dVdxi, *dWdxi;
  const double *uo, *vo, *wo, *uoo, *voo, *woo, *uooo, *vooo, *wooo;
  double *a, *su, *sv, *sw, *spu, *spv, *sp;
  double *rU, *rV, *rW;          // nullptr unless piso
  double gds, timestep, gradPcmf, viscos;
  int cscheme, tscheme, const_mflux;
};
int fvm_update_vel_bnd(fcp_ctx *ctx, double *u, double *v, double *w);
int fvm_uvw_assemble(fcp_ctx *ctx, const UvwArgs &g);
int fvm_uvw_diag(fcp_ctx *ctx, double *a, const double *spq, const double *srcq, const double *phi, double *apq, double *su, double urf, int zero_first);

// ---- fvm_scalar.cu (row f4) ---------------------------------------------------------------------
struct ScParams {
  int kind, cscheme, tscheme, lowre;
  const double *fsst, *walldist, *gte;
  double gds, prtr, viscos, densit, timestep, urf;
  const double *phi, *phio, *phioo, *te, *ed, *den, *vis, *visw, *dnw, *flmass, *u, *v, *w, *magStrain, *su_vol, *sp_vol, *grad;
  double *gen, *tau, *a, *su, *sp, *phi_new, *phi_out;
};
int fvm_strain(fcp_ctx *ctx, const double *gU, const double *gV, const double *gW, double *magStrain, double *vorticity);
int fvm_sc_assemble(fcp_ctx *ctx, const ScParams &q);
int fvm_clip_small(fcp_ctx *ctx, double *phi, int32_t count);
int fvm_mu_eff_sst(fcp_ctx *ctx, double urf, double viscos, double densit, int lowre, const double *magStrain, const double *walldist, const double *te,
                   const double *ed, const double *den, const double *u, const double *v, const double *w, const double *dnw, double *vis, double *visw,
                   double *ypl, double *tau);
int fvm_sst_blend(fcp_ctx *ctx, double viscos, const double *walldist, const double *gte, const double *gom, const double *den, const double *te,
                  const double *ed, double *fsst);
int fvm_mu_eff_rlzb(fcp_ctx *ctx, double urf, double viscos, const double *gU, const double *gV, const double *gW, const double *te, const double *ed,
                    const double *den, const double *u, const double *v, const double *w, const double *dnw, double *vis, double *visw, double *ypl,
                    double *tau);
int fvm_minmax(fcp_ctx *ctx, const double *phi, double **mm_out, int32_t count = -1 /* default: numCells */);
static inline bool fcp_is_gradient_field(int f) { return (f >= FCP_F_DUDXI && f <= FCP_F_G1) || f == FCP_F_DTEDXI || f == FCP_F_DEDDXI; }
int fvm_grad_gauss_fvx(fcp_ctx *ctx, const double *u, double *gtmp, double *g);
int fvm_grad_gauss_passes(fcp_ctx *ctx, const double *u, double *gtmp, double *g, int npass);
int fvm_sgs_viscosity(fcp_ctx *ctx, int model, double urf, double viscos, const double *gU, const double *gV, const double *gW, const double *den,
                      double *vis, double *visw);

// ---- pattern.cu ---------------------------------------------------------------------------------
int sell_from_csr(SellPattern &p, int32_t n, int32_t ncols, const int32_t *ia1, const int32_t *ja1, const int32_t *diag1,
                  const std::vector<std::vector<int32_t>> *halo_cols /* per row extra columns (0-based), may be null */);
void sell_free(SellPattern &p);
int sell_build_levels(SellPattern &p, cudaStream_t st);
int sell_build_tpos(SellPattern &p, cudaStream_t st);
int sell_build_tiles(SellPattern &p, bool with_transposed, cudaStream_t st);
int sell_values_from_csr(const SellPattern &p, const double *d_a_csr, double *d_a_sell, cudaStream_t st);
int sell_values_to_csr(const SellPattern &p, const double *d_a_sell, double *d_a_csr, cudaStream_t st);
template <class T> int dev_upload(T **dptr, const T *h, size_t count);
template <class T> int dev_alloc(T **dptr, size_t count);

// ---- krylov.cu ----------------------------------------------------------------------------------
int krylov_ws_alloc(KrylovWS &ws, int32_t n, int32_t ncols);
void krylov_ws_free(KrylovWS &ws);
int sell_spmv(const SellPattern &p, const double *a, const double *x, double *y, cudaStream_t st);
int krylov_solve(int solver, SellPattern &p, const double *a, double *fi, const double *rhs, KrylovWS &ws,
                 int32_t itr_max, double tol_abs, double tol_rel, fcp_report *rep, cudaStream_t st, FcpComm *comm,
                 fcp_ctx *ctx);

// ---- comm.cu ------------------------------------------------------------------------------------
int comm_exchange(fcp_ctx *ctx, double *field, int ncomp);   // ncomp 1 (scalar) or 3 (gradient)
int comm_allgather_sum(FcpComm *comm, double *d_vals, int count, cudaStream_t st);  // rank-ordered deterministic sum, in place
int comm_allreduce_minmax(FcpComm *comm, double *d_mm /* {min,max} */, cudaStream_t st);
void comm_free(FcpComm *comm);
int comm_nranks(const FcpComm *comm);
const CommDev *comm_dev(const FcpComm *comm);       // device descriptor, nullptr unless the peer-memory path is active
const CommDev *comm_dev_host(const FcpComm *comm);  // its host copy (the pointers inside are device pointers)
const int32_t *comm_chunk_info(const FcpComm *comm); // device copy of CommDev::order, nullptr unless the peer-memory path is active
unsigned int comm_pk_base(const FcpComm *comm);     // sequence base of the fused direction-vector pushes of the next solve
void comm_pk_advance(FcpComm *comm, int32_t iters); // after a solve that ran `iters` iterations (identical on all ranks)
int comm_check_error(fcp_ctx *ctx);                 // synchronises the stream
const int *comm_error_flag(const FcpComm *comm);    // device address of the window's error word, nullptr unless the peer-memory path is active
