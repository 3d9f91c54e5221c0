// tests/emu/emu.cpp -- TEST INFRASTRUCTURE ONLY: the run-time half of the CPU emulation declared in cuda_runtime.h / nccl.h of
// this directory (fibers for the threads of a block, a shared-memory "device" arena, blocking stand-ins for the stream,
// event, IPC and NCCL calls csrc/*.cu makes).  Not part of the product, never measured, never shipped.
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include "cuda_runtime.h"
#include "nccl.h"

thread_local uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
thread_local dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);

// =================================================================================================================
// fibers
// =================================================================================================================
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.hidden emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

namespace {
enum { ST_READY = 0, ST_WARP = 1, ST_BLOCK = 2, ST_GRID = 3, ST_DONE = 4 };
const size_t kStack = 256 * 1024;
const int kMaxThreads = 1024;

struct Fiber {
  void *sp = nullptr;
  int state = ST_DONE;
  unsigned mask = 0;
  char *stack = nullptr;
};
struct Runner {   // one per OS thread
  std::vector<Fiber> f;
  int n = 0, cur = 0, live = 0;
  void *sched_sp = nullptr;
  const std::function<void()> *body = nullptr;
  pthread_barrier_t *gbar = nullptr;
  unsigned char xbuf[kMaxThreads][16];
  unsigned char *dsmem = nullptr;
  int pred[kMaxThreads];
  char *stacks = nullptr;
  ~Runner() { if (stacks) munmap(stacks, kStack * kMaxThreads); free(dsmem); }
};
thread_local Runner *R = nullptr;
Runner *runner() {
  if (!R) {
    static thread_local Runner holder;
    R = &holder;
    R->stacks = (char *)mmap(nullptr, kStack * kMaxThreads, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (R->stacks == MAP_FAILED) { perror("emu: mmap stacks"); abort(); }
    R->dsmem = (unsigned char *)aligned_alloc(1024, 256 * 1024);
    R->f.resize(kMaxThreads);
    for (int i = 0; i < kMaxThreads; ++i) R->f[i].stack = R->stacks + kStack * i;
  }
  return R;
}

void set_thread_index(int t) {
  threadIdx.x = (unsigned)t % blockDim.x;
  threadIdx.y = ((unsigned)t / blockDim.x) % blockDim.y;
  threadIdx.z = (unsigned)t / (blockDim.x * blockDim.y);
}
void yield_to_scheduler() {
  Runner *r = R;
  const int me = r->cur;
  emu_switch(&r->f[me].sp, r->sched_sp);
  // resumed: the scheduler has set cur and the index registers
}
void fiber_entry() {
  Runner *r = R;
  (*r->body)();
  r->f[r->cur].state = ST_DONE;
  r->live--;
  yield_to_scheduler();
  fprintf(stderr, "emu: a finished fiber was resumed\n");
  abort();
}
void run_block(int nthreads, const std::function<void()> &body) {
  Runner *r = runner();
  if (nthreads > kMaxThreads) { fprintf(stderr, "emu: %d threads per block\n", nthreads); abort(); }
  r->n = nthreads; r->live = nthreads; r->body = &body;
  for (int t = 0; t < nthreads; ++t) {
    Fiber &fb = r->f[t];
    void **top = (void **)(fb.stack + kStack);
    top[-1] = nullptr;                    // fake return address: the entry function sees rsp % 16 == 8 like after a call
    top[-2] = (void *)&fiber_entry;       // `ret` of emu_switch jumps here
    for (int k = 3; k <= 8; ++k) top[-k] = nullptr;   // r15 r14 r13 r12 rbx rbp
    fb.sp = (void *)(top - 8);
    fb.state = ST_READY;
    fb.mask = 0;
  }
  static const bool rev_threads = getenv("FCP_EMU_REVERSE") != nullptr;   // race hunting: run the fibers (and the blocks) in descending order
  while (r->live > 0) {
    bool progressed = false;
    for (int tt = 0; tt < nthreads; ++tt) {
      const int t = rev_threads ? nthreads - 1 - tt : tt;
      if (r->f[t].state != ST_READY) continue;
      r->cur = t;
      set_thread_index(t);
      emu_switch(&r->sched_sp, r->f[t].sp);
      progressed = true;
    }
    if (r->live == 0) break;
    // warp-level waits: released when every lane named by the waiter's mask has arrived (or exited)
    for (int w = 0; w * 32 < nthreads; ++w) {
      const int lo = w * 32, hi = std::min(nthreads, lo + 32);
      bool any = false, all = true;
      for (int t = lo; t < hi && all; ++t) {
        if (r->f[t].state != ST_WARP) continue;
        any = true;
        for (int u = lo; u < hi; ++u)
          if (((r->f[t].mask >> (u - lo)) & 1u) && r->f[u].state != ST_WARP && r->f[u].state != ST_DONE) { all = false; break; }
      }
      if (any && all) {
        for (int t = lo; t < hi; ++t) if (r->f[t].state == ST_WARP) r->f[t].state = ST_READY;
        progressed = true;
      }
    }
    int nblock = 0, ngrid = 0;
    for (int t = 0; t < nthreads; ++t) { nblock += r->f[t].state == ST_BLOCK; ngrid += r->f[t].state == ST_GRID; }
    if (nblock == r->live) {
      for (int t = 0; t < nthreads; ++t) if (r->f[t].state == ST_BLOCK) r->f[t].state = ST_READY;
      progressed = true;
    } else if (ngrid == r->live) {
      if (r->gbar) pthread_barrier_wait(r->gbar);
      for (int t = 0; t < nthreads; ++t) if (r->f[t].state == ST_GRID) r->f[t].state = ST_READY;
      progressed = true;
    }
    if (!progressed) {
      fprintf(stderr, "emu: deadlock in block (%u,%u,%u): live=%d at_block_barrier=%d at_grid_barrier=%d; thread states:", blockIdx.x, blockIdx.y,
              blockIdx.z, r->live, nblock, ngrid);
      for (int t = 0; t < nthreads; ++t) fprintf(stderr, "%d", r->f[t].state);
      fprintf(stderr, "\n");
      abort();
    }
  }
}
}   // namespace

namespace emu {
void sync_block() { R->f[R->cur].state = ST_BLOCK; yield_to_scheduler(); }
void sync_grid() { R->f[R->cur].state = ST_GRID; yield_to_scheduler(); }
static void warp_wait(unsigned mask) { Fiber &fb = R->f[R->cur]; fb.state = ST_WARP; fb.mask = mask; yield_to_scheduler(); }
void warp_exchange(unsigned mask, const void *mine, void *out, int src_lane, size_t bytes) {
  Runner *r = R;
  const int me = r->cur, lo = me & ~31;
  if (bytes > 16) { fprintf(stderr, "emu: shuffle of %zu bytes\n", bytes); abort(); }
  memcpy(r->xbuf[me], mine, bytes);
  warp_wait(mask);
  const int src = lo + (src_lane & 31);
  const bool ok = src < r->n && ((mask >> (src_lane & 31)) & 1u) && r->f[src].state != ST_DONE;
  memcpy(out, ok ? (const void *)r->xbuf[src] : mine, bytes);
  warp_wait(mask);
}
unsigned warp_ballot(unsigned mask, int pred) {
  Runner *r = R;
  const int me = r->cur, lo = me & ~31;
  r->pred[me] = pred != 0;
  warp_wait(mask);
  unsigned out = 0;
  for (int l = 0; l < 32 && lo + l < r->n; ++l)
    if (((mask >> l) & 1u) && r->f[lo + l].state != ST_DONE && r->pred[lo + l]) out |= 1u << l;
  warp_wait(mask);
  return out;
}
void yield() { R->f[R->cur].state = ST_READY; yield_to_scheduler(); }
void os_yield() { sched_yield(); }
unsigned char *dyn_smem() { return R->dsmem; }
unsigned long long now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

static void emu_invalid_configuration();
void launch_impl(const Cfg &c, const std::function<void()> &body, bool cooperative) {
  const int nthreads = (int)(c.b.x * c.b.y * c.b.z);
  const long long nblocks = (long long)c.g.x * c.g.y * c.g.z;
  if (nthreads <= 0 || nthreads > 1024 || nblocks <= 0 || c.g.y > 65535 || c.g.z > 65535 || c.g.x > 2147483647u) {
    emu_invalid_configuration();   // a real launch fails with "invalid configuration" (empty grids included): the next FCP_CHECK_LAUNCH sees it
    return;
  }
  const uint3 save_b = blockIdx, save_t = threadIdx;
  const dim3 save_bd = blockDim, save_gd = gridDim;
  if (!cooperative) {
    blockDim = c.b; gridDim = c.g;
    // FCP_EMU_REVERSE: blocks in descending order -- a kernel whose result depends on the order in which its blocks (or the threads of
    // a block between two barriers) happen to run has a race; the parity tests then fail under one of the two orders
    static const bool rev = getenv("FCP_EMU_REVERSE") != nullptr;
    for (long long bb = 0; bb < nblocks; ++bb) {
      const long long b = rev ? nblocks - 1 - bb : bb;
      blockIdx = uint3{(unsigned)(b % c.g.x), (unsigned)((b / c.g.x) % c.g.y), (unsigned)(b / ((long long)c.g.x * c.g.y))};
      run_block(nthreads, body);
    }
  } else {
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, (unsigned)nblocks);
    std::vector<std::thread> th;
    for (long long b = 0; b < nblocks; ++b)
      th.emplace_back([&, b]() {
        blockDim = c.b; gridDim = c.g;
        blockIdx = uint3{(unsigned)(b % c.g.x), (unsigned)((b / c.g.x) % c.g.y), (unsigned)(b / ((long long)c.g.x * c.g.y))};
        runner()->gbar = &bar;
        run_block(nthreads, body);
        R->gbar = nullptr;
      });
    for (auto &t : th) t.join();
    pthread_barrier_destroy(&bar);
  }
  blockIdx = save_b; threadIdx = save_t; blockDim = save_bd; gridDim = save_gd;
}
}   // namespace emu

// =================================================================================================================
// "device" memory: one process-wide arena on a memfd so that another process can map it (cudaIpc*)
// =================================================================================================================
namespace {
const size_t kArena = (size_t)24 << 30;
const unsigned long long kPoison = 0x7ff4dead7ff4deadull;   // a signalling NaN as double, an absurd index as int32
struct Arena {
  std::mutex mu;
  int fd = -1;
  char *base = nullptr;
  std::map<size_t, size_t> free_;   // offset -> size
  std::map<size_t, size_t> used_;   // block offset -> start of its page run
  std::map<size_t, size_t> span_;   // start of a page run -> its length (pages of the block + one guard page)
  std::map<int, char *> peers;      // pid -> mapping of that process's arena
  void init() {
    if (base) return;
    fd = memfd_create("fcpemu_arena", 0);
    if (fd < 0 || ftruncate(fd, (off_t)kArena) != 0) { perror("emu: arena"); abort(); }
    base = (char *)mmap(nullptr, kArena, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
    if (base == MAP_FAILED) { perror("emu: arena mmap"); abort(); }
    free_[0] = kArena;
  }
};
Arena g_arena;
cudaError_t g_last = cudaSuccess;
int g_device = 0;
}   // namespace

// Every allocation ends flush (to 256 bytes) against a PROT_NONE guard page: an overrun of a "device" array faults at once instead of
// silently reading a neighbour, and the slack in front of it is poisoned like the block itself.
static const size_t kPage = 4096;
cudaError_t emuMalloc(void **p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_arena.mu);
  g_arena.init();
  const size_t need = std::max<size_t>((bytes + 255) / 256 * 256, 256);
  const size_t npages = (need + kPage - 1) / kPage, total = (npages + 1) * kPage;
  for (auto it = g_arena.free_.begin(); it != g_arena.free_.end(); ++it) {
    if (it->second < total) continue;
    const size_t off = it->first, sz = it->second;
    g_arena.free_.erase(it);
    if (sz > total) g_arena.free_[off + total] = sz - total;
    const size_t poff = off + npages * kPage - need;          // the block's own offset
    g_arena.used_[poff] = off;                                 // -> start of its page run
    g_arena.span_[off] = total;
    unsigned long long *q = (unsigned long long *)(g_arena.base + off);
    for (size_t i = 0; i < npages * kPage / 8; ++i) q[i] = kPoison;   // reading memory the product never wrote shows up as NaN / a wild index
    if (mprotect(g_arena.base + off + npages * kPage, kPage, PROT_NONE) != 0) { perror("emu: mprotect guard"); abort(); }
    *p = g_arena.base + poff;
    return cudaSuccess;
  }
  *p = nullptr;
  return g_last = cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) {
  if (!p) return cudaSuccess;
  std::lock_guard<std::mutex> lk(g_arena.mu);
  const size_t poff = (size_t)((char *)p - g_arena.base);
  auto it = g_arena.used_.find(poff);
  if (it == g_arena.used_.end()) return g_last = cudaErrorInvalidValue;
  size_t o = it->second, sz = g_arena.span_[o];
  g_arena.used_.erase(it);
  g_arena.span_.erase(o);
  mprotect(g_arena.base + o + sz - kPage, kPage, PROT_READ | PROT_WRITE);
  if (sz >= (1u << 20)) fallocate(g_arena.fd, FALLOC_FL_PUNCH_HOLE | FALLOC_FL_KEEP_SIZE, (off_t)o, (off_t)sz);   // give the pages back
  auto nx = g_arena.free_.lower_bound(o);
  if (nx != g_arena.free_.end() && o + sz == nx->first) { sz += nx->second; nx = g_arena.free_.erase(nx); }
  if (nx != g_arena.free_.begin()) {
    auto pv = std::prev(nx);
    if (pv->first + pv->second == o) { o = pv->first; sz += pv->second; g_arena.free_.erase(pv); }
  }
  g_arena.free_[o] = sz;
  return cudaSuccess;
}
cudaError_t emuMallocHost(void **p, size_t bytes) { return posix_memalign(p, 256, std::max<size_t>(bytes, 256)) == 0 ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) { if (bytes) memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) { if (bytes) memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemset(void *dst, int v, size_t bytes) { if (bytes) memset(dst, v, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *dst, int v, size_t bytes, cudaStream_t) { if (bytes) memset(dst, v, bytes); return cudaSuccess; }

struct emuStream { int id; };
struct emuEvent { unsigned long long t; };
cudaError_t cudaStreamCreate(cudaStream_t *st) { *st = new emuStream{1}; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned) { *st = new emuStream{1}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t st) { delete st; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emuEvent{0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emuEvent{0}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = emu::now_ns(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)((double)(b->t - a->t) * 1e-6); return cudaSuccess; }
namespace emu { static void emu_invalid_configuration() { g_last = cudaErrorInvalidConfiguration; } }
cudaError_t cudaGetLastError() { cudaError_t e = g_last; g_last = cudaSuccess; return e; }
cudaError_t cudaPeekAtLastError() { return g_last; }
const char *cudaGetErrorString(cudaError_t e) {
  switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument (emulation)";
    case cudaErrorMemoryAllocation: return "out of memory (emulation)";
    case cudaErrorInvalidConfiguration: return "invalid configuration argument (emulation: empty grid or block, or a dimension over the limit)";
    default: return "error (emulation)";
  }
}
cudaError_t cudaSetDevice(int dev) { g_device = dev; return cudaSuccess; }
cudaError_t cudaGetDevice(int *dev) { *dev = g_device; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 16; return cudaSuccess; }
static int emu_sm_count() { const char *e = getenv("FCP_EMU_SMS"); const int v = e ? atoi(e) : 2; return v > 0 ? v : 2; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  memset(p, 0, sizeof(*p));
  snprintf(p->name, sizeof(p->name), "CPU emulation of sm_100 (tests only)");
  p->totalGlobalMem = kArena; p->major = 10; p->minor = 0; p->multiProcessorCount = emu_sm_count(); p->cooperativeLaunch = 1;
  p->sharedMemPerBlockOptin = 227 * 1024;
  return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) {
  switch (a) {
    case cudaDevAttrMultiProcessorCount: *v = emu_sm_count(); break;
    case cudaDevAttrComputeCapabilityMajor: *v = 10; break;
    case cudaDevAttrComputeCapabilityMinor: *v = 0; break;
    case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = 227 * 1024; break;
    default: *v = 0; return g_last = cudaErrorInvalidValue;
  }
  return cudaSuccess;
}
cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b) { *free_b = kArena / 2; *total_b = kArena; return cudaSuccess; }

struct EmuIpc { unsigned long long magic; int pid, fd; unsigned long long off; };
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
  if (!p || !g_arena.base) return g_last = cudaErrorInvalidValue;
  EmuIpc r{0x46435045ull, (int)getpid(), g_arena.fd, (unsigned long long)((char *)p - g_arena.base)};
  memset(h, 0, sizeof(*h));
  memcpy(h->reserved, &r, sizeof(r));
  return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
  EmuIpc r;
  memcpy(&r, h.reserved, sizeof(r));
  if (r.magic != 0x46435045ull || getenv("FCP_EMU_NO_IPC")) return g_last = cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lk(g_arena.mu);
  auto it = g_arena.peers.find(r.pid);
  if (it == g_arena.peers.end()) {
    char path[64];
    snprintf(path, sizeof(path), "/proc/%d/fd/%d", r.pid, r.fd);
    const int fd = open(path, O_RDWR);
    if (fd < 0) return g_last = cudaErrorInvalidValue;
    char *m = (char *)mmap(nullptr, kArena, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return g_last = cudaErrorInvalidValue;
    it = g_arena.peers.emplace(r.pid, m).first;
  }
  *p = it->second + r.off;
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }   // peer mappings live until the process exits

// =================================================================================================================
// NCCL stand-in between processes of one host: a POSIX shared-memory segment named by the unique id
// =================================================================================================================
namespace {
const int kMaxRanks = 16;
const size_t kSlot = (size_t)8 << 20;   // largest message / gather contribution
struct Mail { volatile unsigned long long full; unsigned long long bytes; };
struct NcclShm {
  volatile int bar_count, bar_gen;
  volatile int attached;
  Mail mail[kMaxRanks][kMaxRanks];
};
size_t shm_bytes(int n) { return 4096 + sizeof(NcclShm) + (size_t)n * n * kSlot + (size_t)n * kSlot; }
void spin_check(unsigned long long t0, const char *what) {
  sched_yield();
  if (emu::now_ns() - t0 > 120ull * 1000000000ull) { fprintf(stderr, "emu nccl: %s timed out\n", what); abort(); }
}
struct PendingOp { bool send; void *buf; size_t bytes; int peer; struct emuNcclComm *c; };
thread_local int g_group = 0;
thread_local std::vector<PendingOp> g_ops;
size_t dt_size(ncclDataType_t dt) {
  switch (dt) { case ncclInt8: case ncclUint8: return 1; case ncclFloat16: return 2; case ncclInt32: case ncclUint32: case ncclFloat32: return 4; default: return 8; }
}
}   // namespace
struct emuNcclComm {
  int rank, n;
  NcclShm *sh;
  char *base;
  size_t bytes;
  char *slot(int src, int dst) { return base + 4096 + sizeof(NcclShm) + ((size_t)src * n + dst) * kSlot; }
  char *gslot(int r) { return base + 4096 + sizeof(NcclShm) + (size_t)n * n * kSlot + (size_t)r * kSlot; }
  void barrier() {
    const int gen = sh->bar_gen;
    if (__atomic_add_fetch(&sh->bar_count, 1, __ATOMIC_SEQ_CST) == n) {
      sh->bar_count = 0;
      __atomic_add_fetch(&sh->bar_gen, 1, __ATOMIC_SEQ_CST);
    } else {
      const unsigned long long t0 = emu::now_ns();
      while (__atomic_load_n(&sh->bar_gen, __ATOMIC_SEQ_CST) == gen) spin_check(t0, "barrier");
    }
  }
};
static void do_send(const PendingOp &o) {
  emuNcclComm *c = o.c;
  if (o.bytes > kSlot) { fprintf(stderr, "emu nccl: message of %zu bytes\n", o.bytes); abort(); }
  Mail &m = c->sh->mail[c->rank][o.peer];
  const unsigned long long t0 = emu::now_ns();
  while (__atomic_load_n(&m.full, __ATOMIC_ACQUIRE)) spin_check(t0, "send");
  memcpy(c->slot(c->rank, o.peer), o.buf, o.bytes);
  m.bytes = o.bytes;
  __atomic_store_n(&m.full, 1ull, __ATOMIC_RELEASE);
}
static void do_recv(const PendingOp &o) {
  emuNcclComm *c = o.c;
  Mail &m = c->sh->mail[o.peer][c->rank];
  const unsigned long long t0 = emu::now_ns();
  while (!__atomic_load_n(&m.full, __ATOMIC_ACQUIRE)) spin_check(t0, "recv");
  if (m.bytes != o.bytes) { fprintf(stderr, "emu nccl: rank %d expected %zu bytes from %d, got %llu\n", c->rank, o.bytes, o.peer, m.bytes); abort(); }
  memcpy(o.buf, c->slot(o.peer, c->rank), o.bytes);
  __atomic_store_n(&m.full, 0ull, __ATOMIC_RELEASE);
}
extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
  memset(id, 0, sizeof(*id));
  snprintf(id->internal, sizeof(id->internal), "/fcpemu_nccl_%d_%llx", (int)getpid(), emu::now_ns());
  return ncclSuccess;
}
ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank) {
  if (nranks > kMaxRanks) return ncclInvalidArgument;
  const size_t bytes = shm_bytes(nranks);
  int fd = -1;
  const unsigned long long t0 = emu::now_ns();
  if (rank == 0) {
    fd = shm_open(id.internal, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) { perror("emu nccl: shm_open"); return ncclSystemError; }
  } else {
    for (;;) {
      fd = shm_open(id.internal, O_RDWR, 0600);
      struct stat sb;
      if (fd >= 0 && fstat(fd, &sb) == 0 && (size_t)sb.st_size >= bytes) break;
      if (fd >= 0) close(fd);
      spin_check(t0, "attach");
    }
  }
  char *base = (char *)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
  close(fd);
  if (base == MAP_FAILED) return ncclSystemError;
  emuNcclComm *c = new emuNcclComm{rank, nranks, (NcclShm *)(base + 4096), base, bytes};
  __atomic_add_fetch(&c->sh->attached, 1, __ATOMIC_SEQ_CST);
  while (__atomic_load_n(&c->sh->attached, __ATOMIC_SEQ_CST) < nranks) spin_check(t0, "init");
  c->barrier();
  if (rank == 0) shm_unlink(id.internal);   // everyone has it mapped: nothing is left behind in /dev/shm
  *comm = c;
  return ncclSuccess;
}
ncclResult_t ncclCommDestroy(ncclComm_t c) { if (c) { munmap(c->base, c->bytes); delete c; } return ncclSuccess; }
ncclResult_t ncclGroupStart() { ++g_group; return ncclSuccess; }
ncclResult_t ncclGroupEnd() {
  if (--g_group > 0) return ncclSuccess;
  for (auto &o : g_ops) if (o.send) do_send(o);
  for (auto &o : g_ops) if (!o.send) do_recv(o);
  g_ops.clear();
  return ncclSuccess;
}
ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t c, cudaStream_t) {
  PendingOp o{true, (void *)buf, count * dt_size(dt), peer, c};
  if (g_group) g_ops.push_back(o); else do_send(o);
  return ncclSuccess;
}
ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t c, cudaStream_t) {
  PendingOp o{false, buf, count * dt_size(dt), peer, c};
  if (g_group) g_ops.push_back(o); else do_recv(o);
  return ncclSuccess;
}
ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t dt, ncclComm_t c, cudaStream_t) {
  const size_t b = count * dt_size(dt);
  if (b > kSlot) return ncclInvalidArgument;
  memcpy(c->gslot(c->rank), send, b);
  c->barrier();
  for (int r = 0; r < c->n; ++r) memcpy((char *)recv + (size_t)r * b, c->gslot(r), b);
  c->barrier();
  return ncclSuccess;
}
ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t c, cudaStream_t) {
  if (dt != ncclDouble || count * 8 > kSlot) return ncclInvalidArgument;
  memcpy(c->gslot(c->rank), send, count * 8);
  c->barrier();
  double *out = (double *)recv;
  for (size_t i = 0; i < count; ++i) {
    double v = ((const double *)c->gslot(0))[i];
    for (int r = 1; r < c->n; ++r) {
      const double w = ((const double *)c->gslot(r))[i];
      v = op == ncclSum ? v + w : op == ncclProd ? v * w : op == ncclMax ? (w > v ? w : v) : (w < v ? w : v);
    }
    out[i] = v;
  }
  c->barrier();
  return ncclSuccess;
}
ncclResult_t ncclBroadcast(const void *send, void *recv, size_t count, ncclDataType_t dt, int root, ncclComm_t c, cudaStream_t) {
  const size_t b = count * dt_size(dt);
  if (b > kSlot) return ncclInvalidArgument;
  if (c->rank == root) memcpy(c->gslot(root), send, b);
  c->barrier();
  memcpy(recv, c->gslot(root), b);
  c->barrier();
  return ncclSuccess;
}
const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "error (NCCL emulation)"; }
}

// see cuda_runtime.h: FCP_EMU_LIBM_ULP
#undef pow
#undef log
#undef exp
#undef tanh
#undef acos
#undef cos
namespace emu {
double libm_perturb(double x) {
  static const int k = getenv("FCP_EMU_LIBM_ULP") ? atoi(getenv("FCP_EMU_LIBM_ULP")) : 0;
  if (k <= 0 || x == 0.0 || x == 1.0 || x == -1.0 || !std::isfinite(x)) return x;
  unsigned long long b;
  memcpy(&b, &x, 8);
  b ^= b >> 33; b *= 0xff51afd7ed558ccdULL; b ^= b >> 33;
  const int steps = (int)(b % (unsigned)(2 * k + 1)) - k;     // -k .. +k ulp
  for (int i = 0; i < (steps < 0 ? -steps : steps); ++i) x = nextafter(x, steps < 0 ? -INFINITY : INFINITY);
  return x;
}
}   // namespace emu
