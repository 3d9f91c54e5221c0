// tests/emu/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
//
// A CPU *emulation* of the small part of the CUDA programming model that csrc/*.cu uses, so that the logic of the
// product kernels (the same source files, re-compiled by g++ through tests/emu/build.py) can be desk-checked against the
// oracle in a container without a GPU.  It is never built by __graft_entry__.build(), never loaded by the product
// (freecappuccino-dev_b200/lib.py knows nothing about it) and proves nothing about performance: only `tests/` loads
// libfcp_emu.so, and only when FCP_TEST_EMU=1 is set or from tests/test_emu_*.py.
//
// Execution model: a launch runs its blocks one after the other on the calling OS thread; the threads of a block are
// fibers (hand-written context switch) scheduled round-robin; __syncthreads / *_sync warp primitives block a fiber until
// its block / warp has arrived.  A cooperative launch gives every block its own OS thread (grid.sync = barrier across
// them); `__shared__` is `static thread_local`, so each block sees its own copy in both cases.  "Device" memory lives in
// one process-wide shared-memory arena so that cudaIpc*MemHandle can map it into another process (multi-rank tests).
#pragma once
#define FCP_EMU 1
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static thread_local
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned int x, y, z; };
struct dim3 {
  unsigned int x, y, z;
  dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
static const int warpSize = 32;

// ---- runtime API ------------------------------------------------------------------------------------------------
enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorInvalidConfiguration = 9, cudaErrorNoDevice = 100, cudaErrorNotReady = 600, cudaErrorUnknown = 999 };
struct emuStream;
struct emuEvent;
typedef emuStream *cudaStream_t;
typedef emuEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamDefault = 0, cudaStreamNonBlocking = 1 };
enum { cudaEventDefault = 0, cudaEventBlockingSync = 1, cudaEventDisableTiming = 2 };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrComputeCapabilityMajor = 75, cudaDevAttrComputeCapabilityMinor = 76, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
  char name[256];
  size_t totalGlobalMem;
  int major, minor, multiProcessorCount, cooperativeLaunch;
  size_t sharedMemPerBlockOptin;
};
struct cudaIpcMemHandle_t { char reserved[64]; };

cudaError_t emuMalloc(void **p, size_t bytes);
template <class T> inline cudaError_t cudaMalloc(T **p, size_t bytes) { return emuMalloc((void **)p, bytes); }
cudaError_t cudaFree(void *p);
cudaError_t emuMallocHost(void **p, size_t bytes);
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return emuMallocHost((void **)p, bytes); }
template <class T> inline cudaError_t cudaHostAlloc(T **p, size_t bytes, unsigned) { return emuMallocHost((void **)p, bytes); }
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st = nullptr);
cudaError_t cudaMemset(void *dst, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void *dst, int v, size_t bytes, cudaStream_t st = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t *st);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t st);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventQuery(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaSetDevice(int dev);
cudaError_t cudaGetDevice(int *dev);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int dev);
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int dev);
cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
template <class K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 1; return cudaSuccess; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
enum cudaLimit { cudaLimitPersistingL2CacheSize = 6 };
inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }

// ---- launches ---------------------------------------------------------------------------------------------------
namespace emu {
struct Cfg {
  dim3 g, b;
  size_t smem;
  cudaStream_t st;
  Cfg(dim3 g_, dim3 b_, size_t smem_ = 0, cudaStream_t st_ = nullptr) : g(g_), b(b_), smem(smem_), st(st_) {}
};
void launch_impl(const Cfg &c, const std::function<void()> &body, bool cooperative);
// kernel<<<cfg>>>(args)  ->  emu::launch(emu::Cfg(cfg), kernel, args)   (tests/emu/build.py rewrites the launch syntax)
template <class... KA, class... A> inline void launch(const Cfg &c, void (*k)(KA...), A &&...a) {
  std::tuple<typename std::decay<KA>::type...> vals(std::forward<A>(a)...);   // arguments are evaluated once, like a real launch
  launch_impl(c, [&]() { std::apply(k, vals); }, false);
}
template <class... KA, class... A> inline void launch_coop(const Cfg &c, void (*k)(KA...), A &&...a) {
  std::tuple<typename std::decay<KA>::type...> vals(std::forward<A>(a)...);
  launch_impl(c, [&]() { std::apply(k, vals); }, true);
}
void sync_block();
void sync_grid();
void warp_exchange(unsigned mask, const void *mine, void *out, int src_lane, size_t bytes);   // out <- value of src_lane (own if inactive)
unsigned warp_ballot(unsigned mask, int pred);
unsigned long long now_ns();
void yield();            // let the other fibers of the block run (spin loops on something a sibling thread produces)
void os_yield();         // spin loops on something another PROCESS produces
unsigned char *dyn_smem();   // the dynamic shared memory of the running block
}   // namespace emu

// ---- device intrinsics ------------------------------------------------------------------------------------------
inline void __syncthreads() { emu::sync_block(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_ballot(mask, 0); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> inline T emu_shfl(unsigned mask, T v, int src) { T o; emu::warp_exchange(mask, &v, &o, src, sizeof(T)); return o; }
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) { const int l = threadIdx.x & 31; return emu_shfl(mask, v, (l & ~(width - 1)) | (src & (width - 1))); }
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lm, int width = 32) { const int l = threadIdx.x & 31; const int s = l ^ lm; return emu_shfl(mask, v, (s & ~(width - 1)) == (l & ~(width - 1)) ? s : l); }
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) { const int l = threadIdx.x & 31; const int s = l + (int)d; return emu_shfl(mask, v, (s & ~(width - 1)) == (l & ~(width - 1)) ? s : l); }
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) { const int l = threadIdx.x & 31; const int s = l - (int)d; return emu_shfl(mask, v, (s >= 0 && (s & ~(width - 1)) == (l & ~(width - 1))) ? s : l); }
inline unsigned __ballot_sync(unsigned mask, int pred) { return emu::warp_ballot(mask, pred); }
inline int __any_sync(unsigned mask, int pred) { return emu::warp_ballot(mask, pred) != 0u; }
inline int __all_sync(unsigned mask, int pred) { return emu::warp_ballot(mask, !pred) == 0u; }
inline unsigned __activemask() { return 0xffffffffu; }

template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcs(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *(const volatile T *)p; }
template <class T> inline T __ldca(const T *p) { return *p; }
template <class T> inline T __ldlu(const T *p) { return *p; }
template <class T> inline T __ldcv(const T *p) { return *(const volatile T *)p; }
template <class T, class U> inline void __stcs(T *p, U v) { *p = (T)v; }
template <class T, class U> inline void __stcg(T *p, U v) { *p = (T)v; }
template <class T, class U> inline void __stwt(T *p, U v) { *p = (T)v; }

inline long long __double_as_longlong(double v) { long long o; memcpy(&o, &v, 8); return o; }
inline double __longlong_as_double(long long v) { double o; memcpy(&o, &v, 8); return o; }
inline int __double2hiint(double v) { return (int)(__double_as_longlong(v) >> 32); }
inline int __double2loint(double v) { return (int)(__double_as_longlong(v) & 0xffffffffll); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }

template <class T> inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline double atomicAdd(double *p, double v) {
  unsigned long long *q = (unsigned long long *)p, old = __atomic_load_n(q, __ATOMIC_SEQ_CST), nw;
  double o;
  do { memcpy(&o, &old, 8); const double s = o + v; memcpy(&nw, &s, 8); } while (!__atomic_compare_exchange_n(q, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return o;
}
template <class T> inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicCAS(T *p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
template <class T> inline T atomicMax(T *p, T v) { T old = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return old; }
template <class T> inline T atomicMin(T *p, T v) { T old = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return old; }

// CUDA's global-namespace min/max overloads
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }
inline double min(double a, double b) { return fmin(a, b); }
inline double max(double a, double b) { return fmax(a, b); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }

// FCP_EMU_LIBM_ULP=k: the device libm (pow, log, tanh, acos, cos, exp) need not round like glibc (CUDA documents 1-2 ulp for these in
// double precision).  With the variable set every transcendental result is moved by up to k ulp in a direction picked from a hash of its
// bits, so that the parity tests show which of their tolerances would not survive a libm that differs from the oracle's.
#include <cmath>
#include <complex>
#include <map>
#include <string>
#include <vector>
namespace emu { double libm_perturb(double x); }
inline double emu_pow(double a, double b) { return emu::libm_perturb(::pow(a, b)); }
inline double emu_log(double a) { return emu::libm_perturb(::log(a)); }
inline double emu_exp(double a) { return emu::libm_perturb(::exp(a)); }
inline double emu_tanh(double a) { return emu::libm_perturb(::tanh(a)); }
inline double emu_acos(double a) { return emu::libm_perturb(::acos(a)); }
inline double emu_cos(double a) { return emu::libm_perturb(::cos(a)); }
#define pow(a, b) emu_pow(a, b)
#define log(a) emu_log(a)
#define exp(a) emu_exp(a)
#define tanh(a) emu_tanh(a)
#define acos(a) emu_acos(a)
#define cos(a) emu_cos(a)
