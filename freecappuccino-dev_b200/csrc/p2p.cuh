// p2p.cuh -- device-side primitives of the peer-memory protocol (fcp_internal.h: WinHeader / CommDev).
// Producer: data stores into the peer window -> __threadfence_system() -> st.release.sys of the sequence flag.
// Consumer: ld.acquire.sys spin on its own (local) flag -> data loads that bypass L1 (ld.volatile / __ldcg).
#pragma once
#include "fcp_internal.h"

#ifdef FCP_EMU   // tests/emu: the same primitives for the CPU emulation (ranks are processes sharing the window memory; test infrastructure only)
__device__ __forceinline__ unsigned long long p2p_ld_acquire(const unsigned long long *p) { emu::os_yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void p2p_st_release(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ double p2p_ld_data(const double *p) { return *(const volatile double *)p; }
__device__ __forceinline__ unsigned long long p2p_now_ns() { return emu::now_ns(); }
__device__ __forceinline__ void p2p_ll_store_words(unsigned long long *dst, unsigned long long w0, unsigned long long w1) {
  __atomic_store_n(dst, w0, __ATOMIC_RELAXED);
  __atomic_store_n(dst + 1, w1, __ATOMIC_RELAXED);
}
__device__ __forceinline__ void p2p_ll_load_words(const unsigned long long *src, unsigned long long &w0, unsigned long long &w1) {
  w0 = __atomic_load_n(src, __ATOMIC_RELAXED);
  w1 = __atomic_load_n(src + 1, __ATOMIC_RELAXED);
}
#else
__device__ __forceinline__ unsigned long long p2p_ld_acquire(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void p2p_st_release(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double p2p_ld_data(const double *p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long p2p_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void p2p_ll_store_words(unsigned long long *dst /* 16-byte aligned */, unsigned long long w0, unsigned long long w1) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void p2p_ll_load_words(const unsigned long long *src, unsigned long long &w0, unsigned long long &w1) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
}
#endif
// ---- flag-in-data words ("LL" protocol): a double travels as two 8-byte words {32 data bits | 32-bit sequence number}.
// An 8-byte store is atomic, so a word whose upper half equals the expected sequence number carries valid data: no
// fence, no separate flag, one NVLink one-way latency.  16-byte aligned pairs move as one v2 transaction.
__device__ __forceinline__ void p2p_ll_store(unsigned long long *dst /* 16-byte aligned */, double v, unsigned int seq) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (bits & 0xffffffffull) | ((unsigned long long)seq << 32);
  const unsigned long long w1 = (bits >> 32) | ((unsigned long long)seq << 32);
  p2p_ll_store_words(dst, w0, w1);
}
__device__ __forceinline__ double p2p_ll_load(const unsigned long long *src, unsigned int seq, WinHeader *hdr) {
  unsigned long long w0, w1, t0 = 0;
  unsigned int n = 0;
  // The word lives in THIS rank's window (the peer wrote it over NVLink into this GPU's memory, whose point of coherence is this GPU's L2): a
  // relaxed GPU-scope load sees it as soon as the system-scope one does; whether it is cheaper is an A/B switch (FCP_P2P_POLL=gpu), the tag check
  // makes a stale read harmless (the loop just polls again).
#ifndef FCP_EMU
  const bool gpu_scope = hdr->poll_gpu_scope != 0u;
#endif
  for (;;) {
#ifdef FCP_EMU
    p2p_ll_load_words(src, w0, w1);
#else
    if (gpu_scope) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
    else p2p_ll_load_words(src, w0, w1);
#endif
    if ((unsigned int)(w0 >> 32) == seq && (unsigned int)(w1 >> 32) == seq) break;
    if ((++n & 4095u) == 0u) {
      if (*(volatile int *)&hdr->error) break;     // another wait already gave up: fall through fast, the host reports the error
      const unsigned long long t = p2p_now_ns();
      if (!t0) t0 = t;
      else if (t - t0 > hdr->timeout_ns) { hdr->error = 1; break; }
    }
  }
  return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}

// spin until *flag >= seq; gives up after hdr->timeout_ns (a peer died or a protocol bug) and raises hdr->error instead of hanging the GPU
__device__ __forceinline__ void p2p_wait(const unsigned long long *flag, unsigned long long seq, WinHeader *hdr) {
  unsigned long long t0 = 0;
  unsigned int n = 0;
  while (p2p_ld_acquire(flag) < seq) {
    if ((++n & 4095u) == 0u) {
      if (*(volatile int *)&hdr->error) break;
      const unsigned long long t = p2p_now_ns();
      if (!t0) t0 = t;
      else if (t - t0 > hdr->timeout_ns) { hdr->error = 1; break; }
    }
  }
}
