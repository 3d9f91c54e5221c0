// reduce.cuh -- the fixed reduction tree (see fcp_internal.h; CPU mirror: oracle/orc.cpp:orc_sum_tree)
#pragma once
#include "fcp_internal.h"

// Sum NS per-thread values over a 256-thread CTA: xor-butterfly inside each warp (offsets 16,8,4,2,1), then the 8
// warp sums are added in warp order by thread 0.  Result valid in thread 0 only.
template <int NS>
__device__ __forceinline__ void fcp_block_tree(double (&s)[NS], double (&out)[NS]) {
  __shared__ double ws[NS][FCP_TPB / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    double v = s[k];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) ws[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      double t = ws[k][0];
#pragma unroll
      for (int w = 1; w < FCP_TPB / 32; ++w) t = t + ws[k][w];
      out[k] = t;
    }
  }
  __syncthreads();   // ws may be reused by a second call
}

// Grid-level stage (a persistent CTA calls it once per chunk it owns: part = chunk index, nparts = number of chunks):
// every CTA stores its chunk partials, the last CTA to arrive (ticket counter) adds them in the
// fixed order (thread t takes partials t, t+256, ...; same block tree) and returns true in ALL its threads with the
// totals valid in thread 0.  partials: [NS][stride].  The counter is reset for the next kernel.
template <int NS>
__device__ __forceinline__ bool fcp_grid_reduce_part(double (&s)[NS], double *partials, int stride, unsigned int *counter, int part, int nparts,
                                                     double (&total)[NS]) {
  double blk[NS];
  fcp_block_tree<NS>(s, blk);
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) partials[(size_t)k * stride + part] = blk[k];
    __threadfence();
    unsigned int ticket = atomicAdd(counter, 1u);
    is_last = (ticket == (unsigned int)nparts - 1u);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  double acc[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    double t = 0.0;
    for (int i = threadIdx.x; i < nparts; i += FCP_TPB) t = t + __ldcg(&partials[(size_t)k * stride + i]);
    acc[k] = t;
  }
  fcp_block_tree<NS>(acc, total);
  if (threadIdx.x == 0) *counter = 0u;
  return true;
}
// one chunk per CTA: part = blockIdx.x, nparts = gridDim.x
template <int NS>
__device__ __forceinline__ bool fcp_grid_reduce(double (&s)[NS], double *partials, int stride, unsigned int *counter, double (&total)[NS]) {
  return fcp_grid_reduce_part<NS>(s, partials, stride, counter, (int)blockIdx.x, (int)gridDim.x, total);
}
