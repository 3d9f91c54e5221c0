// Micro-benchmark (development tool, not part of the library): latency of one flag-in-data hand-off between two SMs of a B200, the unit the
// barrier-free triangular sweeps are made of.  CTA 0 plays ping-pong with CTA k (k = 1..G-1, one CTA per SM) through two 16-byte LL words; the
// round trip / 2 is printed per partner SM, for relaxed GPU-scope and for volatile (system-scope) accesses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ll_pingpong tools/micro/ll_pingpong.cu && ./ll_pingpong
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

template <bool SYS> __device__ __forceinline__ void ll_load(const unsigned long long *p, unsigned long long &a, unsigned long long &b) {
  if (SYS) asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  else asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
template <bool SYS> __device__ __forceinline__ void ll_store(unsigned long long *p, unsigned long long a, unsigned long long b) {
  if (SYS) asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
  else asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <bool SYS>
__global__ void k_pingpong(unsigned long long *ping, unsigned long long *pong, int rounds, float *ns_out, int *smid_out, unsigned int base) {
  cg::grid_group grid = cg::this_grid();
  const int G = gridDim.x, b = blockIdx.x;
  if (threadIdx.x == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    smid_out[b] = (int)smid;
  }
  for (int k = 1; k < G; ++k) {
    grid.sync();
    if (threadIdx.x != 0) continue;
    const unsigned int s0 = base + (unsigned int)k * (unsigned int)(rounds + 1);
    if (b == 0) {
      unsigned long long a, c;
      const unsigned long long t0 = now();
      for (int r = 0; r < rounds; ++r) {
        const unsigned long long tag = (unsigned long long)(s0 + r) << 32;
        ll_store<SYS>(ping, tag | 1u, tag | 2u);
        do { ll_load<SYS>(pong, a, c); } while ((a >> 32) != (s0 + r) || (c >> 32) != (s0 + r));
      }
      ns_out[k] = (float)(now() - t0) / (2.0f * rounds);
    } else if (b == k) {
      unsigned long long a, c;
      for (int r = 0; r < rounds; ++r) {
        const unsigned long long tag = (unsigned long long)(s0 + r) << 32;
        do { ll_load<SYS>(ping, a, c); } while ((a >> 32) != (s0 + r) || (c >> 32) != (s0 + r));
        ll_store<SYS>(pong, tag | 3u, tag | 4u);
      }
    }
  }
}

template <bool SYS> static void run(const char *name, int G) {
  unsigned long long *ping, *pong;
  float *ns;
  int *smid;
  cudaMalloc(&ping, 256); cudaMalloc(&pong, 256); cudaMalloc(&ns, sizeof(float) * G); cudaMalloc(&smid, sizeof(int) * G);
  cudaMemset(ping, 0, 256); cudaMemset(pong, 0, 256);
  int rounds = 2000;
  unsigned int base = 1;
  void *args[] = {&ping, &pong, &rounds, &ns, &smid, &base};
  cudaLaunchCooperativeKernel((const void *)k_pingpong<SYS>, dim3(G), dim3(32), args, 0, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  float *h = (float *)malloc(sizeof(float) * G);
  int *hs = (int *)malloc(sizeof(int) * G);
  cudaMemcpy(h, ns, sizeof(float) * G, cudaMemcpyDeviceToHost);
  cudaMemcpy(hs, smid, sizeof(int) * G, cudaMemcpyDeviceToHost);
  float mn = 1e30f, mx = 0, sum = 0;
  for (int k = 1; k < G; ++k) { mn = h[k] < mn ? h[k] : mn; mx = h[k] > mx ? h[k] : mx; sum += h[k]; }
  printf("%s: one-way hand-off latency CTA0(sm %d) <-> CTA k: min %.0f ns  mean %.0f ns  max %.0f ns over %d partner SMs\n", name, hs[0], mn, sum / (G - 1), mx, G - 1);
  printf("  per partner (smid:ns):");
  for (int k = 1; k < G; ++k) printf(" %d:%.0f", hs[k], h[k]);
  printf("\n");
}

int main() {
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  run<false>("relaxed.gpu", nsm);
  run<true>("volatile (sys)", nsm);
  return 0;
}
