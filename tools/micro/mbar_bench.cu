// Latency of mbarrier waits on sm_100a: (a) try_wait / test_wait on an already completed phase, (b) wake-up latency of
// a waiter after the last arrival (try_wait suspended vs test_wait spin), (c) plain ld.shared for scale.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ bool test_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{.reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
  return ok;
}
__global__ void k(long long* out) {
  __shared__ uint64_t bar[4];
  __shared__ volatile long long t_arrive[2];
  __shared__ volatile int flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + i)));
    flag = 0;
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar + 0)) : "memory");
    __syncwarp();
    long long a = clock64();
    bool ok = try_wait(bar + 0, 0);
    long long b = clock64();
    bool ok2 = test_wait(bar + 0, 0);
    long long c = clock64();
    int v = flag;
    long long d = clock64();
    if (lane == 0) { out[0] = b - a; out[1] = c - b; out[2] = d - c + (v & ok & ok2 ? 0 : 0); }
  }
  __syncthreads();
  // wake-up latency: warp 1 waits (try_wait), warp 2 waits (test_wait spin); warp 3 arrives on both after a delay
  if (warp == 1) {
    while (!try_wait(bar + 1, 0)) {}
    long long w = clock64();
    if (lane == 0) out[3] = w - t_arrive[0];
  } else if (warp == 2) {
    while (!test_wait(bar + 2, 0)) {}
    long long w = clock64();
    if (lane == 0) out[4] = w - t_arrive[1];
  } else if (warp == 3) {
    long long s = clock64();
    while (clock64() - s < 20000) {}
    if (lane == 0) {
      t_arrive[0] = clock64();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar + 1)) : "memory");
      long long s2 = clock64();
      while (clock64() - s2 < 5000) {}
      t_arrive[1] = clock64();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar + 2)) : "memory");
    }
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  for (int it = 0; it < 3; ++it) { k<<<1, 128>>>(d); cudaDeviceSynchronize(); }
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("satisfied try_wait %lld clk, satisfied test_wait %lld clk, ld.shared %lld clk\n", h[0], h[1], h[2]);
  printf("wake-up after arrive: try_wait (suspended) %lld clk, test_wait (spinning) %lld clk\n", h[3], h[4]);
  return 0;
}
