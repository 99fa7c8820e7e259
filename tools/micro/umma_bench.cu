// How long does the tensor pipe of one SM need per tcgen05.mma at the shapes the attention kernels issue?
// One CTA on one SM, one issuing thread (optionally a second warp issuing a second stream), operands resident in
// shared memory / TMEM (zero-filled: timing only).  Prints clocks per MMA instruction and the fraction of the
// nominal rate (M128: N/8 clocks per K16 instruction = 4096 MAC/clk/SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lemas-tts_b200/csrc -o tools/micro/umma_bench tools/micro/umma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace lemas;

DEVI void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// mode: 0 = SS, chains of 4 K-steps into ONE accumulator (every MMA depends on the previous one)
//       1 = SS, chains of 4, accumulators rotate over 4 TMEM regions (consecutive chains independent)
//       2 = TS (A from TMEM), chains of 4 into one accumulator
//       3 = TS, chains rotate over 4 accumulators
//       4 = attention pattern: [TS chain -> O_x, SS chain -> S_x] for x = 0..3 (4 pipelines), all independent regions
//       5 = SS, K-steps of ONE chain interleaved over 2 accumulators (a0 k0, a1 k0, a0 k1, a1 k1 ...)
template <int N>
__global__ void __launch_bounds__(128, 1) bench(int mode, int reps, int two_issuers, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t idesc = umma_idesc_f16(128, N);
  const uint64_t adesc = umma_desc_sw128(smem_u32(smem)), bdesc = umma_desc_sw128(smem_u32(smem + 16384));
  if ((warp == 0 || (two_issuers && warp == 1)) && (threadIdx.x & 31) == 0) {
    const uint32_t base = tm + (warp == 1 ? 256 : 0);   // second issuer works in the upper half of TMEM
    const int regions = (two_issuers || N > 64) ? (N > 128 ? 1 : 2) : 4;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (mode == 4) {
        const int x = r % (two_issuers ? 2 : 4);
        const uint32_t s_x = base + x * 64 * (two_issuers ? 1 : 1), o_x = base + (two_issuers ? 128 : 0) + x * 64;
        // O_x += P_x V (TS, N = 64), then S_x = Q K^T (SS, N = 64)
        for (int k = 0; k < 4; ++k) umma_ts((two_issuers ? o_x : tm + 256 + x * 64), s_x + 8 * k, bdesc + 2 * k, umma_idesc_f16(128, 64), 1);
        for (int k = 0; k < 4; ++k) umma_f16_ss(s_x, adesc + 2 * k, bdesc + 2 * k, umma_idesc_f16(128, 64), k != 0);
      } else if (mode == 5) {
        for (int k = 0; k < 4; ++k) {
          umma_f16_ss(base, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
          umma_f16_ss(base + N, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
        }
      } else {
        const bool ts = mode == 2 || mode == 3;
        const bool rot = mode == 1 || mode == 3;
        // TS: A occupies columns [448, 480) (32 columns = 64 fp16 per row), accumulators below
        const uint32_t d = base + (rot ? (r % regions) * (N > 64 ? N : 64) : 0);
        for (int k = 0; k < 4; ++k) {
          if (ts) umma_ts(d, tm + 448 + 8 * k, bdesc + 2 * k, idesc, k != 0);
          else umma_f16_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
        }
      }
    }
    const long long t_issued = clock64();
    umma_commit(bar + warp);
    mbar_wait(bar + warp, 0);
    const long long t1 = clock64();
    out[warp * 2] = t1 - t0;
    out[warp * 2 + 1] = t_issued - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

// `issuers` warps, each the single issuer of its own pipeline (S_x / P_x at column 128 w, O_x at 128 w + 64), all
// issuing the attention pattern [O_x += P_x V (TS N64 x4); S_x = Q K^T (SS N64 x4)] concurrently.
// s_n = 64: per-half S (8 MMAs per round);  s_n = 128: one N = 128 S per round (TS x4 + TS x4 into two O, SS N128 x4)
__global__ void __launch_bounds__(128, 1) bench_multi(int issuers, int s_n, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(bar + i, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint64_t adesc = umma_desc_sw128(smem_u32(smem)), bdesc = umma_desc_sw128(smem_u32(smem + 16384));
  if (warp < issuers && (threadIdx.x & 31) == 0) {
    const uint32_t base = tm + warp * 128;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (s_n == 64) {
        for (int k = 0; k < 4; ++k) umma_ts(base + 64, base + 8 * k, bdesc + 2 * k, umma_idesc_f16(128, 64), 1);
        for (int k = 0; k < 4; ++k) umma_f16_ss(base, adesc + 2 * k, bdesc + 2 * k, umma_idesc_f16(128, 64), k != 0);
      } else {  // one tile per issuer in 256 columns: S 128 | O_A 64 | O_B 64 (issuers <= 2)
        const uint32_t b2 = tm + warp * 256;
        for (int k = 0; k < 4; ++k) umma_ts(b2 + 128, b2 + 8 * k, bdesc + 2 * k, umma_idesc_f16(128, 64), 1);
        for (int k = 0; k < 4; ++k) umma_ts(b2 + 192, b2 + 32 + 8 * k, bdesc + 2 * k, umma_idesc_f16(128, 64), 1);
        for (int k = 0; k < 4; ++k) umma_f16_ss(b2, adesc + 2 * k, bdesc + 2 * k, umma_idesc_f16(128, 128), k != 0);
      }
    }
    umma_commit(bar + warp);
    mbar_wait(bar + warp, 0);
    out[warp] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

void run_multi(int issuers, int s_n, long long* d_out) {
  const int reps = 2000;
  cudaFuncSetAttribute(bench_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  long long h[4] = {0, 0, 0, 0};
  for (int it = 0; it < 2; ++it) {
    cudaMemset(d_out, 0, sizeof(h));
    bench_multi<<<1, 128, 65536>>>(issuers, s_n, reps, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("multi: %s\n", cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < issuers; ++i) mx = h[i] > mx ? h[i] : mx;
  // work per round and issuer: s_n 64: 64 queries-keys half block = 128x64x64 x2 MAC; s_n 128: a whole 128x128 block
  const double mac = (s_n == 64 ? 2.0 * 128 * 64 * 64 : 2.0 * 128 * 128 * 64) * reps * issuers;
  printf("attention pattern, S N=%3d, %d issuer(s): %7.1f clk per round per issuer, %5.1f clk/MMA overall, %4.0f %% of the tensor peak\n",
         s_n, issuers, (double)mx / reps, (double)mx / reps / ((s_n == 64 ? 8 : 12) * issuers), 100.0 * mac / 4096.0 / mx);
}

template <int N>
void run(const char* name, int mode, int two, long long* d_out) {
  const int reps = 2000;
  cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  long long h[4] = {0, 0, 0, 0};
  for (int it = 0; it < 2; ++it) {
    cudaMemset(d_out, 0, sizeof(h));
    bench<N><<<1, 128, 65536>>>(mode, reps, two, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const int per_rep = (mode == 4 || mode == 5) ? 8 : 4;
  const double n_mma = (double)reps * per_rep;
  const double ideal = N / 2.0;  // clocks per M128 x N x K16 instruction at 4096 MAC/clk
  printf("%-58s %7.1f clk/MMA (issue %6.1f)  ideal %5.1f  -> %4.0f %% of peak", name, h[0] / n_mma, h[1] / n_mma,
         mode == 4 ? 32.0 : ideal, 100.0 * (mode == 4 ? 32.0 : ideal) / (h[0] / n_mma) * (two ? 2 : 1));
  if (two) printf("   [issuer 2: %7.1f clk/MMA]", h[2] / n_mma);
  printf("\n");
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<64>("SS M128 N64  K16 x4 chains, one accumulator", 0, 0, d_out);
  run<64>("SS M128 N64  K16 x4 chains, rotating accumulators", 1, 0, d_out);
  run<64>("SS M128 N64  two chains interleaved per K step", 5, 0, d_out);
  run<64>("TS M128 N64  K16 x4 chains, one accumulator", 2, 0, d_out);
  run<64>("TS M128 N64  K16 x4 chains, rotating accumulators", 3, 0, d_out);
  run<64>("attention pattern (TS PV + SS S, N64), 4 pipelines, 1 issuer", 4, 0, d_out);
  run<64>("attention pattern, 2 issuers x 2 pipelines", 4, 1, d_out);
  run<64>("SS M128 N64  one accumulator, 2 issuers", 0, 1, d_out);
  run<128>("SS M128 N128 K16 x4 chains, one accumulator", 0, 0, d_out);
  run<128>("SS M128 N128 K16 x4 chains, rotating accumulators", 1, 0, d_out);
  run<128>("TS M128 N128 K16 x4 chains, one accumulator", 2, 0, d_out);
  run<256>("SS M128 N256 K16 x4 chains, one accumulator", 0, 0, d_out);
  run<256>("TS M128 N256 K16 x4 chains, one accumulator", 2, 0, d_out);
  for (int n : {1, 2, 3, 4}) run_multi(n, 64, d_out);
  for (int n : {1, 2}) run_multi(n, 128, d_out);
  return 0;
}
