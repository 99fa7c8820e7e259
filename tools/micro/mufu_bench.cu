// Throughput of MUFU-based exp2 flavours on sm_100a: f32 vs packed f16x2 / bf16x2, and a FMA-pipe polynomial.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu ; run on the GPU box
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xB800B400u + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 5) {  // Cody-Waite + degree-3 polynomial on the FMA/ALU pipes
        float x = a[i];
        float t = x + 12582912.0f;
        float f = x - (t - 12582912.0f);
        float p = fmaf(fmaf(fmaf(0.0555041f, f, 0.2402265f), f, 0.6931472f), f, 1.0f);
        a[i] = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23)) - 1.5f;
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_instr) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 1024>>>(out, 16);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 1024>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double instr = 148.0 * 8 * 1024 * iters * 8;
  double per_clk_sm = instr / (ms * 1e-3) / 148 / 1.965e9;
  printf("%-28s %8.3f ms  %6.2f thread-instr/clk/SM  -> %6.2f results/clk/SM (at 1965 MHz)\n", name, ms, per_clk_sm,
         per_clk_sm * per_instr);
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("tanh.approx.f32", 1);
  run<4>("tanh.approx.f16x2", 2);
  run<5>("poly3 exp2 (FMA pipe)", 1);
  return 0;
}
