// How long does ONE softmax warp of attention.cu need for the exponential phase of a 64-key half block, alone on
// its SM sub-partition and with 1 / 3 sibling warps?  (SFU floor: 64 MUFU.EX2 x 8 clk = 512 clk per warp-block.)
// Variants: 0 = source order of the kernel (consumers right behind their MUFU pair, ptxas schedule),
//           1 = consumers deferred by DEPTH pairs in source, 2 = all 64 MUFU first, then all consumers (forced by
//           an opaque dependency), 3 = variant 0 plus FMNMX3 max pass + vote in front (the whole non-memory body).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_block_bench softmax_block_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t f32x2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void split(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ void unpack_h2(uint32_t v, float& lo, float& hi) {
  asm("{ .reg .f16 a, b; mov.b32 {a, b}, %2; cvt.f32.f16 %0, a; cvt.f32.f16 %1, b; }" : "=f"(lo), "=f"(hi) : "r"(v));
}
__device__ __forceinline__ float fmax3f(float a, float b, float c) { float y; asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }

template <int VAR, int DEPTH>
__global__ void __launch_bounds__(512, 1) k(const float* in, uint32_t* out, long long* clk, int rounds, int zero) {
  float s[64];
  for (int i = 0; i < 64; ++i) s[i] = in[(threadIdx.x * 64 + i) & 4095];
  uint32_t keep = 0;
  float mc = in[threadIdx.x & 63];
  const float c = 0.18033688f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < rounds; ++r) {
    uint64_t rs2[4] = {0, 0, 0, 0};
    uint32_t pk[32];
    const uint64_t c2 = f32x2(c, c), nmc2 = f32x2(-mc, -mc);
    if (VAR == 3) {
      float m4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
      for (int i = 0; i < 32; ++i) m4[i & 3] = fmax3f(m4[i & 3], s[i], s[32 + i]);
      const float mx = fmaxf(fmax3f(m4[0], m4[1], m4[2]), m4[3]);
      if (__any_sync(0xffffffffu, (mx - mc) * c > 8.f)) mc = mx;
    }
    if (VAR == 9 || VAR == 10) {  // max pass: FMNMX3, 4 chains
      float m4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
      for (int i = 0; i < 32; ++i) m4[i & 3] = fmax3f(m4[i & 3], s[i], s[32 + i]);
      const float mx = fmaxf(fmax3f(m4[0], m4[1], m4[2]), m4[3]);
      if (__any_sync(0xffffffffu, (mx - mc) * c > 8.f)) mc = mx;
#pragma unroll
      for (int i = 0; i < 32; ++i) {   // (s c - m c) in fp32, packed to f16x2, ONE MUFU per key pair; P is already packed
        float x0, x1;
        split(ffma2(f32x2(s[2 * i], s[2 * i + 1]), c2, nmc2), x0, x1);
        pk[i] = ex2h2(pack(x0, x1));
        if (VAR == 10) {               // fp32 row sum from the packed halves (instead of a ones-column in V)
          float e0, e1;
          unpack_h2(pk[i], e0, e1);
          rs2[i & 3] = fadd2(rs2[i & 3], f32x2(e0, e1));
        }
      }
    }
    if (VAR == 5 || VAR == 7) {  // max pass: FMNMX3, 4 chains (the kernels' current form)
      float m4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
      for (int i = 0; i < 32; ++i) m4[i & 3] = fmax3f(m4[i & 3], s[i], s[32 + i]);
      const float mx = fmaxf(fmax3f(m4[0], m4[1], m4[2]), m4[3]);
      if (__any_sync(0xffffffffu, (mx - mc) * c > 8.f)) mc = mx;
    }
    if (VAR == 6 || VAR == 8) {  // max pass: 2-input FMNMX, 8 chains
      float m8[8] = {-1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
      for (int i = 0; i < 64; ++i) m8[i & 7] = fmaxf(m8[i & 7], s[i]);
      const float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      if (__any_sync(0xffffffffu, (mx - mc) * c > 8.f)) mc = mx;
    }
    if (VAR == 5 || VAR == 6) {
      for (int i = 0; i < 32; ++i) pk[i] = __float_as_uint(s[i]) ^ __float_as_uint(mc);
    }
    if (VAR == 0 || VAR == 3) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x0, x1;
        split(ffma2(f32x2(s[2 * i], s[2 * i + 1]), c2, nmc2), x0, x1);
        const float e0 = ex2f(x0), e1 = ex2f(x1);
        rs2[i & 3] = fadd2(rs2[i & 3], f32x2(e0, e1));
        pk[i] = pack(e0, e1);
      }
    } else if (VAR == 1) {
      float e[64];
#pragma unroll
      for (int i = 0; i < 32 + DEPTH; ++i) {
        if (i < 32) {
          float x0, x1;
          split(ffma2(f32x2(s[2 * i], s[2 * i + 1]), c2, nmc2), x0, x1);
          e[2 * i] = ex2f(x0); e[2 * i + 1] = ex2f(x1);
        }
        if (i >= DEPTH) {
          const int q = i - DEPTH;
          rs2[q & 3] = fadd2(rs2[q & 3], f32x2(e[2 * q], e[2 * q + 1]));
          pk[q] = pack(e[2 * q], e[2 * q + 1]);
        }
      }
    } else if (VAR == 4 || VAR == 7 || VAR == 8) {
      // DEPTH of every 4 key pairs get their exp2 from the FMA/ALU pipes (Cody-Waite + degree-3 polynomial)
      const uint64_t magic2 = f32x2(12582912.f, 12582912.f), nmagic2 = f32x2(-12582912.f, -12582912.f);
      const uint64_t k3 = f32x2(0.0555041f, 0.0555041f), k2 = f32x2(0.2402265f, 0.2402265f),
                     k1 = f32x2(0.6931472f, 0.6931472f), k0 = f32x2(1.f, 1.f), m1 = f32x2(-1.f, -1.f);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x0, x1, e0, e1;
        split(ffma2(f32x2(s[2 * i], s[2 * i + 1]), c2, nmc2), x0, x1);
        if ((i & 3) < DEPTH) {
          x0 = fmaxf(x0, -125.f); x1 = fmaxf(x1, -125.f);
          const uint64_t x2 = f32x2(x0, x1);
          const uint64_t t2 = fadd2(x2, magic2);
          const uint64_t f2 = ffma2(fadd2(t2, nmagic2), m1, x2);
          uint64_t p2 = ffma2(k3, f2, k2);
          p2 = ffma2(p2, f2, k1);
          p2 = ffma2(p2, f2, k0);
          float p0, p1, t0f, t1f;
          split(p2, p0, p1); split(t2, t0f, t1f);
          e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0f) << 23));
          e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1f) << 23));
        } else {
          e0 = ex2f(x0); e1 = ex2f(x1);
        }
        rs2[i & 3] = fadd2(rs2[i & 3], f32x2(e0, e1));
        pk[i] = pack(e0, e1);
      }
    } else if (VAR == 2) {
      float e[64];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x0, x1;
        split(ffma2(f32x2(s[2 * i], s[2 * i + 1]), c2, nmc2), x0, x1);
        e[2 * i] = ex2f(x0); e[2 * i + 1] = ex2f(x1);
      }
      // opaque zero derived from the last results: every consumer depends on it
      const float z = __uint_as_float(__float_as_uint(e[62]) & __float_as_uint(e[63]) & (uint32_t)zero);
      const uint64_t z2 = f32x2(z, z);
#pragma unroll
      for (int i = 0; i < 4; ++i) rs2[i] = z2;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        rs2[i & 3] = fadd2(rs2[i & 3], f32x2(e[2 * i], e[2 * i + 1]));
        pk[i] = pack(e[2 * i], e[2 * i + 1]);
      }
    }
    float a, b;
    split(fadd2(fadd2(rs2[0], rs2[1]), fadd2(rs2[2], rs2[3])), a, b);
    mc += (a + b) * 1e-9f;
#pragma unroll
    for (int i = 0; i < 32; ++i) keep ^= pk[i];
    // new "scores" for the next round without memory traffic
#pragma unroll
    for (int i = 0; i < 64; i += 8) s[i] += 0.001f;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = keep;
  if ((threadIdx.x & 31) == 0) clk[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int VAR, int DEPTH>
void run(const char* name, const float* in, uint32_t* out, long long* clk) {
  const int rounds = 200;
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = warps_per_smsp * 4 * 32;
    k<VAR, DEPTH><<<148, threads>>>(in, out, clk, rounds, 0);
    k<VAR, DEPTH><<<148, threads>>>(in, out, clk, rounds, 0);
    cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, clk, sizeof(long long) * (threads / 32), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < threads / 32; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-44s %d warp(s)/SMSP: %7.0f clk per warp-block  (%5.0f clk per block of SFU time / SMSP)\n", name,
           warps_per_smsp, mx / rounds, mx / rounds / warps_per_smsp);
  }
}

int main() {
  float* in; uint32_t* out; long long* clk;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 148 * 16 * 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = -0.01f * (i % 977);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0, 0>("kernel order (ptxas schedule)", in, out, clk);
  run<1, 4>("consumers deferred 4 pairs in source", in, out, clk);
  run<1, 8>("consumers deferred 8 pairs in source", in, out, clk);
  run<2, 0>("all MUFU first, consumers after (forced)", in, out, clk);
  run<3, 0>("max pass + vote + kernel order", in, out, clk);
  run<4, 1>("1 of 4 pairs on the FMA pipe (poly3)", in, out, clk);
  run<4, 2>("2 of 4 pairs on the FMA pipe (poly3)", in, out, clk);
  run<5, 0>("max pass only: FMNMX3 x32, 4 chains", in, out, clk);
  run<6, 0>("max pass only: FMNMX x64, 8 chains", in, out, clk);
  run<7, 1>("FMNMX3 max + 1/4 poly (v6 body)", in, out, clk);
  run<8, 1>("FMNMX 8-chain max + 1/4 poly", in, out, clk);
  run<8, 2>("FMNMX 8-chain max + 2/4 poly", in, out, clk);
  run<9, 0>("FMNMX3 max + ex2.f16x2, no row sum", in, out, clk);
  run<10, 0>("FMNMX3 max + ex2.f16x2 + fp32 sum (unpack)", in, out, clk);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
