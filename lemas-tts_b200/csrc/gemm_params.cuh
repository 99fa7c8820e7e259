// Kernel-side parameters and activation helpers shared by the GEMM kernels (gemm.cu: single-CTA tiles with taps /
// groups; gemm2.cu: CTA-pair 256x256 tiles).
#pragma once
#include "ptx.cuh"

namespace lemas {

struct GemmParams {
  int batches, rows;      // A tiling: tiles never straddle a batch item
  int n, k_iters, kc_per_tap, tap_pad, w_tap_stride, group_cols, tap_dil;
  const float* bias;
  __half* out16; int ld16;
  float* out32; int ld32;
  const float* resid; int ldr;
  const float* gate; int gate_bstride;
  const int* row_valid;
  const int* row_limit;   // gemm2 only: tiles starting at or beyond row_limit[batch item] are skipped
  int red_add;            // gemm2 gated-residual epilogue, in place (resid == out32): x += gate * (acc + bias) as one
                          // vector reduction in L2 (red.global.add.v4.f32) instead of load + add + store
  // LayerNorm folded into the surrounding GEMMs (gemm2 only; see lemas_gemm_desc)
  const float* ln_scale; __half* ln_out16; int ln_ld16; float* ln_stats;
  const float* ln_stats_in; int ln_parts; const float* ln_uv; const int* ln_step; float ln_inv_k;
  int seq_len;
  const float2* rope; int rope_cols; int inner;
  __half* vt; int vt_ld;
};

DEVI float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh(u)),  u = sqrt(2/pi) (x + 0.044715 x^3); one SFU op (tanh.approx, rel. error 2^-11 — the output
  // is rounded to fp16 right after).  The ex2 + rcp form costs two SFU ops per element and made the FF1 epilogue
  // SFU-bound (9 M elements x 2 / (148 SMs x 16/clk) = 3.9 us per launch).
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
DEVI float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f)); }
DEVI float mish_f(float x) {
  // x * tanh(softplus(x)); tanh(log(1+e^x)) = (n^2 + 2n) / (n^2 + 2n + 2), n = e^x  (softplus threshold 20 as torch)
  if (x > 20.0f) return x;
  float n = __expf(x);
  float a = n * (n + 2.0f);
  return x * __fdividef(a, a + 2.0f);
}

}  // namespace lemas
