// BigVGAN-v2 generator (the reference's `mel_spec_type: bigvgan` vocoder branch, utils_infer.py:144-158, :550-551;
// SURVEY.md §8 row f3): mel [B, num_mels, T] -> waveform [B, T * prod(rates)].
//
// The generator itself lives in an un-vendored submodule of the reference (third_party/BigVGAN, NVIDIA/BigVGAN, model
// nvidia/bigvgan_v2_24khz_100band_256x); its published algorithm is restated in oracle/bigvgan_oracle.py, which this
// file is checked against (parity unpinned).
//
// Layout: every activation is TIME-major, [B, T, Cpad] with channels contiguous and Cpad = channels rounded up to a
// multiple of 64 (padded channels carry exact zeros end to end: zero weights, zero bias, snake(0) = 0).  That makes
//   * every Conv1d (k taps, dilation d) an im2col-free tap GEMM over rows (lemas_gemm_desc.taps / tap_dilation, TMA
//     zero fill outside the sequence = the conv's zero padding), tcgen05 tensor cores, fp16 operands / fp32 accumulate;
//   * every ConvTranspose1d (k = 2r, stride r, pad r/2) a 3-tap GEMM with N = r * Cpad_out: output row m, column block
//     ph is output sample m*r + ph, which receives input rows {m, m-1} (ph + r/2 < r) or {m+1, m} (otherwise); the
//     unused (tap, phase) weight blocks are zero.  [B, T, r*Cpad] IS [B, T*r, Cpad] — no scatter, no trimming;
//   * the anti-aliased SnakeBeta activation (2x Kaiser-sinc up-sampling, snake, 2x down-sampling, replicate padding)
//     one fused kernel per call: a tile of x rows is staged in shared memory, the 2x-rate snake samples are formed
//     there once, and the 12-tap decimation reads them back — the 2x-rate signal never exists in HBM.
// The residual stream of each AMP block is fp32 (gated-residual GEMM epilogue with gate = 1); GEMM operands are fp16.
#include "common.h"
#include "ptx.cuh"

namespace lemas {

int gemm_launch(const lemas_gemm_desc& d, cudaStream_t stream);

constexpr int AA_L = 30;           // outputs per thread (one channel, consecutive rows); AA_L + 6 is a multiple of 6
constexpr int AA_CH = 64;          // channels per block (consecutive threads = consecutive channels: coalesced rows)
constexpr int AA_THREADS = 256;    // 4 time chunks x 64 channels

struct AaFilter { float f[12]; };

template <typename T> DEVI float ld_as_float(const T* p);
template <> DEVI float ld_as_float<float>(const float* p) { return *p; }
template <> DEVI float ld_as_float<__half>(const __half* p) { return __half2float(*p); }

// SnakeBeta with precomputed e^alpha (ea) and 1 / (e^beta + 1e-9) (ib):  u + ib sin^2(u ea).
// sin^2(th) = (1 - cos 2 th) / 2 with 2 th reduced to [-pi, pi] in turns (one MUFU.COS; absolute error ~1e-6 for the
// |th| < 100 these activations see — the result is rounded to fp16 two steps later).
DEVI float snake_beta(float u, float ea, float ib) {
  float turns = u * ea * 0.318309886183790672f;
  turns -= rintf(turns);
  return fmaf(ib, fmaf(-0.5f, __cosf(turns * 6.283185307179586f), 0.5f), u);
}

// y = DownSample2(SnakeBeta(UpSample2(x)))   (Activation1d).  x = (x0 [+ x1 + x2]) * in_scale, [B, T, ld] of TIn;
// ab: fp32 [2, ld]: row 0 = e^alpha, row 1 = 1 / (e^beta + 1e-9);  out: fp16 [B, T, ld].
//   u[n] = 2 sum_j x[clamp(j)] f[n + 5 - 2 j]            (6 taps: j in [a-3, a+2] for n = 2a, [a-2, a+3] for n = 2a+1)
//   s[n] = snake_beta(u[n])
//   y[t] = sum_k s[clamp(2 t + k - 5, 0, 2T-1)] f[k]
// A block stages AA_ROWS + 12 input rows x 64 channels in shared memory (coalesced, index-clamped = the replicate
// padding of x); then one thread = one channel x AA_L consecutive outputs: the 2x-rate snake samples slide through a
// 12-deep register window and every output is formed the moment its last sample exists — the 2x-rate signal never
// touches shared or global memory.  Chunks in which the 2x-rate signal itself is replicate-padded (sequence ends) take
// a clamped-index path.
constexpr int AA_ROWS = AA_L * (AA_THREADS / AA_CH);   // 128 output rows per block

template <typename TIn, int NIN>
__global__ void __launch_bounds__(AA_THREADS)
snake_aa_kernel(const TIn* __restrict__ x0, const TIn* __restrict__ x1, const TIn* __restrict__ x2, float in_scale,
                const float* __restrict__ ab, __half* __restrict__ out, int T, int ld, AaFilter flt) {
  __shared__ float xs[AA_ROWS + 12][AA_CH];    // xs[r] = x[clamp(tile0 - 6 + r)]
  const int c = threadIdx.x & (AA_CH - 1);
  const int chunk = threadIdx.x / AA_CH;
  const int ch = blockIdx.y * AA_CH + c;
  const int tile0 = blockIdx.x * AA_ROWS;
  const long base = (long)blockIdx.z * T * ld + ch;
  for (int r = chunk; r < AA_ROWS + 12; r += AA_THREADS / AA_CH) {
    int j = tile0 - 6 + r;
    j = j < 0 ? 0 : (j > T - 1 ? T - 1 : j);
    const long off = base + (long)j * ld;
    float v = ld_as_float(x0 + off);
    if constexpr (NIN == 3) v = (v + ld_as_float(x1 + off) + ld_as_float(x2 + off)) * in_scale;
    xs[r][c] = v;
  }
  const float ea = ab[ch], ib = ab[ld + ch];
  __syncthreads();
  const int t0 = tile0 + chunk * AA_L;
  if (t0 >= T) return;
  if (ea == 0.f) {   // padded channel (e^alpha > 0 for every real one): snake(0) = 0, nothing to compute
    const int t_end = min(T, t0 + AA_L);
    for (int t = t0; t < t_end; ++t) out[base + (long)t * ld] = __float2half_rn(0.f);
    return;
  }
  const int r0 = chunk * AA_L;                 // xs row of x[t0 - 6]
  if (t0 >= 3 && t0 + AA_L + 2 <= T - 1) {     // every 2x-rate index 2 t0 - 5 .. 2 t0 + 2 L + 4 lies inside [0, 2T - 1]
    // Iteration i (a = t0 - 3 + i, x[a - 3] = xs[r0 + i]) forms s[2a], then the output t = a - 3 from the 12 latest
    // samples, then s[2a + 1].  Six iterations push 12 samples = one full turn of the window, so inside a group of
    // six every window slot is a compile-time register (no shifting); the groups run in a rolled loop.  The first
    // even sample (i = 0) and the last odd one are not needed — pushing them anyway is harmless (they fall out of /
    // never enter a window that is read) and keeps the groups uniform.
    float sw[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) sw[i] = 0.f;
    static_assert((AA_L + 6) % 6 == 0, "AA_L + 6 must be a multiple of 6");
#pragma unroll 1
    for (int g = 0; g < (AA_L + 6) / 6; ++g) {
#pragma unroll
      for (int ii = 0; ii < 6; ++ii) {
        const int i = g * 6 + ii;
        float xw[7];
#pragma unroll
        for (int q = 0; q < 7; ++q) xw[q] = xs[r0 + i + q][c];
        float u = 0.f;                          // s[2a]: j = a-3 .. a+2, taps 11, 9, .., 1
#pragma unroll
        for (int q = 0; q < 6; ++q) u = fmaf(xw[q], flt.f[11 - 2 * q], u);
        sw[2 * ii] = snake_beta(2.0f * u, ea, ib);
        if (g > 0) {                            // chronological order: slots 2ii+1 .. 11, then 0 .. 2ii
          float y = 0.f;
#pragma unroll
          for (int k = 0; k < 12; ++k) y = fmaf(sw[(2 * ii + 1 + k) % 12], flt.f[k], y);
          out[base + (long)(t0 + i - 6) * ld] = __float2half_rn(y);
        }
        u = 0.f;                                // s[2a + 1]: j = a-2 .. a+3, taps 10, 8, .., 0
#pragma unroll
        for (int q = 0; q < 6; ++q) u = fmaf(xw[1 + q], flt.f[10 - 2 * q], u);
        sw[2 * ii + 1] = snake_beta(2.0f * u, ea, ib);
      }
    }
    return;
  }
  // sequence ends: the 2x-rate index is clamped too, straight from the definition
  const int t_end = min(T, t0 + AA_L);
  for (int t = t0; t < t_end; ++t) {
    float y = 0.f;
    for (int k = 0; k < 12; ++k) {
      int n = 2 * t + k - 5;
      n = n < 0 ? 0 : (n > 2 * T - 1 ? 2 * T - 1 : n);
      const int a = n >> 1;
      const int j0 = (n & 1) ? a - 2 : a - 3;  // first input row of this sample
      const int tap0 = (n & 1) ? 10 : 11;
      float u = 0.f;
      for (int q = 0; q < 6; ++q) u = fmaf(xs[j0 + q - (tile0 - 6)][c], flt.f[tap0 - 2 * q], u);
      y = fmaf(snake_beta(2.0f * u, ea, ib), flt.f[k], y);
    }
    out[base + (long)t * ld] = __float2half_rn(y);
  }
}

// (a + b + c) * scale -> fp16   (mean of the AMP blocks = the next stage's up-sampling operand)
__global__ void sum3_to_half_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                    float scale, __half* __restrict__ out, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i],
                 z = reinterpret_cast<const float4*>(c)[i];
    __half2 lo = __floats2half2_rn((x.x + y.x + z.x) * scale, (x.y + y.y + z.y) * scale);
    __half2 hi = __floats2half2_rn((x.z + y.z + z.z) * scale, (x.w + y.w + z.w) * scale);
    reinterpret_cast<uint2*>(out)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// conv_post: Conv1d(C -> 1, k 7, pad 3, optional bias) + clamp(-1, 1) or tanh.  a: fp16 [B, T, ld]; w: fp32 [7, ld].
__global__ void __launch_bounds__(256)
conv_post_kernel(const __half* __restrict__ a, const float* __restrict__ w, float bias, float* __restrict__ wav, int T,
                 int ld, int use_tanh) {
  extern __shared__ float wsm[];   // [7 * ld]
  for (int i = threadIdx.x; i < 7 * ld; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    float acc = bias;
    for (int k = 0; k < 7; ++k) {
      const int tt = t + k - 3;
      if (tt < 0 || tt >= T) continue;
      const uint4* row = reinterpret_cast<const uint4*>(a + ((long)b * T + tt) * ld);
      const float* wk = wsm + k * ld;
      for (int c8 = 0; c8 < ld / 8; ++c8) {
        const uint4 v = row[c8];
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(h[i]);
          acc = fmaf(f.x, wk[c8 * 8 + 2 * i], acc);
          acc = fmaf(f.y, wk[c8 * 8 + 2 * i + 1], acc);
        }
      }
    }
    wav[(long)b * T + t] = use_tanh ? tanhf(acc) : fminf(fmaxf(acc, -1.0f), 1.0f);
  }
}

__global__ void mel_rows128_kernel(const float* __restrict__ mel, __half* __restrict__ out, int batch, int ch, int t) {
  // [b, ch, t] fp32 -> [b, t, 128] fp16 (zero padded channels)
  const long total = (long)batch * t * 128;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int cidx = (int)(i & 127);
    const long bt = i >> 7;
    const int b = (int)(bt / t), tt = (int)(bt - (long)b * t);
    out[i] = __float2half_rn(cidx < ch ? mel[((long)b * ch + cidx) * t + tt] : 0.f);
  }
}

static int ew_grid(long n) {
  long g = (n + 255) / 256;
  const long cap = (long)sm_count() * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

static int pick_bn(int n) { return n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : 64); }

template <typename TIn>
static int snake_aa(const TIn* x0, const TIn* x1, const TIn* x2, float scale, const float* ab, __half* out, int batch, int T,
                    int ld, const AaFilter& flt, cudaStream_t st) {
  dim3 grid((T + AA_ROWS - 1) / AA_ROWS, ld / AA_CH, batch);
  if (x1 != nullptr)
    snake_aa_kernel<TIn, 3><<<grid, AA_THREADS, 0, st>>>(x0, x1, x2, scale, ab, out, T, ld, flt);
  else
    snake_aa_kernel<TIn, 1><<<grid, AA_THREADS, 0, st>>>(x0, x1, x2, scale, ab, out, T, ld, flt);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

// Conv1d(cin -> cout, k taps, dilation) over [B, T, cin] as a tap GEMM
static lemas_gemm_desc conv_desc(const __half* a, int batch, int T, int cin, const void* w, int cout, int taps, int dil,
                                 int epilogue) {
  lemas_gemm_desc d = {};
  d.a = a; d.batches = batch; d.rows = T; d.lda = cin; d.a_cols = cin;
  d.w = w; d.w_rows = taps * cout; d.ldw = cin; d.n = cout; d.k_per_tap = cin;
  d.taps = taps; d.tap_pad = (taps - 1) / 2; d.tap_dilation = dil; d.w_tap_stride = cout;
  d.block_n = pick_bn(cout); d.epilogue = epilogue; d.seq_len = T;
  return d;
}

struct BigvganBuffers {
  __half *mel16, *x16, *a16, *h16;
  float *xup, *r[3];
  int64_t bytes;
};

static int64_t max_stream_elems(const lemas_bigvgan_weights& w, int batch, int t, int64_t* max_in16) {
  int64_t T = t, best = 0, in16 = (int64_t)batch * t * w.ch0;
  for (int i = 0; i < w.stages; ++i) {
    T *= w.stage[i].rate;
    const int64_t e = (int64_t)batch * T * w.stage[i].ch_out;
    if (e > best) best = e;
    if (i + 1 < w.stages && e > in16) in16 = e;
  }
  *max_in16 = in16;
  return best;
}

static BigvganBuffers carve_bigvgan(const lemas_bigvgan_weights& w, int batch, int t, void* ws) {
  BigvganBuffers b;
  int64_t in16 = 0;
  const int64_t e = max_stream_elems(w, batch, t, &in16);
  // Carver from engine.cu is file-local there; same 1 KB-aligned bump allocation
  uint8_t* base = static_cast<uint8_t*>(ws);
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    off = align_up(off, 1024);
    uint8_t* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  };
  b.mel16 = reinterpret_cast<__half*>(take((int64_t)batch * t * 128 * 2));
  b.x16 = reinterpret_cast<__half*>(take(in16 * 2));
  b.a16 = reinterpret_cast<__half*>(take(e * 2));
  b.h16 = reinterpret_cast<__half*>(take(e * 2));
  b.xup = reinterpret_cast<float*>(take(e * 4));
  for (int j = 0; j < 3; ++j) b.r[j] = reinterpret_cast<float*>(take(e * 4));
  b.bytes = align_up(off, 1024);
  return b;
}

}  // namespace lemas

using namespace lemas;

extern "C" {

int64_t lemas_bigvgan_workspace_bytes(const lemas_bigvgan_weights* w, int32_t batch, int32_t t) {
  if (!w || !w->stage || w->stages < 1) return -1;
  return carve_bigvgan(*w, batch, t, nullptr).bytes;
}

int lemas_bigvgan_decode(const lemas_bigvgan_weights* w, const float* mel, float* wav, int32_t batch, int32_t t,
                         void* workspace, int64_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LEMAS_REQUIRE(w && mel && wav && workspace && w->stage, "lemas_bigvgan_decode: null argument");
  LEMAS_REQUIRE(w->num_mels >= 1 && w->num_mels <= 128 && w->ch0 % 64 == 0 && w->stages >= 1 && w->stages <= 8,
                "lemas_bigvgan_decode: unsupported dims");
  LEMAS_REQUIRE(batch >= 1 && t >= 1, "lemas_bigvgan_decode: bad shape");
  if (!lemas_device_supported())
    return fail(LEMAS_ERR_UNSUPPORTED,
                "CUDA error: no kernel image is available for execution on the device (liblemas_b200 is sm_100a only)");
  int cin = w->ch0;
  for (int i = 0; i < w->stages; ++i) {
    const lemas_bigvgan_stage& S = w->stage[i];
    LEMAS_REQUIRE(S.rate >= 1 && S.ch_in == cin && S.ch_out % 64 == 0 && S.ch_out >= 64 && S.up_w && S.up_b,
                  "lemas_bigvgan_decode: bad stage description");
    for (int j = 0; j < 3; ++j)
      LEMAS_REQUIRE(S.block[j].kernel >= 1 && (S.block[j].kernel & 1) == 1, "lemas_bigvgan_decode: odd resblock kernels only");
    cin = S.ch_out;
  }
  BigvganBuffers b = carve_bigvgan(*w, batch, t, workspace);
  LEMAS_REQUIRE(workspace_bytes >= b.bytes, "lemas_bigvgan_decode: workspace too small");
  AaFilter flt;
  for (int i = 0; i < 12; ++i) flt.f[i] = w->aa_filter[i];

  mel_rows128_kernel<<<ew_grid((long)batch * t * 128), 256, 0, st>>>(mel, b.mel16, batch, w->num_mels, t);
  LEMAS_LAUNCHED(1);
  {  // conv_pre: Conv1d(num_mels -> ch0, k 7, pad 3); the up-sampling GEMM that follows wants fp16 rows
    lemas_gemm_desc d = conv_desc(b.mel16, batch, t, 128, w->pre_w, w->ch0, 7, 1, LEMAS_EPI_BIAS_F16);
    d.bias = w->pre_b; d.out16 = b.x16; d.ld16 = w->ch0;
    LEMAS_TRY(gemm_launch(d, st));
  }
  int T = t;
  cin = w->ch0;
  for (int i = 0; i < w->stages; ++i) {
    const lemas_bigvgan_stage& S = w->stage[i];
    const int C = S.ch_out, r = S.rate;
    {  // ups[i]: ConvTranspose1d(cin -> C, k 2r, stride r, pad r/2) = 3-tap GEMM, N = r * C, rows [B, T] -> [B, T*r, C]
      lemas_gemm_desc d = conv_desc(b.x16, batch, T, cin, S.up_w, r * C, 3, 1, LEMAS_EPI_BIAS_F32);
      d.bias = S.up_b; d.out32 = b.xup; d.ld32 = r * C;
      LEMAS_TRY(gemm_launch(d, st));
    }
    T *= r;
    for (int j = 0; j < 3; ++j) {  // AMPBlock1 j on its own fp32 stream r[j]; the first residual add reads xup
      const lemas_bigvgan_block& K = S.block[j];
      for (int dd = 0; dd < 3; ++dd) {
        const float* src = dd == 0 ? b.xup : b.r[j];
        LEMAS_TRY(snake_aa<float>(src, nullptr, nullptr, 1.0f, K.act[2 * dd], b.a16, batch, T, C, flt, st));
        {
          lemas_gemm_desc d = conv_desc(b.a16, batch, T, C, K.w1[dd], C, K.kernel, K.dilation[dd], LEMAS_EPI_BIAS_F16);
          d.bias = K.b1[dd]; d.out16 = b.h16; d.ld16 = C;
          LEMAS_TRY(gemm_launch(d, st));
        }
        LEMAS_TRY(snake_aa<__half>(b.h16, nullptr, nullptr, 1.0f, K.act[2 * dd + 1], b.a16, batch, T, C, flt, st));
        {
          lemas_gemm_desc d = conv_desc(b.a16, batch, T, C, K.w2[dd], C, K.kernel, 1, LEMAS_EPI_GATE_RESID_F32);
          d.bias = K.b2[dd]; d.resid = src; d.ldr = C; d.out32 = b.r[j]; d.ld32 = C;
          LEMAS_TRY(gemm_launch(d, st));
        }
      }
    }
    if (i + 1 < w->stages) {  // x = mean of the three blocks -> the next up-sampling operand
      const long n4 = (long)batch * T * C / 4;
      sum3_to_half_kernel<<<ew_grid(n4), 256, 0, st>>>(b.r[0], b.r[1], b.r[2], 1.0f / 3.0f, b.x16, n4);
      LEMAS_LAUNCHED(1);
    }
    cin = C;
  }
  // activation_post on the mean of the last stage's blocks, conv_post, clamp
  LEMAS_TRY(snake_aa<float>(b.r[0], b.r[1], b.r[2], 1.0f / 3.0f, w->post_act, b.a16, batch, T, cin, flt, st));
  {
    int grid = (T + 255) / 256;
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    conv_post_kernel<<<dim3(grid, batch), 256, 7 * cin * sizeof(float), st>>>(b.a16, w->post_w, w->post_bias, wav, T, cin,
                                                                              w->use_tanh);
    LEMAS_LAUNCHED(1);
  }
  LEMAS_CUDA_OK(cudaGetLastError());
  return LEMAS_OK;
}
}
