// Prosody path of the prosody-conditioned model, once per utterance:
//
//   torchaudio.functional.resample(24 kHz -> 16 kHz)                                  cfm.py:254
//   extract_fbank_16k = torchaudio.compliance.kaldi.fbank(num_mel_bins=80)             prosody_encoder.py:337-361
//   ECAPA_TDNN.forward (TDNN / SE-Res2Net blocks, MFA, attentive statistics pooling,
//                       LayerNorm, fc, L2 normalisation)                               prosody_encoder.py:30-334
//
// Everything is fp32 on the CUDA cores: the whole encoder is ~7 GFLOP per 10 s utterance (0.01 % of the sampler) and
// its embedding conditions every frame, so exact fp32 arithmetic is worth more than tensor-core speed here.
// Activations are channels-last [batch, frames, channels]; every kernel takes a row stride and a channel offset so
// that the Res2Net channel groups and the MFA concatenation are views of one buffer, never copies.
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace lemas {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_TANH = 3 };

DEVI float act_f(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// Conv1d over frames, channels-last, "same" zero padding, dilation, groups, fused bias + activation.
//   y[b, t, yo + co] = act(bias[co] + sum_{k, ci} w[k][ci][co] * (x + add)[b, t + (k - (K-1)/2) * dil, xo + g*cin_g + ci])
// w is tap-major / input-major / output-contiguous: [K][cin_g][cout] (host-transposed nn.Conv1d weight).
// One CTA = 32 frames x 64 output channels; 256 threads, thread = (channel tx, 8 frames).  K-loop over taps x 32-channel
// chunks staged in shared memory.
// ---------------------------------------------------------------------------------------------------------------
struct ConvArgs {
  const float* x; int x_ld, x_off;
  const float* add; int add_ld, add_off;     // optional second input added to x before the convolution (Res2Net)
  const float* w; const float* bias;
  float* y; int y_ld, y_off;
  int batch, t, cin_g, cout, k, dil, groups, act;
};

__global__ void __launch_bounds__(256)
conv_cl_kernel(const ConvArgs a) {
  __shared__ float xs[32][33];
  __shared__ float ws[32][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // ty in [0,4): frames ty*8 .. ty*8+7
  const int t0 = blockIdx.x * 32;
  const int co0 = blockIdx.y * 64;
  const int b = blockIdx.z;
  const int cout_g = a.cout / a.groups;
  const int g = co0 / cout_g;                               // a 64-channel tile never straddles groups (host-checked)
  const int co = co0 + tx;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const float* xb = a.x + (long)b * a.t * a.x_ld + a.x_off + g * a.cin_g;
  const float* ab = a.add ? a.add + (long)b * a.t * a.add_ld + a.add_off + g * a.cin_g : nullptr;
  for (int k = 0; k < a.k; ++k) {
    const int shift = (k - (a.k - 1) / 2) * a.dil;
    for (int c0 = 0; c0 < a.cin_g; c0 += 32) {
      // stage x tile [32 frames][32 channels] (zero outside the sequence / channel range)
      for (int i = threadIdx.x; i < 32 * 32; i += 256) {
        const int f = i >> 5, c = i & 31;
        const int t = t0 + f + shift;
        float v = 0.f;
        if (t >= 0 && t < a.t && c0 + c < a.cin_g) {
          v = xb[(long)t * a.x_ld + c0 + c];
          if (ab) v += ab[(long)t * a.add_ld + c0 + c];
        }
        xs[f][c] = v;
      }
      // stage w tile [32 in][64 out]
      for (int i = threadIdx.x; i < 32 * 64; i += 256) {
        const int c = i >> 6, o = i & 63;
        float v = 0.f;
        if (c0 + c < a.cin_g && co0 + o < a.cout) v = a.w[((long)k * a.cin_g + c0 + c) * a.cout + co0 + o];
        ws[c][o] = v;
      }
      __syncthreads();
#pragma unroll 8
      for (int c = 0; c < 32; ++c) {
        const float wv = ws[c][tx];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(xs[ty * 8 + i][c], wv, acc[i]);
      }
      __syncthreads();
    }
  }
  if (co < a.cout) {
    const float bv = a.bias ? a.bias[co] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = t0 + ty * 8 + i;
      if (t < a.t) a.y[((long)b * a.t + t) * a.y_ld + a.y_off + co] = act_f(acc[i] + bv, a.act);
    }
  }
}

// LayerNorm over `c` channels of every row (view: ld / offset), affine, optional activation after.  Warp per row.
__global__ void __launch_bounds__(256)
ln_cl_kernel(const float* __restrict__ x, int x_ld, int x_off, float* __restrict__ y, int y_ld, int y_off,
             const float* __restrict__ gw, const float* __restrict__ gb, int rows, int c, float eps, int act) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long)row * x_ld + x_off;
  float s = 0.f;
  for (int i = lane; i < c; i += 32) s += xr[i];
  const float mean = warp_sum(s) / c;
  float v = 0.f;
  for (int i = lane; i < c; i += 32) { const float d = xr[i] - mean; v += d * d; }
  const float rstd = rsqrtf(warp_sum(v) / c + eps);
  float* yr = y + (long)row * y_ld + y_off;
  for (int i = lane; i < c; i += 32) yr[i] = act_f((xr[i] - mean) * rstd * gw[i] + gb[i], act);
}

// Per (batch, channel) statistics over frames: mean and, optionally, std = sqrt(clamp(mean((x - mean)^2), eps))
// (AttentiveStatisticsPooling._compute_statistics with uniform weights 1/T, prosody_encoder.py:247-252; SEBlock mean).
// CTA = 32 channels x 8 frame lanes.
__global__ void __launch_bounds__(256)
time_stats_kernel(const float* __restrict__ x, int x_ld, int x_off, int t_len, int c, float* __restrict__ mean_out,
                  float* __restrict__ std_out, int out_ld, float eps) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, tl = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cl;
  const int b = blockIdx.y;
  const float* xb = x + (long)b * t_len * x_ld + x_off + ch;
  float s = 0.f;
  if (ch < c) for (int t = tl; t < t_len; t += 8) s += xb[(long)t * x_ld];
  red[tl][cl] = s;
  __syncthreads();
  float mean = 0.f;
  for (int i = 0; i < 8; ++i) mean += red[i][cl];
  mean /= t_len;
  __syncthreads();
  if (std_out == nullptr) {
    if (tl == 0 && ch < c) mean_out[(long)b * out_ld + ch] = mean;
    return;
  }
  float v = 0.f;
  if (ch < c) for (int t = tl; t < t_len; t += 8) { const float d = xb[(long)t * x_ld] - mean; v += d * d; }
  red[tl][cl] = v;
  __syncthreads();
  if (tl == 0 && ch < c) {
    float var = 0.f;
    for (int i = 0; i < 8; ++i) var += red[i][cl];
    mean_out[(long)b * out_ld + ch] = mean;
    std_out[(long)b * out_ld + ch] = sqrtf(fmaxf(var / t_len, eps));
  }
}

// attn input of the pooling layer: [x | mean | std] per frame (prosody_encoder.py:262-267)
__global__ void asp_concat_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                  const float* __restrict__ sd, float* __restrict__ out, int batch, int t_len, int c) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)batch * t_len * 3 * c;
  if (i >= total) return;
  const int col = (int)(i % (3 * c));
  const long row = i / (3 * c);
  const int b = (int)(row / t_len);
  float v;
  if (col < c) v = x[row * c + col];
  else if (col < 2 * c) v = mean[(long)b * c + col - c];
  else v = sd[(long)b * c + col - 2 * c];
  out[i] = v;
}

// Attentive statistics: per (batch, channel) softmax of the logits over frames, weighted mean and std
// (prosody_encoder.py:271-278).  out: [batch, 2c] = [mean | std].  CTA = 32 channels x 8 frame lanes.
__global__ void __launch_bounds__(256)
asp_pool_kernel(const float* __restrict__ x, const float* __restrict__ logit, int t_len, int c, float* __restrict__ out,
                float eps) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, tl = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cl;
  const int b = blockIdx.y;
  const bool ok = ch < c;
  const float* xb = x + (long)b * t_len * c + ch;
  const float* lb = logit + (long)b * t_len * c + ch;
  auto reduce = [&](float v, bool is_max) {
    red[tl][cl] = v;
    __syncthreads();
    float r = red[0][cl];
    for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i][cl]) : r + red[i][cl];
    __syncthreads();
    return r;
  };
  float m = -INFINITY;
  if (ok) for (int t = tl; t < t_len; t += 8) m = fmaxf(m, lb[(long)t * c]);
  m = reduce(m, true);
  float se = 0.f, sx = 0.f;
  if (ok) for (int t = tl; t < t_len; t += 8) { const float e = expf(lb[(long)t * c] - m); se += e; sx += e * xb[(long)t * c]; }
  se = reduce(se, false);
  sx = reduce(sx, false);
  const float mean = sx / se;
  float sv = 0.f;
  if (ok) for (int t = tl; t < t_len; t += 8) {
    const float e = expf(lb[(long)t * c] - m);
    const float d = xb[(long)t * c] - mean;
    sv += e * d * d;
  }
  sv = reduce(sv, false);
  if (ok && tl == 0) {
    out[(long)b * 2 * c + ch] = mean;
    out[(long)b * 2 * c + c + ch] = sqrtf(fmaxf(sv / se, eps));
  }
}

// y[b, t, yo + c] = s[b, c] * x[b, t, c] + res[b, t, ro + c]        (SEBlock scale + SERes2NetBlock residual)
__global__ void se_scale_res_kernel(const float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ res,
                                    int res_ld, int res_off, float* __restrict__ y, int y_ld, int y_off, int batch,
                                    int t_len, int c) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)batch * t_len * c) return;
  const int col = (int)(i % c);
  const long row = i / c;
  const int b = (int)(row / t_len);
  y[row * y_ld + y_off + col] = s[(long)b * c + col] * x[i] + res[row * res_ld + res_off + col];
}

// copy a channel slice: y[row, yo + c] = x[row, xo + c]
__global__ void slice_copy_kernel(const float* __restrict__ x, int x_ld, int x_off, float* __restrict__ y, int y_ld,
                                  int y_off, long rows, int c) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= rows * c) return;
  const int col = (int)(i % c);
  const long row = i / c;
  y[row * y_ld + y_off + col] = x[row * x_ld + x_off + col];
}

// F.normalize(x, dim=-1): x / max(||x||, 1e-12); one warp per row
__global__ void l2norm_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int c) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int i = lane; i < c; i += 32) { const float v = x[(long)row * c + i]; s += v * v; }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  for (int i = lane; i < c; i += 32) y[(long)row * c + i] = x[(long)row * c + i] * inv;
}

// ---------------------------------------------------------------------------------------------------------------
// torchaudio.functional.resample (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) as a polyphase FIR:
//   out[q * up + j] = sum_k taps[j][k] * xpad[q * down + k],  xpad[p] = x[p - width] (zero outside), n_out = ceil(up n / down)
// ---------------------------------------------------------------------------------------------------------------
__global__ void resample_kernel(const float* __restrict__ x, int n, int x_ld, const float* __restrict__ taps, int n_taps,
                                int width, int up, int down, float* __restrict__ y, int n_out, int y_ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= n_out) return;
  const int j = i % up, q = i / up;
  const float* xb = x + (long)b * x_ld;
  const float* tp = taps + j * n_taps;
  float acc = 0.f;
  for (int k = 0; k < n_taps; ++k) {
    const int p = q * down + k - width;
    if (p >= 0 && p < n) acc = fmaf(tp[k], xb[p], acc);
  }
  y[(long)b * y_ld + i] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// torchaudio.compliance.kaldi.fbank(num_mel_bins, 16 kHz): frames of 400 samples every 160 (snip_edges), DC removal,
// pre-emphasis 0.97 (first sample against itself), povey window, zero pad to 512, |FFT|^2, mel filterbank (kaldi mel
// scale, 20 Hz .. Nyquist), log(max(., eps)).  One CTA per frame.
// ---------------------------------------------------------------------------------------------------------------
constexpr int FB_WIN = 400, FB_SHIFT = 160, FB_NFFT = 512;

__global__ void __launch_bounds__(256)
kaldi_fbank_kernel(const float* __restrict__ wav, int wav_ld, const float* __restrict__ window,
                   const float* __restrict__ banks, const int* __restrict__ bank_range, int n_mels, int n_frames,
                   float* __restrict__ out, float eps) {
  __shared__ float fr[FB_WIN];
  __shared__ float2 buf[FB_NFFT];
  __shared__ float2 tw[FB_NFFT / 2];
  __shared__ float pw[FB_NFFT / 2 + 1];
  __shared__ float red[8];
  const int frame = blockIdx.x % n_frames;
  const int b = blockIdx.x / n_frames;
  const float* w = wav + (long)b * wav_ld + (long)frame * FB_SHIFT;
  for (int k = threadIdx.x; k < FB_NFFT / 2; k += blockDim.x) {
    float s, c;
    sincospif((float)k * (-2.0f / FB_NFFT), &s, &c);
    tw[k] = make_float2(c, s);
  }
  float part = 0.f;
  for (int n = threadIdx.x; n < FB_WIN; n += blockDim.x) { fr[n] = w[n]; part += fr[n]; }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  float mean = 0.f;
  for (int i = 0; i < 8; ++i) mean += red[i];
  mean /= FB_WIN;
  for (int n = threadIdx.x; n < FB_NFFT; n += blockDim.x) {
    float v = 0.f;
    if (n < FB_WIN) {
      const float cur = fr[n] - mean;
      const float prev = fr[n > 0 ? n - 1 : 0] - mean;
      v = (cur - 0.97f * prev) * window[n];
    }
    buf[__brev((unsigned)n) >> 23] = make_float2(v, 0.f);
  }
  __syncthreads();
#pragma unroll 1
  for (int len = 2; len <= FB_NFFT; len <<= 1) {
    const int half = len >> 1;
    const int tstep = FB_NFFT / len;
    for (int i = threadIdx.x; i < FB_NFFT / 2; i += blockDim.x) {
      const int grp = i / half, j = i - grp * half;
      const int i0 = grp * len + j, i1 = i0 + half;
      const float2 t = tw[j * tstep];
      const float2 p = buf[i0], c = buf[i1];
      const float2 tc = make_float2(c.x * t.x - c.y * t.y, c.x * t.y + c.y * t.x);
      buf[i0] = make_float2(p.x + tc.x, p.y + tc.y);
      buf[i1] = make_float2(p.x - tc.x, p.y - tc.y);
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k <= FB_NFFT / 2; k += blockDim.x) pw[k] = buf[k].x * buf[k].x + buf[k].y * buf[k].y;
  __syncthreads();
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    const int lo = bank_range[2 * m], hi = bank_range[2 * m + 1];
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc = fmaf(pw[k], __ldg(banks + (long)m * (FB_NFFT / 2 + 1) + k), acc);
    out[((long)b * n_frames + frame) * n_mels + m] = logf(fmaxf(acc, eps));
  }
}

}  // namespace lemas

using namespace lemas;

namespace {

int conv(const ConvArgs& a, cudaStream_t st) {
  const int cout_g = a.cout / a.groups;
  if (a.groups > 1 && cout_g % 64 != 0) return fail(LEMAS_ERR_UNSUPPORTED, "prosody conv: grouped conv needs cout/groups % 64 == 0");
  dim3 grid((a.t + 31) / 32, (a.cout + 63) / 64, a.batch);
  conv_cl_kernel<<<grid, 256, 0, st>>>(a);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int ln(const float* x, int x_ld, int x_off, float* y, int y_ld, int y_off, const float* w, const float* b, int rows, int c,
       int act, cudaStream_t st) {
  ln_cl_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, x_ld, x_off, y, y_ld, y_off, w, b, rows, c, 1e-12f, act);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

// TDNNBlock: conv -> ReLU -> LayerNorm (prosody_encoder.py:135-158); tmp holds the pre-norm activations
int tdnn(const lemas_prosody_tdnn& p, const float* x, int x_ld, int x_off, const float* add, int add_ld, int add_off,
         float* tmp, float* y, int y_ld, int y_off, int batch, int t, int act_after, cudaStream_t st) {
  ConvArgs a;
  a.x = x; a.x_ld = x_ld; a.x_off = x_off;
  a.add = add; a.add_ld = add_ld; a.add_off = add_off;
  a.w = p.w; a.bias = p.b;
  a.y = tmp; a.y_ld = p.cout; a.y_off = 0;
  a.batch = batch; a.t = t; a.cin_g = p.cin / p.groups; a.cout = p.cout; a.k = p.k; a.dil = p.dil; a.groups = p.groups;
  a.act = ACT_RELU;
  LEMAS_TRY(conv(a, st));
  return ln(tmp, p.cout, 0, y, y_ld, y_off, p.ln_w, p.ln_b, batch * t, p.cout, act_after, st);
}

int dense(const float* x, int cin, const float* w, const float* b, float* y, int cout, int rows, int act, cudaStream_t st) {
  ConvArgs a;   // a k=1 convolution over `rows` frames of one batch item
  a.x = x; a.x_ld = cin; a.x_off = 0; a.add = nullptr; a.add_ld = 0; a.add_off = 0;
  a.w = w; a.bias = b; a.y = y; a.y_ld = cout; a.y_off = 0;
  a.batch = 1; a.t = rows; a.cin_g = cin; a.cout = cout; a.k = 1; a.dil = 1; a.groups = 1; a.act = act;
  return conv(a, st);
}

}  // namespace

extern "C" {

int64_t lemas_prosody_workspace_bytes(const lemas_prosody_weights* w, int32_t batch, int32_t t) {
  if (!w || batch < 1 || t < 1) return -1;
  const int64_t rows = (int64_t)batch * t;
  const int64_t c = w->channels, cm = w->mfa_channels;
  // x0, t1, y, t2, tmp (c each) | cat, mfa, logits, tmp (cm each) | concat (3 cm) | attn (att) x2 | small vectors
  int64_t floats = rows * (5 * c + 4 * cm + 3 * cm + 2 * (int64_t)w->att_channels) + (int64_t)batch * (8 * cm + 4 * c) + 1024;
  return floats * 4 + 4096;
}

int lemas_prosody_encode(const lemas_prosody_weights* w, const float* fbank, int32_t batch, int32_t t, float* out,
                         void* workspace, int64_t workspace_bytes, void* stream) {
  LEMAS_REQUIRE(w && fbank && out && workspace, "lemas_prosody_encode: null pointer");
  LEMAS_REQUIRE(batch >= 1 && t >= 1, "lemas_prosody_encode: bad shape");
  LEMAS_REQUIRE(workspace_bytes >= lemas_prosody_workspace_bytes(w, batch, t), "lemas_prosody_encode: workspace too small");
  LEMAS_REQUIRE(w->n_blocks >= 1 && w->n_blocks <= 8 && w->scale >= 2 && w->channels % w->scale == 0 &&
                w->mfa_channels == w->channels * w->n_blocks, "lemas_prosody_encode: unsupported architecture");
  cudaStream_t st = (cudaStream_t)stream;
  const int c = w->channels, cm = w->mfa_channels, att = w->att_channels, sc = w->scale, cg = c / sc;
  const long rows = (long)batch * t;
  float* p = static_cast<float*>(workspace);
  auto take = [&](long n) { float* r = p; p += (n + 63) / 64 * 64; return r; };
  float* x0 = take(rows * c);      // block 0 output
  float* t1 = take(rows * c);
  float* y = take(rows * c);
  float* t2 = take(rows * c);
  float* tmp = take(rows * cm);    // pre-norm scratch (largest conv output is mfa_channels wide)
  float* cat = take(rows * cm);    // outputs of the SE-Res2Net blocks side by side == torch.cat(xl[1:], dim=1)
  float* mfa = take(rows * cm);
  float* logits = take(rows * cm);
  float* concat = take(rows * 3 * cm);
  float* a1 = take(rows * att);
  float* gmean = take((long)batch * cm);
  float* gstd = take((long)batch * cm);
  float* pooled = take((long)batch * 2 * cm);
  float* pooled_n = take((long)batch * 2 * cm);
  float* se_m = take((long)batch * c);
  float* se_h = take((long)batch * w->se_channels);
  float* se_s = take((long)batch * c);
  float* emb = take((long)batch * w->embed_dim);

  // blocks[0]: TDNN(input_dim -> c)
  LEMAS_TRY(tdnn(w->block0, fbank, w->input_dim, 0, nullptr, 0, 0, tmp, x0, c, 0, batch, t, ACT_NONE, st));
  const float* xin = x0; int xin_ld = c, xin_off = 0;
  for (int bi = 0; bi < w->n_blocks; ++bi) {
    const lemas_prosody_block& blk = w->blocks[bi];
    // tdnn1 (1x1)
    LEMAS_TRY(tdnn(blk.tdnn1, xin, xin_ld, xin_off, nullptr, 0, 0, tmp, t1, c, 0, batch, t, ACT_NONE, st));
    // Res2Net: group 0 passes through, group i runs its own TDNN on x_i (+ y_{i-1})      (prosody_encoder.py:187-199)
    slice_copy_kernel<<<(int)((rows * cg + 255) / 256), 256, 0, st>>>(t1, c, 0, y, c, 0, rows, cg);
    LEMAS_LAUNCHED(1);
    for (int i = 1; i < sc; ++i)
      LEMAS_TRY(tdnn(blk.res2[i - 1], t1, c, i * cg, i >= 2 ? y : nullptr, c, (i - 1) * cg, tmp, y, c, i * cg, batch, t,
                     ACT_NONE, st));
    // tdnn2 (1x1)
    LEMAS_TRY(tdnn(blk.tdnn2, y, c, 0, nullptr, 0, 0, tmp, t2, c, 0, batch, t, ACT_NONE, st));
    // SE: s = sigmoid(W2 relu(W1 mean_t(x)))                                              (prosody_encoder.py:215-226)
    time_stats_kernel<<<dim3((c + 31) / 32, batch), 256, 0, st>>>(t2, c, 0, t, c, se_m, nullptr, c, 0.f);
    LEMAS_LAUNCHED(1);
    LEMAS_TRY(dense(se_m, c, blk.se_w1, blk.se_b1, se_h, w->se_channels, batch, ACT_RELU, st));
    LEMAS_TRY(dense(se_h, w->se_channels, blk.se_w2, blk.se_b2, se_s, c, batch, ACT_SIGMOID, st));
    // out = s * x + residual -> column slice bi of the concatenation buffer              (prosody_encoder.py:326-334)
    se_scale_res_kernel<<<(int)((rows * c + 255) / 256), 256, 0, st>>>(t2, se_s, xin, xin_ld, xin_off, cat, cm, bi * c,
                                                                       batch, t, c);
    LEMAS_LAUNCHED(1);
    xin = cat; xin_ld = cm; xin_off = bi * c;
  }
  // MFA: TDNN(cm -> cm, k=1, groups)
  LEMAS_TRY(tdnn(w->mfa, cat, cm, 0, nullptr, 0, 0, tmp, mfa, cm, 0, batch, t, ACT_NONE, st));
  // attentive statistics pooling with global context                                       (prosody_encoder.py:239-279)
  time_stats_kernel<<<dim3((cm + 31) / 32, batch), 256, 0, st>>>(mfa, cm, 0, t, cm, gmean, gstd, cm, 1e-12f);
  LEMAS_LAUNCHED(1);
  asp_concat_kernel<<<(int)((rows * 3 * cm + 255) / 256), 256, 0, st>>>(mfa, gmean, gstd, concat, batch, t, cm);
  LEMAS_LAUNCHED(1);
  LEMAS_TRY(tdnn(w->asp_tdnn, concat, 3 * cm, 0, nullptr, 0, 0, tmp, a1, att, 0, batch, t, ACT_TANH, st));
  {
    ConvArgs a;
    a.x = a1; a.x_ld = att; a.x_off = 0; a.add = nullptr; a.add_ld = 0; a.add_off = 0;
    a.w = w->asp_conv_w; a.bias = w->asp_conv_b; a.y = logits; a.y_ld = cm; a.y_off = 0;
    a.batch = batch; a.t = t; a.cin_g = att; a.cout = cm; a.k = 1; a.dil = 1; a.groups = 1; a.act = ACT_NONE;
    LEMAS_TRY(conv(a, st));
  }
  asp_pool_kernel<<<dim3((cm + 31) / 32, batch), 256, 0, st>>>(mfa, logits, t, cm, pooled, 1e-12f);
  LEMAS_LAUNCHED(1);
  LEMAS_TRY(ln(pooled, 2 * cm, 0, pooled_n, 2 * cm, 0, w->asp_norm_w, w->asp_norm_b, batch, 2 * cm, ACT_NONE, st));
  LEMAS_TRY(dense(pooled_n, 2 * cm, w->fc_w, w->fc_b, emb, w->embed_dim, batch, ACT_NONE, st));
  l2norm_kernel<<<(batch + 7) / 8, 256, 0, st>>>(emb, out, batch, w->embed_dim);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int lemas_resample_sinc(const float* x, int32_t batch, int32_t n, int32_t x_ld, const float* taps, int32_t n_taps,
                        int32_t width, int32_t up, int32_t down, float* y, int32_t n_out, int32_t y_ld, void* stream) {
  LEMAS_REQUIRE(x && taps && y, "lemas_resample_sinc: null pointer");
  LEMAS_REQUIRE(batch >= 1 && n >= 1 && up >= 1 && down >= 1 && n_taps >= 1 && n_out >= 1, "lemas_resample_sinc: bad shape");
  resample_kernel<<<dim3((n_out + 255) / 256, batch), 256, 0, (cudaStream_t)stream>>>(x, n, x_ld, taps, n_taps, width, up,
                                                                                      down, y, n_out, y_ld);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int lemas_kaldi_fbank_16k(const float* wav, int32_t batch, int32_t n, int32_t wav_ld, const float* window,
                          const float* banks, const int32_t* bank_range, int32_t n_mels, float* out, void* stream) {
  LEMAS_REQUIRE(wav && window && banks && bank_range && out, "lemas_kaldi_fbank_16k: null pointer");
  LEMAS_REQUIRE(batch >= 1 && n >= FB_WIN && n_mels >= 1, "lemas_kaldi_fbank_16k: needs at least 400 samples");
  const int n_frames = 1 + (n - FB_WIN) / FB_SHIFT;
  kaldi_fbank_kernel<<<batch * n_frames, 256, 0, (cudaStream_t)stream>>>(wav, wav_ld, window, banks, bank_range, n_mels,
                                                                         n_frames, out, 1.1920928955078125e-07f);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}
}
