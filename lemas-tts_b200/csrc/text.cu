// Text embedding of the DiT (dit.py:51-81) on the device: token embedding + absolute position table + filler mask,
// then `layers` ConvNeXt-V2 blocks (modules.py:241-269): depthwise conv k7 -> LayerNorm -> Linear(C -> 2C) -> GELU(erf)
// -> GRN over the SEQUENCE dimension (modules.py:225-234) -> Linear(2C -> C) -> + residual -> filler rows zeroed.
// Runs once per CFM.sample for the conditional and the unconditional copy of the text (dit.py:212-220), not per ODE
// step.  The two Linears run on the tcgen05 GEMMs (GELU / residual fused in their epilogues), dwconv+LN reuses the
// Vocos kernel; this file adds the gather, the GRN reduction / apply and the row mask, plus the sequencing.
#include "common.h"
#include "ptx.cuh"

namespace lemas {

int gemm_launch(const lemas_gemm_desc& d, cudaStream_t stream);

// out[r, :] = filler(r) ? 0 : table[drop(b) ? 0 : id] + pos[n]     (dit.py:52-71); mask[r] = filler(r)
__global__ void text_init_kernel(const int* __restrict__ ids, const uint8_t* __restrict__ drop,
                                 const float* __restrict__ table, const float* __restrict__ pos, float* __restrict__ out,
                                 uint8_t* __restrict__ mask, int rows, int seq, int dim, int mask_padding) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int b = r / seq, n = r - b * seq;
  const int id = ids[r];
  const bool filler = id == 0;  // taken before the ids are dropped (dit.py:56-60)
  if (threadIdx.x == 0) mask[r] = (filler && mask_padding) ? 1 : 0;
  const float* trow = table + (long)(drop[b] ? 0 : id) * dim;
  const float* prow = pos ? pos + (long)min(n, 4095) * dim : nullptr;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float v = trow[c] + (prow ? prow[c] : 0.f);
    if (filler && mask_padding && pos) v = 0.f;  // the reference masks only on the extra-modeling path
    out[(long)r * dim + c] = v;
  }
}

// GRN statistics, stage 1: part[b, chunk, c] = sum over a chunk of rows of h[b, n, c]^2.  No atomics: the order of
// summation is fixed, so the embedding (and everything downstream) is bit-reproducible from call to call.
__global__ void grn_stats_kernel(const __half* __restrict__ h, float* __restrict__ part, int seq, int ch,
                                 int rows_per_block) {
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * rows_per_block;
  const int n1 = min(seq, n0 + rows_per_block);
  for (int c8 = threadIdx.x; c8 < ch / 8; c8 += blockDim.x) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int n = n0; n < n1; ++n) {
      const uint4 u = *reinterpret_cast<const uint4*>(h + ((long)b * seq + n) * ch + c8 * 8);
      const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(hp[i]);
        acc[2 * i] += f.x * f.x;
        acc[2 * i + 1] += f.y * f.y;
      }
    }
    float* dst = part + ((long)b * gridDim.x + blockIdx.x) * ch + c8 * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = acc[i];
  }
}

// stage 2: sumsq[b, c] = sum over chunks (ascending)
__global__ void grn_reduce_kernel(const float* __restrict__ part, float* __restrict__ sumsq, int chunks, int ch) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ch) return;
  float t = 0.f;
  for (int k = 0; k < chunks; ++k) t += part[((long)b * chunks + k) * ch + c];
  sumsq[(long)b * ch + c] = t;
}

// GRN apply: Gx = sqrt(sumsq), Nx = Gx / (mean_c Gx + 1e-6), out = gamma * (h * Nx) + beta + h   -> fp16
__global__ void __launch_bounds__(256)
grn_apply_kernel(const __half* __restrict__ h, const float* __restrict__ sumsq, const float* __restrict__ gamma,
                 const float* __restrict__ beta, __half* __restrict__ out, int seq, int ch, int rows_per_block) {
  __shared__ float red[8];
  __shared__ float s_mean;
  const int b = blockIdx.y;
  float part = 0.f;
  for (int c = threadIdx.x; c < ch; c += blockDim.x) part += sqrtf(sumsq[(long)b * ch + c]);
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    s_mean = t / ch;
  }
  __syncthreads();
  const float inv = 1.0f / (s_mean + 1e-6f);
  const int n0 = blockIdx.x * rows_per_block;
  const int n1 = min(seq, n0 + rows_per_block);
  for (int c2 = threadIdx.x; c2 < ch / 2; c2 += blockDim.x) {
    const int c = 2 * c2;
    const float nx0 = sqrtf(sumsq[(long)b * ch + c]) * inv, nx1 = sqrtf(sumsq[(long)b * ch + c + 1]) * inv;
    const float g0 = gamma[c], g1 = gamma[c + 1], b0 = beta[c], b1 = beta[c + 1];
    for (int n = n0; n < n1; ++n) {
      const long o = ((long)b * seq + n) * ch + c;
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(h + o));
      *reinterpret_cast<__half2*>(out + o) = __floats2half2_rn(g0 * (f.x * nx0) + b0 + f.x, g1 * (f.y * nx1) + b1 + f.y);
    }
  }
}

__global__ void zero_masked_rows_kernel(float* __restrict__ x, const uint8_t* __restrict__ mask, int rows, int dim) {
  const int r = blockIdx.x;
  if (r >= rows || !mask[r]) return;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) x[(long)r * dim + c] = 0.f;
}

constexpr int GRN_ROWS = 64;  // sequence rows per statistics block

struct TextBuffers {
  __half *a16, *h16, *g16;
  float *sumsq, *part;
  uint8_t* mask;
  int64_t bytes;
};

static TextBuffers carve_text(const lemas_text_weights& w, int batch, int seq, void* ws) {
  TextBuffers b;
  uint8_t* base = static_cast<uint8_t*>(ws);
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    off = align_up(off, 1024);
    uint8_t* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  };
  const int64_t R = (int64_t)batch * seq;
  b.a16 = reinterpret_cast<__half*>(take(R * w.dim * 2));
  b.h16 = reinterpret_cast<__half*>(take(R * w.inter * 2));
  b.g16 = reinterpret_cast<__half*>(take(R * w.inter * 2));
  b.sumsq = reinterpret_cast<float*>(take((int64_t)batch * w.inter * 4));
  b.part = reinterpret_cast<float*>(take((int64_t)batch * ((seq + GRN_ROWS - 1) / GRN_ROWS) * w.inter * 4));
  b.mask = take(R);
  b.bytes = align_up(off, 1024);
  return b;
}

}  // namespace lemas

using namespace lemas;

extern "C" {

int64_t lemas_text_workspace_bytes(const lemas_text_weights* w, int32_t batch, int32_t seq) {
  if (!w) return -1;
  return carve_text(*w, batch, seq, nullptr).bytes;
}

int lemas_text_embedding(const lemas_text_weights* w, const int32_t* ids, const uint8_t* drop, float* out, int32_t batch,
                         int32_t seq, void* workspace, int64_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LEMAS_REQUIRE(w && ids && drop && out && workspace, "lemas_text_embedding: null argument");
  LEMAS_REQUIRE(batch >= 1 && seq >= 1, "lemas_text_embedding: bad shape");
  LEMAS_REQUIRE(w->layers == 0 || (w->dim % 128 == 0 && w->dim <= 1024 && w->inter % 64 == 0 && w->inter % 8 == 0),
                "lemas_text_embedding: text_dim must be a multiple of 128 (<= 1024)");
  if (!lemas_device_supported())
    return fail(LEMAS_ERR_UNSUPPORTED,
                "CUDA error: no kernel image is available for execution on the device (liblemas_b200 is sm_100a only)");
  TextBuffers b = carve_text(*w, batch, seq, workspace);
  LEMAS_REQUIRE(workspace_bytes >= b.bytes, "lemas_text_embedding: workspace too small");
  const int R = batch * seq;
  const int dim = w->dim, inter = w->inter;
  text_init_kernel<<<R, 128, 0, st>>>(ids, drop, w->table, w->layers > 0 ? w->pos : nullptr, out, b.mask, R, seq, dim,
                                      w->mask_padding);
  LEMAS_LAUNCHED(1);
  const int rpb = GRN_ROWS;
  const dim3 grid_rows((seq + rpb - 1) / rpb, batch);
  for (int l = 0; l < w->layers; ++l) {
    const lemas_text_block& L = w->blocks[l];
    LEMAS_TRY(lemas_dwconv7_ln(out, L.dw_w, L.dw_b, L.ln_w, L.ln_b, b.a16, batch, seq, dim, st));
    {
      lemas_gemm_desc d = {};
      d.a = b.a16; d.batches = 1; d.rows = R; d.lda = dim; d.a_cols = dim;
      d.w = L.w1; d.w_rows = inter; d.ldw = dim; d.n = inter; d.k_per_tap = dim; d.taps = 1;
      d.block_n = inter % 256 == 0 ? 256 : 128; d.epilogue = LEMAS_EPI_GELU_ERF_F16; d.bias = L.b1;
      d.out16 = b.h16; d.ld16 = inter; d.seq_len = seq;
      LEMAS_TRY(gemm_launch(d, st));
    }
    grn_stats_kernel<<<grid_rows, 128, 0, st>>>(b.h16, b.part, seq, inter, rpb);
    grn_reduce_kernel<<<dim3((inter + 255) / 256, batch), 256, 0, st>>>(b.part, b.sumsq, (int)grid_rows.x, inter);
    grn_apply_kernel<<<grid_rows, 256, 0, st>>>(b.h16, b.sumsq, L.grn_gamma, L.grn_beta, b.g16, seq, inter, rpb);
    LEMAS_LAUNCHED(3);
    {
      lemas_gemm_desc d = {};
      d.a = b.g16; d.batches = 1; d.rows = R; d.lda = inter; d.a_cols = inter;
      d.w = L.w2; d.w_rows = dim; d.ldw = inter; d.n = dim; d.k_per_tap = inter; d.taps = 1;
      d.block_n = dim % 256 == 0 ? 256 : 128; d.epilogue = LEMAS_EPI_GATE_RESID_F32; d.bias = L.b2;
      d.resid = out; d.ldr = dim; d.out32 = out; d.ld32 = dim; d.seq_len = seq;
      LEMAS_TRY(gemm_launch(d, st));
    }
    if (w->mask_padding) {
      zero_masked_rows_kernel<<<R, 128, 0, st>>>(out, b.mask, R, dim);
      LEMAS_LAUNCHED(1);
    }
  }
  return LEMAS_OK;
}
}
