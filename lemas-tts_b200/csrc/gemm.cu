// Persistent warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   acc[m, n] = sum_k A[m, k] * W[n, k]            fp16 operands, fp32 accumulation in TMEM
//
// One CTA per SM loops over 128 x BLOCK_N output tiles.  Roles (192 threads):
//   warp 0    TMA producer   : cp.async.bulk.tensor (128B swizzle) into a STAGES-deep smem ring, mbarrier tx-count
//   warp 1    MMA issuer     : one elected lane issues tcgen05.mma (M=128, N=BLOCK_N, K=16) x4 per 64-wide k block,
//                              tcgen05.commit releases smem slots and publishes the accumulator
//   warps 2-5 epilogue       : tcgen05.ld (32 lanes x 32 columns) -> registers -> fused epilogue -> global
// The accumulator is double buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.
//
// "taps" turns the GEMM into an im2col-free 1-D convolution along rows: each tap is one more k block whose A tile
// is the same activation matrix shifted by (tap - pad) rows; rows outside the sequence are zero-filled by TMA's
// out-of-bounds handling, which is exactly Conv1d zero padding.  Grouped convs select the A column block by n tile.
//
// Replaces the cuBLASLt calls behind nn.Linear / nn.Conv1d on the reference hot path
// (modules.py:171-176, 349-350, 452-454, 495; dit.py:97, 252) — see include/lemas_b200.h for the epilogues.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"
#include "gemm_params.cuh"

namespace lemas {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

DEVI void store16x32(__half* dst, const float (&v)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
    d[i] = u;
  }
}
DEVI void store32x32(float* dst, const float (&v)[32]) {
  float4* d = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
DEVI void add32(float (&v)[32], const float* src) {
  const float4* s = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 t = __ldg(s + i);
    v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
  }
}

template <int EPI>
DEVI void epilogue_chunk(const GemmParams& p, float (&v)[32], long grow, int col0) {
  const int nvalid = p.n - col0;  // > 0 guaranteed by caller
  const bool full = nvalid >= 32;
  if (EPI != LEMAS_EPI_ADD_F32_F16 && p.bias != nullptr) {
    if (full) {
      add32(v, p.bias + col0);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) v[i] += __ldg(p.bias + col0 + i);
    }
  }
  if constexpr (EPI == LEMAS_EPI_BIAS_F16) {
    store16x32(p.out16 + grow * p.ld16 + col0, v);
  } else if constexpr (EPI == LEMAS_EPI_QKV_ROPE) {
    const int b = (int)(grow / p.seq_len);
    const int pos = (int)(grow - (long)b * p.seq_len);
    if (col0 < 2 * p.inner) {
      const int within = col0 % p.inner;
      if (within < p.rope_cols) {
        const float4* cs = reinterpret_cast<const float4*>(p.rope + (long)pos * 32 + (within & 63) / 2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // two (cos, sin) pairs per float4
          float4 t = __ldg(cs + j);
          float x0 = v[4 * j], x1 = v[4 * j + 1], x2 = v[4 * j + 2], x3 = v[4 * j + 3];
          v[4 * j] = x0 * t.x - x1 * t.y;
          v[4 * j + 1] = x1 * t.x + x0 * t.y;
          v[4 * j + 2] = x2 * t.z - x3 * t.w;
          v[4 * j + 3] = x3 * t.z + x2 * t.w;
        }
      }
      store16x32(p.out16 + grow * p.ld16 + col0, v);
    } else {
      const int vcol = col0 - 2 * p.inner;
      const int heads = p.inner >> 6;
      __half* dst = p.vt + ((long)(b * heads + (vcol >> 6)) * 64 + (vcol & 63)) * p.vt_ld + pos;
#pragma unroll
      for (int i = 0; i < 32; ++i) dst[(long)i * p.vt_ld] = __float2half_rn(v[i]);  // lanes = consecutive pos
    }
  } else if constexpr (EPI == LEMAS_EPI_GELU_TANH_F16) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_tanh_f(v[i]);
    store16x32(p.out16 + grow * p.ld16 + col0, v);
  } else if constexpr (EPI == LEMAS_EPI_GELU_ERF_F16) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf_f(v[i]);
    store16x32(p.out16 + grow * p.ld16 + col0, v);
  } else if constexpr (EPI == LEMAS_EPI_GATE_RESID_F32) {
    const int b = (int)(grow / p.seq_len);
    const int pos = (int)(grow - (long)b * p.seq_len);
    const bool dead = p.row_valid != nullptr && pos >= __ldg(p.row_valid + b);
    const float4* r = reinterpret_cast<const float4*>(p.resid + grow * p.ldr + col0);
    float4* o = reinterpret_cast<float4*>(p.out32 + grow * p.ld32 + col0);
    const float4* g = p.gate ? reinterpret_cast<const float4*>(p.gate + (long)b * p.gate_bstride + col0) : nullptr;
    if (p.red_add == 2) {   // in place: x += gate (acc + bias) as 16-byte reductions in the L2 (see gemm2.cu)
      if (!dead) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 gg = g ? __ldg(g + i) : make_float4(1.f, 1.f, 1.f, 1.f);
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i), "f"(gg.x * v[4 * i]),
                       "f"(gg.y * v[4 * i + 1]), "f"(gg.z * v[4 * i + 2]), "f"(gg.w * v[4 * i + 3]) : "memory");
        }
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 rr = r[i];
      float4 gg = g ? __ldg(g + i) : make_float4(1.f, 1.f, 1.f, 1.f);
      float a0 = dead ? 0.f : v[4 * i], a1 = dead ? 0.f : v[4 * i + 1];
      float a2 = dead ? 0.f : v[4 * i + 2], a3 = dead ? 0.f : v[4 * i + 3];
      o[i] = make_float4(rr.x + gg.x * a0, rr.y + gg.y * a1, rr.z + gg.z * a2, rr.w + gg.w * a3);
    }
  } else if constexpr (EPI == LEMAS_EPI_BIAS_F32) {
    float* o = p.out32 + grow * p.ld32 + col0;
    if (full) {
      store32x32(o, v);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) o[i] = v[i];
    }
  } else if constexpr (EPI == LEMAS_EPI_ADD_F32_F16) {
    add32(v, p.resid + grow * p.ldr + col0);
    store32x32(p.out32 + grow * p.ld32 + col0, v);
    if (p.out16) store16x32(p.out16 + grow * p.ld16 + col0, v);
  } else if constexpr (EPI == LEMAS_EPI_MISH_F16) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = mish_f(v[i]);
    store16x32(p.out16 + grow * p.ld16 + col0, v);
  } else if constexpr (EPI == LEMAS_EPI_MISH_RESID_F32) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = mish_f(v[i]);
    add32(v, p.resid + grow * p.ldr + col0);
    store32x32(p.out32 + grow * p.ld32 + col0, v);
  }
}

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int STAGES = BLOCK_N == 256 ? 4 : 6;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + slack for 1024 B alignment
};

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ GemmParams p) {
  using S = GemmSmem<BLOCK_N>;
  constexpr int STAGES = S::STAGES;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles_per_batch = (p.rows + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (p.n + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = p.batches * m_tiles_per_batch * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full + a, 1);
      mbar_init(acc_empty + a, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_idx = tile % n_tiles;
        const int mt = tile / n_tiles;
        const int b = mt / m_tiles_per_batch;
        const int row0 = (mt - b * m_tiles_per_batch) * BLOCK_M;
        const int n0 = n_idx * BLOCK_N;
        const int a_col0 = n_idx * p.group_cols;
        for (int it = 0; it < p.k_iters; ++it) {
          const int tap = it / p.kc_per_tap;
          const int kc = it - tap * p.kc_per_tap;
          mbar_wait(empty_bar + stage, phase ^ 1);
          mbar_arrive_expect_tx(full_bar + stage, S::STAGE_BYTES);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          tma_load_3d(sa, &tmA, full_bar + stage, a_col0 + kc * BLOCK_K, row0 + (tap - p.tap_pad) * p.tap_dil, b);
          tma_load_2d(sa + S::A_BYTES, &tmW, full_bar + stage, kc * BLOCK_K, tap * p.w_tap_stride + n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16(BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(acc_empty + acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int it = 0; it < p.k_iters; ++it) {
        mbar_wait(full_bar + stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + S::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row: start-address field is in 16 B units
            umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar + stage);
          if (it == p.k_iters - 1) umma_commit(acc_full + acc);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    const int sub = warp & 3;  // TMEM sub-partition this warp may read: lanes [32*sub, 32*sub+32)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_idx = tile % n_tiles;
      const int mt = tile / n_tiles;
      const int b = mt / m_tiles_per_batch;
      const int row_in_batch = (mt - b * m_tiles_per_batch) * BLOCK_M + sub * 32 + lane;
      const bool row_ok = row_in_batch < p.rows;
      const long grow = (long)b * p.rows + row_in_batch;
      const int n0 = n_idx * BLOCK_N;
      mbar_wait(acc_full + acc, acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(sub * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        if (n0 + c >= p.n) break;
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c, r);
        tmem_ld_wait();
        if (row_ok) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          epilogue_chunk<EPI>(p, v, grow, n0 + c);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int BLOCK_N, int EPI>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmParams& p, int max_ctas,
                  cudaStream_t stream) {
  using S = GemmSmem<BLOCK_N>;
  static unsigned long long configured = 0;
  auto kern = gemm_kernel<BLOCK_N, EPI>;
  LEMAS_CUDA_OK(ensure_dynamic_smem(kern, S::TOTAL, configured));
  const int m_tiles = p.batches * ((p.rows + BLOCK_M - 1) / BLOCK_M);
  const int n_tiles = (p.n + BLOCK_N - 1) / BLOCK_N;
  int grid = m_tiles * n_tiles;
  const int cap = max_ctas > 0 ? max_ctas : sm_count();
  if (grid > cap) grid = cap;
  kern<<<grid, GEMM_THREADS, S::TOTAL, stream>>>(tmA, tmW, p);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

template <int EPI>
static int dispatch_bn(int bn, const CUtensorMap& a, const CUtensorMap& w, const GemmParams& p, int max_ctas,
                       cudaStream_t s) {
  switch (bn) {
    case 64: return launch<64, EPI>(a, w, p, max_ctas, s);
    case 128: return launch<128, EPI>(a, w, p, max_ctas, s);
    case 256: return launch<256, EPI>(a, w, p, max_ctas, s);
  }
  return fail(LEMAS_ERR_INVALID, "lemas_gemm_f16: block_n must be 64, 128 or 256");
}

bool gemm2_eligible(const lemas_gemm_desc& d);
int gemm2_launch(const lemas_gemm_desc& d, const GemmParams& p, cudaStream_t st);

// LEMAS_GEMM_PAIR=0 forces the single-CTA kernel everywhere (A/B measurements)
static bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LEMAS_GEMM_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int gemm_launch(const lemas_gemm_desc& d, cudaStream_t stream) {
  LEMAS_REQUIRE(d.a && d.w, "lemas_gemm_f16: null operand");
  LEMAS_REQUIRE(d.k_per_tap > 0 && d.k_per_tap % BLOCK_K == 0, "lemas_gemm_f16: k_per_tap must be a multiple of 64");
  LEMAS_REQUIRE(d.lda % 8 == 0 && d.ldw % 8 == 0, "lemas_gemm_f16: lda/ldw must be multiples of 8 (16 B rows)");
  LEMAS_REQUIRE(d.taps >= 1 && d.batches >= 1 && d.rows >= 1 && d.n >= 1, "lemas_gemm_f16: bad shape");
  LEMAS_REQUIRE((reinterpret_cast<uintptr_t>(d.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.w) & 15) == 0,
                "lemas_gemm_f16: operands must be 16-byte aligned");
  LEMAS_REQUIRE(d.seq_len > 0, "lemas_gemm_f16: seq_len must be set");

  GemmParams p;
  p.batches = d.batches; p.rows = d.rows; p.n = d.n;
  p.kc_per_tap = d.k_per_tap / BLOCK_K;
  p.k_iters = p.kc_per_tap * d.taps;
  p.tap_pad = d.tap_pad; p.w_tap_stride = d.w_tap_stride; p.group_cols = d.group_cols;
  p.tap_dil = d.tap_dilation > 0 ? d.tap_dilation : 1;
  p.bias = d.bias;
  p.out16 = static_cast<__half*>(d.out16); p.ld16 = d.ld16;
  p.out32 = d.out32; p.ld32 = d.ld32;
  p.resid = d.resid; p.ldr = d.ldr;
  p.gate = d.gate; p.gate_bstride = d.gate_bstride;
  p.row_valid = d.row_valid; p.seq_len = d.seq_len; p.row_limit = d.row_limit;
  {
    static const int red = [] { const char* e = getenv("LEMAS_G2_RED"); return e ? atoi(e) : 1; }();   // 0: load + add + store
    p.red_add = (red != 0 && d.epilogue == LEMAS_EPI_GATE_RESID_F32 && d.resid == d.out32 && d.ldr == d.ld32 &&
                 d.ln_out16 == nullptr) ? 1 : 0;
  }
  p.ln_scale = d.ln_scale; p.ln_out16 = static_cast<__half*>(d.ln_out16); p.ln_ld16 = d.ln_ld16; p.ln_stats = d.ln_stats;
  p.ln_stats_in = d.ln_stats_in; p.ln_parts = d.ln_parts; p.ln_uv = d.ln_uv; p.ln_step = d.ln_step;
  p.ln_inv_k = d.ln_k > 0 ? 1.0f / (float)d.ln_k : 0.f;
  p.rope = reinterpret_cast<const float2*>(d.rope); p.rope_cols = d.rope_cols; p.inner = d.inner;
  p.vt = static_cast<__half*>(d.vt); p.vt_ld = d.vt_ld;

  const bool need16 = d.epilogue == LEMAS_EPI_BIAS_F16 || d.epilogue == LEMAS_EPI_QKV_ROPE ||
                      d.epilogue == LEMAS_EPI_GELU_TANH_F16 || d.epilogue == LEMAS_EPI_GELU_ERF_F16 ||
                      d.epilogue == LEMAS_EPI_MISH_F16;
  const bool need32 = d.epilogue == LEMAS_EPI_GATE_RESID_F32 || d.epilogue == LEMAS_EPI_BIAS_F32 ||
                      d.epilogue == LEMAS_EPI_ADD_F32_F16 || d.epilogue == LEMAS_EPI_MISH_RESID_F32;
  LEMAS_REQUIRE(!need16 || (d.out16 && d.ld16 % 8 == 0), "lemas_gemm_f16: fp16 output missing or ld16 % 8 != 0");
  LEMAS_REQUIRE(!need32 || (d.out32 && d.ld32 % 4 == 0), "lemas_gemm_f16: fp32 output missing or ld32 % 4 != 0");
  if (need16 || d.epilogue == LEMAS_EPI_ADD_F32_F16 || d.epilogue == LEMAS_EPI_GATE_RESID_F32 ||
      d.epilogue == LEMAS_EPI_MISH_RESID_F32)
    LEMAS_REQUIRE(d.n % 32 == 0, "lemas_gemm_f16: this epilogue needs n % 32 == 0");
  if (d.epilogue == LEMAS_EPI_GATE_RESID_F32 || d.epilogue == LEMAS_EPI_ADD_F32_F16 ||
      d.epilogue == LEMAS_EPI_MISH_RESID_F32)
    LEMAS_REQUIRE(d.resid && d.ldr % 4 == 0, "lemas_gemm_f16: residual missing or ldr % 4 != 0");
  if (d.epilogue == LEMAS_EPI_QKV_ROPE)
    LEMAS_REQUIRE(d.rope && d.vt && d.inner % 64 == 0 && d.n == 3 * d.inner && d.rope_cols % 64 == 0,
                  "lemas_gemm_f16: QKV epilogue needs rope table, vt buffer and n == 3*inner");

  if (pair_enabled() && gemm2_eligible(d)) return gemm2_launch(d, p, stream);
  {
    static const int red1 = [] { const char* e = getenv("LEMAS_G1_RED"); return e ? atoi(e) : 1; }();   // 0: load + add + store
    p.red_add = (red1 != 0 && d.epilogue == LEMAS_EPI_GATE_RESID_F32 && d.resid == d.out32 && d.ldr == d.ld32 &&
                 d.n % 32 == 0) ? 2 : 0;
  }
  LEMAS_REQUIRE(d.ln_out16 == nullptr && d.ln_stats_in == nullptr,
                "lemas_gemm_f16: the folded LayerNorm needs the CTA-pair kernel (block_n 256, n % 256 == 0)");

  CUtensorMap tmA, tmW;
  {
    uint64_t dims[3] = {(uint64_t)d.a_cols, (uint64_t)d.rows, (uint64_t)d.batches};
    uint64_t strides[2] = {(uint64_t)d.lda * 2, (uint64_t)d.rows * d.lda * 2};
    uint32_t box[3] = {BLOCK_K, BLOCK_M, 1};
    LEMAS_TRY(make_tensor_map_f16(&tmA, d.a, 3, dims, strides, box));
  }
  {
    uint64_t dims[2] = {(uint64_t)d.ldw, (uint64_t)d.w_rows};
    uint64_t strides[1] = {(uint64_t)d.ldw * 2};
    uint32_t box[2] = {BLOCK_K, (uint32_t)d.block_n};
    LEMAS_TRY(make_tensor_map_f16(&tmW, d.w, 2, dims, strides, box));
  }

  switch (d.epilogue) {
    case LEMAS_EPI_BIAS_F16: return dispatch_bn<LEMAS_EPI_BIAS_F16>(d.block_n, tmA, tmW, p, d.max_ctas, stream);
    case LEMAS_EPI_QKV_ROPE: return dispatch_bn<LEMAS_EPI_QKV_ROPE>(d.block_n, tmA, tmW, p, d.max_ctas, stream);
    case LEMAS_EPI_GELU_TANH_F16:
      return dispatch_bn<LEMAS_EPI_GELU_TANH_F16>(d.block_n, tmA, tmW, p, d.max_ctas, stream);
    case LEMAS_EPI_GELU_ERF_F16:
      return dispatch_bn<LEMAS_EPI_GELU_ERF_F16>(d.block_n, tmA, tmW, p, d.max_ctas, stream);
    case LEMAS_EPI_GATE_RESID_F32:
      return dispatch_bn<LEMAS_EPI_GATE_RESID_F32>(d.block_n, tmA, tmW, p, d.max_ctas, stream);
    case LEMAS_EPI_BIAS_F32: return dispatch_bn<LEMAS_EPI_BIAS_F32>(d.block_n, tmA, tmW, p, d.max_ctas, stream);
    case LEMAS_EPI_ADD_F32_F16:
      return dispatch_bn<LEMAS_EPI_ADD_F32_F16>(d.block_n, tmA, tmW, p, d.max_ctas, stream);
    case LEMAS_EPI_MISH_F16: return dispatch_bn<LEMAS_EPI_MISH_F16>(d.block_n, tmA, tmW, p, d.max_ctas, stream);
    case LEMAS_EPI_MISH_RESID_F32:
      return dispatch_bn<LEMAS_EPI_MISH_RESID_F32>(d.block_n, tmA, tmW, p, d.max_ctas, stream);
  }
  return fail(LEMAS_ERR_INVALID, "lemas_gemm_f16: unknown epilogue");
}

}  // namespace lemas

extern "C" int lemas_gemm_f16(const lemas_gemm_desc* d, void* stream) {
  if (!d) return lemas::fail(LEMAS_ERR_INVALID, "lemas_gemm_f16: null descriptor");
  return lemas::gemm_launch(*d, static_cast<cudaStream_t>(stream));
}
