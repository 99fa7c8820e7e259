// v9: ONE 64-key pipeline per CTA, FOUR CTAs per SM (experimental, LEMAS_ATT_VARIANT=40).
//
// attention.cu (v3) runs two independent 64-key half pipelines inside a 128-query CTA and merges them at the end; its
// measured cost model is t = 2.0 us + 1.58 us x (128-key blocks) per CTA (59.3 us at C2's 18 blocks, 238 us at C4's 6:
// the fixed part is 6.5 % / 17 % of the CTA's life).  Here a CTA owns a 128-query tile and walks ALL keys in 64-key
// blocks with a single pipeline: twice the blocks per CTA for the same fixed cost, no merge of halves, no exchange
// buffer, 128 TMEM columns (S 64 + O 64), 48 KB of shared memory (two-slot K / V^T rings) and 160 threads — the TMA
// producer and the MMA issuer are ONE warp, so that four CTAs (= v3's four pipelines per SM) fit the register file at 102
// registers per thread.  (First version: 192 threads, three-slot rings, three CTAs per SM: 13 % behind v3 in steady state.)
// Everything inside the pipeline — P written back over the scores in TMEM, TS-form P V, lazy rescaling, packed
// arithmetic, a quarter of the exponentials on the FMA pipe — is v3's.
#include <type_traits>

#include "att_common.cuh"

namespace lemas {

constexpr int A9_THREADS = 160;   // warp 0 TMA + MMA, warps 1-4 softmax (thread == query row)
constexpr int A9_BM = 128;
constexpr int A9_BN = 64;         // keys per block
constexpr int A9_D = 64;
constexpr int A9_STAGES = 2;
constexpr int A9_Q_BYTES = A9_BM * A9_D * 2;     // 16 KB
constexpr int A9_K_BYTES = A9_BN * A9_D * 2;     // 8 KB
constexpr int A9_V_BYTES = A9_D * A9_BN * 2;     // 8 KB
constexpr int A9_OFF_K = A9_Q_BYTES;
constexpr int A9_OFF_V = A9_OFF_K + A9_STAGES * A9_K_BYTES;
constexpr int A9_OFF_BAR = A9_OFF_V + A9_STAGES * A9_V_BYTES;
constexpr int A9_SMEM = A9_OFF_BAR + 256;        // 48.25 KB: four CTAs per SM

constexpr int B9_Q = 0, B9_KF = 1, B9_KE = 3, B9_VF = 5, B9_VE = 7, B9_SF = 9, B9_PF = 10, B9_OF = 11, B9_COUNT = 12;
constexpr float A9_RESCALE_LOG2 = 8.0f;
#ifndef A9_POLY_EVERY
#define A9_POLY_EVERY 4
#endif

__global__ void __launch_bounds__(A9_THREADS, 4)
attention9_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmVT, const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A9_OFF_BAR);
  uint64_t* q_full = bars + B9_Q;
  uint64_t* k_full = bars + B9_KF;
  uint64_t* k_empty = bars + B9_KE;
  uint64_t* v_full = bars + B9_VF;
  uint64_t* v_empty = bars + B9_VE;
  uint64_t* s_full = bars + B9_SF;
  uint64_t* p_full = bars + B9_PF;
  uint64_t* o_full = bars + B9_OF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B9_COUNT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * A9_BM;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int kvl = p.kv_len ? min(__ldg(p.kv_len + b), p.seq) : p.seq;
  const int n_blocks = (kvl + A9_BN - 1) / A9_BN;
  if (q0 >= kvl) return;   // tiles made only of padding rows (see attention.cu)

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("lemas attention: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVT);
    mbar_init(q_full, 1);
    for (int s = 0; s < A9_STAGES; ++s) {
      mbar_init(k_full + s, 1);
      mbar_init(k_empty + s, 1);
      mbar_init(v_full + s, 1);
      mbar_init(v_empty + s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);   // one arrival per softmax warp
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_s = tmem_base;          // S (64 fp32 columns) / P (32 columns of fp16 pairs)
  const uint32_t tmem_o = tmem_base + 64;     // O

  if (warp == 0) {
    // ---- control warp: TMA producer and MMA issuer in one (the loads of block j + 2 are issued right after the MMAs
    // of block j: their ring slots are released by exactly those MMAs)
    constexpr uint32_t idesc = umma_idesc_f16(A9_BM, 64);
    const uint32_t sq = smem_u32(smem);
    auto load_k = [&](int j) {
      const int s = j % A9_STAGES;
      mbar_arrive_expect_tx(k_full + s, A9_K_BYTES);
      tma_load_3d(smem + A9_OFF_K + s * A9_K_BYTES, &tmK, k_full + s, p.inner + h * A9_D, j * A9_BN, b);
    };
    auto load_v = [&](int j) {
      const int s = j % A9_STAGES;
      mbar_arrive_expect_tx(v_full + s, A9_V_BYTES);
      tma_load_3d(smem + A9_OFF_V + s * A9_V_BYTES, &tmVT, v_full + s, j * A9_BN, 0, b * p.heads + h);
    };
    auto issue_s = [&](int j) {  // S(j) = Q K_j^T
      const int s = j % A9_STAGES;
      const uint64_t adesc = umma_desc_sw128(sq), bdesc = umma_desc_sw128(smem_u32(smem + A9_OFF_K + s * A9_K_BYTES));
#pragma unroll
      for (int k = 0; k < A9_D / 16; ++k) umma_f16_ss(tmem_s, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
      umma_commit(s_full);
      umma_commit(k_empty + s);
    };
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, A9_Q_BYTES);
      tma_load_3d(smem, &tmQ, q_full, h * A9_D, q0, b);
      for (int j = 0; j < A9_STAGES && j < n_blocks; ++j) { load_k(j); load_v(j); }
    }
    __syncwarp();
    ATT_WAIT_P(q_full, 0, 3, 0);
    ATT_WAIT_P(k_full + 0, 0, 4, 0);
    tc_fence_after();
    if (elect_one()) issue_s(0);
    __syncwarp();
    for (int j = 0; j < n_blocks; ++j) {
      const bool last = j + 1 == n_blocks;
      const int s = j % A9_STAGES;
      ATT_WAIT_P(v_full + s, (j / A9_STAGES) & 1, 5, j);
      if (!last) ATT_WAIT_P(k_full + ((j + 1) % A9_STAGES), ((j + 1) / A9_STAGES) & 1, 4, j + 1);
      ATT_WAIT_P(p_full, j & 1, 6, j);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + A9_OFF_V + s * A9_V_BYTES));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)   // O (+)= P(j) V_j; A = P from TMEM: 8 columns (16 fp16) per K16 step
          umma_f16_ts(tmem_o, tmem_s + 8 * ks, bdesc + 2 * ks, idesc, (j | ks) != 0 ? 1u : 0u);
        umma_commit(v_empty + s);
        if (last) umma_commit(o_full);
        else issue_s(j + 1);   // overwrites P(j): executes behind the P V just issued
      }
      __syncwarp();
      if (j + A9_STAGES < n_blocks) {   // refill the two slots block j occupied (K: free since S(j), V: once P V(j) retires)
        ATT_WAIT_P(k_empty + s, (j / A9_STAGES) & 1, 1, j);
        if (elect_one()) load_k(j + A9_STAGES);
        __syncwarp();
        ATT_WAIT_P(v_empty + s, (j / A9_STAGES) & 1, 2, j);
        if (elect_one()) load_v(j + A9_STAGES);
        __syncwarp();
      }
    }
  } else {
    const int sub = warp & 3;
    const int r = sub * 32 + lane;
    const uint32_t lane_addr = uint32_t(sub * 32) << 16;
    const float c = 0.125f * 1.4426950408889634f;
    float m_ref = -INFINITY;
    float l_run = 0.f;
    const uint32_t sb = smem_u32(smem);
    const uint32_t a_sfull = sb + A9_OFF_BAR + B9_SF * 8, a_pfull = sb + A9_OFF_BAR + B9_PF * 8;
    const uint32_t t_s = tmem_s + lane_addr;
    const uint32_t t_o = tmem_o + lane_addr;
    const bool rows_dead = q0 + sub * 32 >= p.seq;
    for (int j = 0; j < n_blocks; ++j) {
      if (rows_dead) {
        ATT_WAIT_A(a_sfull, j & 1, 12, j);
        __syncwarp();
        if (lane == 0) mbar_arrive_s(a_pfull);
        continue;
      }
      const int valid = min(max(kvl - j * A9_BN, 0), 64);
      ATT_WAIT_A(a_sfull, j & 1, 8, j);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(t_s, s0);
      tmem_ld_32x32(t_s + 32, s1);
      tmem_ld_wait();

      float mx = -INFINITY;
      if (valid == 64) {  // four independent FMNMX3 chains of depth 8 instead of one of depth 32
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; ++i) m4[i & 3] = fmax3f(m4[i & 3], __uint_as_float(s0[i]), __uint_as_float(s1[i]));
        mx = fmaxf(fmax3f(m4[0], m4[1], m4[2]), m4[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < valid) mx = fmaxf(mx, __uint_as_float(s0[i]));
          if (i + 32 < valid) mx = fmaxf(mx, __uint_as_float(s1[i]));
        }
      }
      // lazy rescale: advance the reference max only when this block exceeds it by more than 2^8 (warp-uniform
      // decision, tcgen05.ld/st are warp-collective)
      const bool grow = (mx - m_ref) * c > A9_RESCALE_LOG2;  // also true for the first finite max (m_ref = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mx : m_ref;
        const float alpha = (m_ref == -INFINITY) ? 0.f : ex2f((m_ref - m_new) * c);
        l_run *= alpha;
        if (j > 0) {  // O_half holds the sum of blocks < j (retired, see the s_full wait): rescale it in TMEM
#pragma unroll 1
          for (int cc = 0; cc < A9_D; cc += 8) {  // narrow chunks: S_j (64 registers) stays live across this
            uint32_t v[8];
            tmem_ld_32x32_x8(t_o + cc, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32_x8(t_o + cc, v);
          }
        }
        m_ref = m_new;
      }
      const float mc = (m_ref == -INFINITY) ? 0.f : m_ref * c;

      // Per key pair: one FFMA2 (scale, subtract the reference max), two MUFU.EX2, one FADD2 into one of four
      // independent packed row-sum accumulators, one F2FP pack.
      uint64_t rs2[4] = {0ull, 0ull, 0ull, 0ull};   // bit pattern of (0.f, 0.f)
      uint32_t pk[32];
      const uint64_t c2 = f32x2(c, c), nmc2 = f32x2(-mc, -mc);
      auto exp_block = [&](auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;  // full half-block: no per-element masking code at all
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = 2 * i;
          if (!kFull && col >= valid) {  // warp-uniform: masked key pairs cost no SFU work
            pk[i] = 0u;
            continue;
          }
          float x0, x1;
          f32x2_split(ffma2(f32x2(__uint_as_float(col < 32 ? s0[col & 31] : s1[col & 31]),
                                  __uint_as_float(col + 1 < 32 ? s0[(col + 1) & 31] : s1[(col + 1) & 31])),
                            c2, nmc2), x0, x1);
          float e0, e1;
          if (kFull && A9_POLY_EVERY > 0 && (i % (A9_POLY_EVERY > 0 ? A9_POLY_EVERY : 1)) == 0) {
            // exp2 on the FMA / ALU pipes for one key pair in A9_POLY_EVERY (the SFU is the contended unit):
            // x = n + f, n = round(x) via the 1.5 * 2^23 magic constant, f in [-0.5, 0.5]; 2^f by a degree-3 minimax
            // polynomial (max relative error 7.5e-5, below the fp16 rounding of P); 2^n added into the exponent field.
            // x <= 8 by the lazy-rescale bound; the clamp keeps n inside the exponent range (result < 2^-125 ~ 0).
            const uint64_t xc = f32x2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
            const uint64_t t2 = fadd2(xc, f32x2(12582912.f, 12582912.f));
            const uint64_t f2 = ffma2(fadd2(t2, f32x2(-12582912.f, -12582912.f)), f32x2(-1.f, -1.f), xc);
            uint64_t p2 = ffma2(f32x2(0.055171460f, 0.055171460f), f2, f32x2(0.24261086f, 0.24261086f));
            p2 = ffma2(p2, f2, f32x2(0.69326097f, 0.69326097f));
            p2 = ffma2(p2, f2, f32x2(0.99992812f, 0.99992812f));
            float p0, p1, t0, t1;
            f32x2_split(p2, p0, p1);
            f32x2_split(t2, t0, t1);
            e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
            e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
          } else {
            e0 = ex2f(x0);
            e1 = ex2f(x1);
          }
          if (!kFull && col + 1 >= valid) e1 = 0.f;
          rs2[i & 3] = fadd2(rs2[i & 3], f32x2(e0, e1));
          pk[i] = pack_half2(e0, e1);
        }
      };
      if (valid == 64) exp_block(std::true_type{}); else exp_block(std::false_type{});
      tmem_st_32x32(t_s, pk);   // P(j) over the first 32 of the 64 score columns: column k holds keys (2k, 2k+1)
      {
        float lo, hi, lo2, hi2;
        f32x2_split(fadd2(rs2[0], rs2[1]), lo, hi);
        f32x2_split(fadd2(rs2[2], rs2[3]), lo2, hi2);
        l_run += (lo + hi) + (lo2 + hi2);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_s(a_pfull);
    }

    // ---- normalise and store: this warp's 32 rows, all 64 head-dim columns
    ATT_WAIT_A(sb + A9_OFF_BAR + B9_OF * 8, 0, 10, 0);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const int row = q0 + r;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t o[32];
      tmem_ld_32x32(t_o + half * 32, o);
      tmem_ld_wait();
      if (row < p.seq) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((long)b * p.seq + row) * p.inner + h * A9_D + half * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 w;
          w.x = pack_half2(__uint_as_float(o[8 * u + 0]) * inv, __uint_as_float(o[8 * u + 1]) * inv);
          w.y = pack_half2(__uint_as_float(o[8 * u + 2]) * inv, __uint_as_float(o[8 * u + 3]) * inv);
          w.z = pack_half2(__uint_as_float(o[8 * u + 4]) * inv, __uint_as_float(o[8 * u + 5]) * inv);
          w.w = pack_half2(__uint_as_float(o[8 * u + 6]) * inv, __uint_as_float(o[8 * u + 7]) * inv);
          dst[u] = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem_base);
}

int attention_v9_launch(const void* qk, int32_t ld_qk, const void* vt, int32_t vt_ld, const AttnParams& p, int batch,
                        void* stream) {
  CUtensorMap tmQ, tmK, tmVT;
  {
    uint64_t dims[3] = {(uint64_t)2 * p.inner, (uint64_t)p.seq, (uint64_t)batch};
    uint64_t strides[2] = {(uint64_t)ld_qk * 2, (uint64_t)p.seq * ld_qk * 2};
    uint32_t box_q[3] = {64, A9_BM, 1}, box_k[3] = {64, A9_BN, 1};
    LEMAS_TRY(make_tensor_map_f16(&tmQ, qk, 3, dims, strides, box_q));
    LEMAS_TRY(make_tensor_map_f16(&tmK, qk, 3, dims, strides, box_k));
  }
  {
    uint64_t dims[3] = {(uint64_t)p.seq, (uint64_t)A9_D, (uint64_t)batch * p.heads};
    uint64_t strides[2] = {(uint64_t)vt_ld * 2, (uint64_t)A9_D * vt_ld * 2};
    uint32_t box[3] = {64, A9_D, 1};
    LEMAS_TRY(make_tensor_map_f16(&tmVT, vt, 3, dims, strides, box));
  }
  static unsigned long long configured = 0;
  LEMAS_CUDA_OK(ensure_dynamic_smem(attention9_kernel, A9_SMEM, configured));
  dim3 grid((p.seq + A9_BM - 1) / A9_BM, p.heads, batch);
  LEMAS_CUDA_OK(launch_pdl(attention9_kernel, grid, dim3(A9_THREADS), A9_SMEM, (cudaStream_t)stream, tmQ, tmK, tmVT, p));
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

}  // namespace lemas
