// Non-causal self-attention for head_dim 64 on tcgen05 tensor cores (sm_100a), key-padding aware.
//
//   O = softmax(Q K^T / 8 + keymask) V        modules.py:483-491 (F.scaled_dot_product_attention call site)
//
// What bounds this kernel on B200: with head_dim 64 a 128 x 128 score block costs 512 tensor-core clocks (QK^T + PV)
// but 16 384 exponentials, and the SFU delivers 16 exp2 / clk / SM (measured, profiles/r01b_mufu_exp2_throughput.log;
// the packed f16x2 / bf16x2 forms are split into two MUFU ops and gain nothing) = 1024 clocks.  The design goal is
// therefore to keep the SFU saturated: many independent softmax warps, no per-block work besides max / exp / sum / pack.
//
// One CTA = 128 query rows of one (batch, head); two CTAs are co-resident per SM.  Roles (320 threads):
//   warp 0    TMA producer : Q tile once, then K tile [128 keys x 64] and V^T tile [64 x 128 keys] per KV block
//                            into two 2-slot rings (128B swizzle, mbarrier tx-count); K slots recycle after S = Q K^T,
//                            V slots after P V, so K runs a block further ahead
//   warp 1    MMA issuer   : S = Q K^T (M128 N128 K16 x4) into TMEM cols [0,128); S for block j+1 is issued while the
//                            softmax of block j runs.  P V is issued as TWO independent streams, one per 64-key half
//                            of the block: O_A += P[:, 0:64] V[0:64], O_B += P[:, 64:128] V[64:128] (M128 N64 K16 x4
//                            each), accumulated IN TMEM across all KV blocks (cols [128,192) and [192,256)).
//   warps 2-9 softmax      : warps 2-5 own key half A, warps 6-9 key half B; thread == query row (tcgen05.ld 32x32b
//                            gives each lane one row).  Each half runs its own online softmax (own running max and
//                            row sum) over its 64 keys of every block — intra-CTA split-KV — so the two warps that
//                            share a row never synchronise inside the loop; the halves are merged once at the end:
//                            O = (w_A O_A + w_B O_B) / (w_A l_A + w_B l_B),  w_X = 2^((m_X - max(m_A, m_B)) c).
//                            The accumulators are rescaled lazily: O_X and l_X keep the scale of a reference max that
//                            is only advanced when a block's max exceeds it by more than 2^8 (P then stays <= 256,
//                            exact in fp16 terms); after the first blocks this almost never fires, so the steady
//                            state per block is: one TMEM read of S, max, exp2, sum, fp16 pack, swizzled smem store.
// Fully masked KV blocks (keys >= kv_len) are skipped; the partial block is masked to zero probability.
// K comes from the fused QKV buffer [b*seq, ld_qk]; V is read from the transposed copy [b, head, 64, vt_ld] that
// the QKV GEMM epilogue writes, so both MMAs use K-major operands.
#include <type_traits>

#include "common.h"
#include "ptx.cuh"

namespace lemas {

constexpr int ATT_THREADS = 320;
constexpr int ATT_BM = 128;   // query rows per CTA
constexpr int ATT_BN = 128;   // keys per KV block
constexpr int ATT_D = 64;
constexpr int ATT_STAGES = 2;

constexpr int ATT_Q_BYTES = ATT_BM * ATT_D * 2;          // 16 KB
constexpr int ATT_K_BYTES = ATT_BN * ATT_D * 2;          // 16 KB
constexpr int ATT_V_BYTES = ATT_D * ATT_BN * 2;          // 16 KB (two 8 KB halves of 64 keys)
constexpr int ATT_P_BYTES = ATT_BM * ATT_BN * 2;         // 32 KB (two 16 KB halves of 64 keys)
constexpr int ATT_KV_STAGE = ATT_K_BYTES + ATT_V_BYTES;
constexpr int ATT_OFF_KV = ATT_Q_BYTES;
constexpr int ATT_OFF_P = ATT_OFF_KV + ATT_STAGES * ATT_KV_STAGE;
constexpr int ATT_OFF_XCH = ATT_OFF_P;                   // float2 [2 halves][128 rows] (reference max, row sum):
                                                         // reuses the P buffer after the last P V has retired
constexpr int ATT_OFF_BAR = ATT_OFF_P + ATT_P_BYTES;
constexpr int ATT_SMEM = ATT_OFF_BAR + 128;              // 112.1 KB: two CTAs per SM

constexpr float ATT_RESCALE_LOG2 = 8.0f;  // advance the reference max only past 2^8 growth

struct AttnParams {
  long long* trace;   // debug: clock64 stamps of CTA (1,0,0) [warp][block][8]; nullptr in production
  const int* kv_len;
  __half* out;
  int seq, heads, inner;
};

DEVI float fmax3f(float a, float b, float c) {  // 3-input max: one FMNMX3 on sm_100
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
DEVI float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Packed fp32 pairs (sm_100 FFMA2 / FADD2): one issue slot for two lanes of arithmetic.  The softmax warps are
// issue-limited, so the scale-and-subtract and the row sum are done on register pairs.
DEVI uint64_t f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
DEVI void f32x2_split(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
DEVI uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
DEVI uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVT,
                 const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;                  // [2]  K and V^T travel through separate 2-slot rings: a K slot is
  uint64_t* k_empty = bars + 3;                 // [2]  free as soon as S_j = Q K_j^T has retired, a V slot only after
  uint64_t* v_full = bars + 5;                  // [2]  P V_j — so K_{j+2} streams in a whole block earlier than a
  uint64_t* v_empty = bars + 7;                 // [2]  shared ring would allow and S_{j+1} is never late
  uint64_t* s_full = bars + 9;
  uint64_t* s_empty = bars + 10;
  uint64_t* p_full = bars + 11;                 // [2] per key half
  uint64_t* o_full = bars + 13;                 // [2] per key half
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef LEMAS_ATT_TRACE  // per-CTA record after the warp stamps: [cta][8] = smid, globaltimer at entry / loop / merge / exit
  long long* cta_rec = p.trace ? p.trace + 8 * 32 * 8 +
      8 * ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) : nullptr;
  auto gtime = [] { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
  if (cta_rec && threadIdx.x == 64) {
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    cta_rec[0] = smid; cta_rec[1] = gtime();
  }
#endif
  const int q0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int kvl = p.kv_len ? min(__ldg(p.kv_len + b), p.seq) : p.seq;
  const int n_blocks = (kvl + ATT_BN - 1) / ATT_BN;
  // Query tiles made only of padding rows: their output is zeroed by the row mask of the to_out epilogue
  // (modules.py:499-501), so nothing needs to be computed.
  if (q0 >= kvl) return;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("lemas attention: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmVT);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT_STAGES; ++s) {
      mbar_init(k_full + s, 1);
      mbar_init(k_empty + s, 1);
      mbar_init(v_full + s, 1);
      mbar_init(v_empty + s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 8);            // one arrival per softmax warp
    for (int x = 0; x < 2; ++x) {
      mbar_init(p_full + x, 4);         // one arrival per softmax warp of the half
      mbar_init(o_full + x, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();  // prologue overlapped the previous kernel's tail; q/k/v are visible from here on
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_o = tmem_base + ATT_BN;   // + 64 * half

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, ATT_Q_BYTES);
      tma_load_3d(smem, &tmQK, q_full, h * ATT_D, q0, b);
      for (int j = 0; j < n_blocks; ++j) {
        const int s = j & 1;
        const uint32_t ph = ((j >> 1) & 1) ^ 1;
        uint8_t* sk = smem + ATT_OFF_KV + s * ATT_KV_STAGE;
        mbar_wait(k_empty + s, ph);
        mbar_arrive_expect_tx(k_full + s, ATT_K_BYTES);
        tma_load_3d(sk, &tmQK, k_full + s, p.inner + h * ATT_D, j * ATT_BN, b);
        mbar_wait(v_empty + s, ph);
        mbar_arrive_expect_tx(v_full + s, ATT_V_BYTES);
        tma_load_3d(sk + ATT_K_BYTES, &tmVT, v_full + s, j * ATT_BN, 0, b * p.heads + h);
        tma_load_3d(sk + ATT_K_BYTES + ATT_V_BYTES / 2, &tmVT, v_full + s, j * ATT_BN + 64, 0, b * p.heads + h);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc_f16(ATT_BM, ATT_BN);
    constexpr uint32_t idesc_o = umma_idesc_f16(ATT_BM, ATT_D);
    const uint32_t sq = smem_u32(smem);
    const uint32_t sp = smem_u32(smem + ATT_OFF_P);
    auto issue_s = [&](int j) {  // S = Q K_j^T
      const uint32_t sk = smem_u32(smem + ATT_OFF_KV + (j & 1) * ATT_KV_STAGE);
      const uint64_t adesc = umma_desc_sw128(sq), bdesc = umma_desc_sw128(sk);
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k) umma_f16_ss(tmem_s, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
      umma_commit(s_full);
      umma_commit(k_empty + (j & 1));
    };
    mbar_wait(q_full, 0);
    mbar_wait(k_full + 0, 0);
    tc_fence_after();
    if (elect_one()) issue_s(0);
    __syncwarp();
    for (int j = 0; j < n_blocks; ++j) {
      if (j + 1 < n_blocks) {
        mbar_wait(k_full + ((j + 1) & 1), ((j + 1) >> 1) & 1);
        mbar_wait(s_empty, j & 1);  // every softmax thread has pulled its part of S_j out of TMEM
        tc_fence_after();
        if (elect_one()) issue_s(j + 1);
        __syncwarp();
      }
      const uint32_t sv = smem_u32(smem + ATT_OFF_KV + (j & 1) * ATT_KV_STAGE + ATT_K_BYTES);
      mbar_wait(v_full + (j & 1), (j >> 1) & 1);
#pragma unroll
      for (int x = 0; x < 2; ++x) {  // O_x (+)= P_j[:, 64x : 64x+64] V_j[64x : 64x+64]
        mbar_wait(p_full + x, j & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = umma_desc_sw128(sp + x * (ATT_P_BYTES / 2));
          const uint64_t bdesc = umma_desc_sw128(sv + x * (ATT_V_BYTES / 2));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16_ss(tmem_o + x * ATT_D, adesc + 2 * ks, bdesc + 2 * ks, idesc_o, (j | ks) != 0 ? 1u : 0u);
          umma_commit(o_full + x);
          if (x == 1) umma_commit(v_empty + (j & 1));
        }
        __syncwarp();
      }
    }
  } else {
    const int sub = warp & 3;          // TMEM sub-partition: lanes [32*sub, 32*sub+32)
    const int half = (warp - 2) >> 2;  // key half of every KV block this warp owns
    const int r = sub * 32 + lane;     // query row inside the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(sub * 32) << 16;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float m_ref = -INFINITY;           // max the accumulators O_half / l are currently scaled by
    float l_run = 0.f;
    // 32-bit shared-window addresses (one register each, immediates for the rest): the softmax loop is issue-bound,
    // every instruction of address arithmetic or pointer conversion in it costs throughput
    const uint32_t sb = smem_u32(smem);
    const uint32_t a_sfull = sb + ATT_OFF_BAR + 9 * 8, a_sempty = sb + ATT_OFF_BAR + 10 * 8;
    const uint32_t a_pfull = sb + ATT_OFF_BAR + (11 + half) * 8, a_ofull = sb + ATT_OFF_BAR + (13 + half) * 8;
    const uint32_t a_prow = sb + ATT_OFF_P + half * (ATT_P_BYTES / 2) + (r >> 3) * 1024 + (r & 7) * 128;
    const uint32_t xor7 = (r & 7) << 4;
    const uint32_t t_s = tmem_s + lane_addr + half * 64;
    const uint32_t t_o = tmem_o + lane_addr + half * ATT_D;

    // Warps whose 32 query rows all lie beyond the sequence (last query tile) keep the barrier protocol going but do
    // no softmax work: their P rows only feed output rows that are never stored.
    const bool rows_dead = q0 + sub * 32 >= p.seq;
#ifdef LEMAS_ATT_TRACE
    if (cta_rec && threadIdx.x == 64) cta_rec[2] = gtime();
#endif
    for (int j = 0; j < n_blocks; ++j) {
      if (rows_dead) {
        mbar_wait_lean(a_sfull, j & 1);
        if (j > 0) mbar_wait_lean(a_ofull, (j - 1) & 1);  // never run a p_full phase ahead of the live warps
        if (lane == 0) { mbar_arrive_s(a_sempty); mbar_arrive_s(a_pfull); }
        continue;
      }
      const int valid = min(max(kvl - j * ATT_BN - half * 64, 0), 64);  // keys of this half-block that exist
#ifdef LEMAS_ATT_TRACE  // clock64 stamps of one CTA (tools/trace_att.py); costs ~8 % of the kernel, off by default
      // traced CTA = linear id stored in the unused stamp slot [warp 0][block 0][7]
      const bool tr = p.trace != nullptr && lane == 0 && j < 32 &&
                      (long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x == p.trace[7];
      long long* tp = p.trace + ((warp - 2) * 32 + j) * 8;
#define ATT_STAMP(i) do { if (tr) tp[i] = clock64(); } while (0)
#else
#define ATT_STAMP(i) do { } while (0)
#endif
      ATT_STAMP(0);
      if (j == 0) mbar_wait_lean(a_sfull, 0);  // later blocks: S_j was awaited before P_{j-1} was stored (below)
      ATT_STAMP(1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(t_s, s0);
      tmem_ld_32x32(t_s + 32, s1);
      tmem_ld_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive_s(a_sempty);  // S_j is in registers (tcgen05.wait::ld is warp-wide): S_{j+1} may land
      ATT_STAMP(2);

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
      const bool grow = (mx - m_ref) * c > ATT_RESCALE_LOG2;  // also true for the first finite max (m_ref = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mx : m_ref;
        const float alpha = (m_ref == -INFINITY) ? 0.f : ex2f((m_ref - m_new) * c);
        l_run *= alpha;
        if (j > 0) {  // O_half holds the sum of blocks < j: rescale it in TMEM once P V_{j-1} has retired
          mbar_wait_lean(a_ofull, (j - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < ATT_D; cc += 8) {  // narrow chunks: S_j (64 registers) stays live across this
            uint32_t v[8];
            tmem_ld_32x32_x8(t_o + cc, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32_x8(t_o + cc, v);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        m_ref = m_new;
      }
      const float mc = (m_ref == -INFINITY) ? 0.f : m_ref * c;
      ATT_STAMP(3);

      // exp2 of the whole half-row into packed fp16 registers first; only then wait for the P buffer (free once
      // P V_{j-1} has read it) — waiting before the exponentials re-synchronised the four warps of a half every block
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
          const float e0 = ex2f(x0);
          float e1 = ex2f(x1);
          if (!kFull && col + 1 >= valid) e1 = 0.f;
          rs2[i & 3] = fadd2(rs2[i & 3], f32x2(e0, e1));
          pk[i] = pack_half2(e0, e1);
        }
      };
      if (valid == 64) exp_block(std::true_type{}); else exp_block(std::false_type{});
      float rs4[4];
      {
        float lo, hi;
        const uint64_t t01 = fadd2(rs2[0], rs2[1]), t23 = fadd2(rs2[2], rs2[3]);
        f32x2_split(t01, lo, hi);
        rs4[0] = lo; rs4[1] = hi;
        f32x2_split(t23, lo, hi);
        rs4[2] = lo; rs4[3] = hi;
      }
      ATT_STAMP(4);
      // The P buffer is free once P V_{j-1} has retired.  The MMA warp issues S_{j+1} after P V_{j-1} and commits it
      // to s_full, and a commit tracks every MMA issued before it — so ONE wait on s_full(j+1) covers both "P is
      // free" and "S_{j+1} is ready" (a successful mbarrier wait costs ~120 clk of pure latency in this loop).
      if (j + 1 < n_blocks) mbar_wait_lean(a_sfull, (j + 1) & 1);
      else if (j > 0) mbar_wait_lean(a_ofull, (j - 1) & 1);
      ATT_STAMP(5);
#pragma unroll
      for (int u = 0; u < 8; ++u)  // 16-byte units of the 128-byte (64 keys x fp16) swizzled row
        sts128(a_prow + ((u << 4) ^ xor7), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
      l_run += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      fence_proxy_async_smem();  // make the generic-proxy P stores visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive_s(a_pfull);
      ATT_STAMP(6);
    }

    // ---- merge the two key halves and normalise
#ifdef LEMAS_ATT_TRACE
    if (cta_rec && threadIdx.x == 64) cta_rec[3] = gtime();
#endif
    mbar_wait_lean(sb + ATT_OFF_BAR + 13 * 8, (n_blocks - 1) & 1);
    mbar_wait_lean(sb + ATT_OFF_BAR + 14 * 8, (n_blocks - 1) & 1);
    tc_fence_after();
    float2* xch = reinterpret_cast<float2*>(smem + ATT_OFF_XCH);  // P buffer: free now that every P V has retired
    xch[half * ATT_BM + r] = make_float2(m_ref, l_run);
    named_bar_sync(1 + sub, 64);  // the two warps that share these 32 rows
    const float2 other = xch[(half ^ 1) * ATT_BM + r];
    const float m_all = fmaxf(m_ref, other.x);
    const float w_me = (m_ref == -INFINITY) ? 0.f : ex2f((m_ref - m_all) * c);
    const float w_ot = (other.x == -INFINITY) ? 0.f : ex2f((other.x - m_all) * c);
    const float inv = 1.0f / (w_me * l_run + w_ot * other.y);
    const float wa = (half == 0 ? w_me : w_ot) * inv, wb = (half == 0 ? w_ot : w_me) * inv;
    uint32_t oa[32], ob[32];  // this warp outputs head-dim columns [32*half, 32*half+32)
    tmem_ld_32x32(tmem_o + lane_addr + half * 32, oa);
    tmem_ld_32x32(tmem_o + lane_addr + ATT_D + half * 32, ob);
    tmem_ld_wait();
    const int row = q0 + r;
    if (row < p.seq) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((long)b * p.seq + row) * p.inner + h * ATT_D + half * 32);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o[i] = __uint_as_float(oa[8 * u + i]) * wa + __uint_as_float(ob[8 * u + i]) * wb;
        uint4 w;
        w.x = pack_half2(o[0], o[1]);
        w.y = pack_half2(o[2], o[3]);
        w.z = pack_half2(o[4], o[5]);
        w.w = pack_half2(o[6], o[7]);
        dst[u] = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
#ifdef LEMAS_ATT_TRACE
  if (cta_rec && threadIdx.x == 64) cta_rec[4] = gtime();
#endif
}

}  // namespace lemas

using namespace lemas;

static long long* g_att_trace = nullptr;
// debug aid (not part of the public header): device buffer of 8 warps x 32 blocks x 8 int64 clock stamps
extern "C" void lemas_debug_attention_trace(void* buf) { g_att_trace = static_cast<long long*>(buf); }

extern "C" int lemas_attention_f16(const void* qk, int32_t ld_qk, const void* vt, int32_t vt_ld, const int32_t* kv_len,
                                   void* out16, int32_t batch, int32_t seq, int32_t heads, void* stream) {
  LEMAS_REQUIRE(qk && vt && out16, "lemas_attention_f16: null pointer");
  LEMAS_REQUIRE(ld_qk % 8 == 0 && vt_ld % 8 == 0 && vt_ld >= seq, "lemas_attention_f16: ld_qk/vt_ld must be multiples of 8");
  LEMAS_REQUIRE(batch >= 1 && seq >= 1 && heads >= 1, "lemas_attention_f16: bad shape");
  const int inner = heads * ATT_D;
  CUtensorMap tmQK, tmVT;
  {
    uint64_t dims[3] = {(uint64_t)2 * inner, (uint64_t)seq, (uint64_t)batch};
    uint64_t strides[2] = {(uint64_t)ld_qk * 2, (uint64_t)seq * ld_qk * 2};
    uint32_t box[3] = {64, ATT_BM, 1};
    LEMAS_TRY(make_tensor_map_f16(&tmQK, qk, 3, dims, strides, box));
  }
  {
    uint64_t dims[3] = {(uint64_t)seq, (uint64_t)ATT_D, (uint64_t)batch * heads};
    uint64_t strides[2] = {(uint64_t)vt_ld * 2, (uint64_t)ATT_D * vt_ld * 2};
    uint32_t box[3] = {64, ATT_D, 1};
    LEMAS_TRY(make_tensor_map_f16(&tmVT, vt, 3, dims, strides, box));
  }
  static unsigned long long configured = 0;
  LEMAS_CUDA_OK(ensure_dynamic_smem(attention_kernel, ATT_SMEM, configured));
  AttnParams p;
  p.trace = g_att_trace;
  p.kv_len = kv_len;
  p.out = static_cast<__half*>(out16);
  p.seq = seq;
  p.heads = heads;
  p.inner = inner;
  dim3 grid((seq + ATT_BM - 1) / ATT_BM, heads, batch);
  LEMAS_CUDA_OK(launch_pdl(attention_kernel, grid, dim3(ATT_THREADS), ATT_SMEM, (cudaStream_t)stream, tmQK, tmVT, p));
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}
