// Non-causal self-attention for head_dim 64 on tcgen05 tensor cores (sm_100a), key-padding aware.
//
//   O = softmax(Q K^T / 8 + keymask) V        modules.py:483-491 (F.scaled_dot_product_attention call site)
//
// What bounds this kernel on B200 (measured, tools/trace_att.py + tools/micro/softmax_block_bench.cu):
//  * SFU: a 128 x 128 score block costs 512 tensor-core clocks (QK^T + PV) but 16 384 exponentials at 16 exp2 / clk /
//    SM = 1024 clocks (the packed f16x2 / bf16x2 forms are split into two MUFU ops and gain nothing);
//  * shared-memory bandwidth (128 B / clk / SM): every tcgen05.mma streams its operands from shared memory.  With P
//    staged through shared memory (generic stores + an SS-form P V) a block moved 144 KB = 1150 clk per CTA, more than
//    the SFU needs.  P therefore never touches shared memory here: the softmax warps write it back INTO the TMEM
//    columns its scores came from (fp16 pairs, 32 columns per 64 keys) and P V reads its A operand from TMEM
//    (tcgen05.mma ... [d], [a_tmem], b_desc) — 96 KB per block, no proxy fence, no P buffer to wait for.
//
// One CTA = 128 query rows of one (batch, head); two CTAs are co-resident per SM.  Roles (320 threads):
//   warp 0    TMA producer : Q tile once, then K tile [128 keys x 64] and V^T tile [64 x 128 keys] per KV block
//                            into two 3-slot rings (128B swizzle, mbarrier tx-count)
//   warp 1    MMA issuer   : every KV block is handled as two independent 64-key halves x = A, B, each with its own
//                            TMEM columns, barriers and online-softmax state (intra-CTA split-KV):
//                              S_x = Q K_x^T            (M128 N64 K16 x4, SS)   -> TMEM cols [64x, 64x+64)
//                              O_x += P_x V_x           (M128 N64 K16 x4, TS)   -> TMEM cols [128+64x, 192+64x)
//                            P_x(j) overwrites S_x(j) in place, so S_x(j+1) is issued right behind P_x(j) V_x(j) (the
//                            tensor core executes one thread's MMAs in order).  The two halves never wait for each
//                            other; half B is started half a block period after half A so that the two warps sharing
//                            an SM sub-partition are not in their exponential phase at the same time.
//   warps 2-9 softmax      : warps 2-5 own key half A, warps 6-9 key half B; thread == query row (tcgen05.ld 32x32b
//                            gives each lane one row).  Per block: wait S_x, read 64 scores, max, exp2 (packed
//                            FFMA2 / FADD2 arithmetic), fp16 pack, tcgen05.st, arrive.  The accumulators are rescaled
//                            lazily: O_x and l_x keep the scale of a reference max that is only advanced when a
//                            block's max exceeds it by more than 2^8 (P then stays <= 256); after the first blocks
//                            this almost never fires.  The halves are merged once at the end:
//                            O = (w_A O_A + w_B O_B) / (w_A l_A + w_B l_B),  w_X = 2^((m_X - max(m_A, m_B)) c).
// Fully masked KV blocks (keys >= kv_len) are skipped; the partial block is masked to zero probability.
// K comes from the fused QKV buffer [b*seq, ld_qk]; V is read from the transposed copy [b, head, 64, vt_ld] that
// the QKV GEMM epilogue writes, so both MMAs use K-major operands.
#include <type_traits>

#include "att_common.cuh"

namespace lemas {

constexpr int ATT_THREADS = 320;
constexpr int ATT_BM = 128;   // query rows per CTA
constexpr int ATT_BN = 128;   // keys per KV block
constexpr int ATT_D = 64;
constexpr int ATT_STAGES = 3;

constexpr int ATT_Q_BYTES = ATT_BM * ATT_D * 2;          // 16 KB
constexpr int ATT_K_BYTES = ATT_BN * ATT_D * 2;          // 16 KB (two 8 KB halves of 64 keys)
constexpr int ATT_V_BYTES = ATT_D * ATT_BN * 2;          // 16 KB (two 8 KB halves of 64 keys)
constexpr int ATT_KV_STAGE = ATT_K_BYTES + ATT_V_BYTES;
constexpr int ATT_OFF_KV = ATT_Q_BYTES;
constexpr int ATT_OFF_XCH = ATT_OFF_KV;                  // float2 [2 halves][128 rows] (reference max, row sum):
                                                         // reuses ring slot 0 after the last MMA has retired
constexpr int ATT_OFF_BAR = ATT_OFF_KV + ATT_STAGES * ATT_KV_STAGE;
constexpr int ATT_SMEM = ATT_OFF_BAR + 256;              // 112.25 KB: two CTAs per SM

// barrier slots (8 bytes each) from ATT_OFF_BAR
constexpr int BAR_Q = 0, BAR_KF = 1, BAR_KE = 4, BAR_VF = 7, BAR_VE = 10, BAR_SF = 13, BAR_PF = 15, BAR_OF = 17,
              BAR_COUNT = 19;

constexpr float ATT_RESCALE_LOG2 = 8.0f;  // advance the reference max only past 2^8 growth
#ifndef ATT_POLY_EVERY
#define ATT_POLY_EVERY 4                  // one key pair in 4 gets its exp2 from the FMA pipe (0 = all on the SFU)
#endif
#ifndef ATT_DEPHASE_CLK
#define ATT_DEPHASE_CLK 1000              // head start of key half A over key half B (about half a block period)
#endif

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVT,
                 const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_OFF_BAR);
  uint64_t* q_full = bars + BAR_Q;
  uint64_t* k_full = bars + BAR_KF;    // [3]  K and V^T travel through separate rings: a K slot is free as soon as
  uint64_t* k_empty = bars + BAR_KE;   // [3]  both halves of S_j have retired, a V slot only after both halves of
  uint64_t* v_full = bars + BAR_VF;    // [3]  P V_j
  uint64_t* v_empty = bars + BAR_VE;   // [3]
  uint64_t* s_full = bars + BAR_SF;    // [2] per key half: S_x(j) is in TMEM (and P_x(j-1) V_x(j-1) has retired)
  uint64_t* p_full = bars + BAR_PF;    // [2] per key half: P_x(j) is in TMEM
  uint64_t* o_full = bars + BAR_OF;    // [2] per key half: the last P V has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);

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
    for (int x = 0; x < 2; ++x) {
      mbar_init(s_full + x, 1);
      mbar_init(p_full + x, 4);         // one arrival per softmax warp of the half
      mbar_init(o_full + x, 1);
    }
    fence_barrier_init();
  }
#ifdef ATT_ALLOC_AFTER_PDL
  __syncthreads();
  pdl_trigger();
  pdl_wait();
#endif
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#ifndef ATT_ALLOC_AFTER_PDL
  pdl_trigger();
  pdl_wait();  // prologue overlapped the previous kernel's tail; q/k/v are visible from here on
#endif
  const uint32_t tmem_s = tmem_base;            // + 64 * half : S_x (64 fp32 columns) / P_x (32 columns of fp16 pairs)
  const uint32_t tmem_o = tmem_base + ATT_BN;   // + 64 * half : O_x

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, ATT_Q_BYTES);
      tma_load_3d(smem, &tmQK, q_full, h * ATT_D, q0, b);
      for (int j = 0; j < n_blocks; ++j) {
        const int s = j % ATT_STAGES;
        const uint32_t ph = ((j / ATT_STAGES) & 1) ^ 1;
        uint8_t* sk = smem + ATT_OFF_KV + s * ATT_KV_STAGE;
        ATT_WAIT_P(k_empty + s, ph, 1, j);
        mbar_arrive_expect_tx(k_full + s, ATT_K_BYTES);
        tma_load_3d(sk, &tmQK, k_full + s, p.inner + h * ATT_D, j * ATT_BN, b);
        ATT_WAIT_P(v_empty + s, ph, 2, j);
        mbar_arrive_expect_tx(v_full + s, ATT_V_BYTES);
        tma_load_3d(sk + ATT_K_BYTES, &tmVT, v_full + s, j * ATT_BN, 0, b * p.heads + h);
        tma_load_3d(sk + ATT_K_BYTES + ATT_V_BYTES / 2, &tmVT, v_full + s, j * ATT_BN + 64, 0, b * p.heads + h);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16(ATT_BM, 64);   // both MMA shapes are M128 N64 K16
    const uint32_t sq = smem_u32(smem);
    auto issue_s = [&](int x, int j) {  // S_x(j) = Q K_j[64x : 64x+64]^T
      const uint32_t sk = smem_u32(smem + ATT_OFF_KV + (j % ATT_STAGES) * ATT_KV_STAGE) + x * (ATT_K_BYTES / 2);
      const uint64_t adesc = umma_desc_sw128(sq), bdesc = umma_desc_sw128(sk);
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k) umma_f16_ss(tmem_s + x * 64, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
      umma_commit(s_full + x);
      if (x == 1) umma_commit(k_empty + (j % ATT_STAGES));
    };
    ATT_WAIT_P(q_full, 0, 3, 0);
    ATT_WAIT_P(k_full + 0, 0, 4, 0);
    tc_fence_after();
    if (elect_one()) issue_s(0, 0);
    __syncwarp();
    if (ATT_DEPHASE_CLK > 0 && n_blocks > 2) {  // head start for key half A (see the header)
      const long long t_go = clock64() + ATT_DEPHASE_CLK;
      while (clock64() < t_go) { }
    }
    if (elect_one()) issue_s(1, 0);
    __syncwarp();
    for (int j = 0; j < n_blocks; ++j) {
      const bool last = j + 1 == n_blocks;
      const uint32_t sv = smem_u32(smem + ATT_OFF_KV + (j % ATT_STAGES) * ATT_KV_STAGE + ATT_K_BYTES);
      ATT_WAIT_P(v_full + (j % ATT_STAGES), (j / ATT_STAGES) & 1, 5, j);
      if (!last) ATT_WAIT_P(k_full + ((j + 1) % ATT_STAGES), ((j + 1) / ATT_STAGES) & 1, 4, j + 1);
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        ATT_WAIT_P(p_full + x, j & 1, 6 + x, j);
        tc_fence_after();
        if (elect_one()) {
          // O_x (+)= P_x(j) V_j[64x : 64x+64]; A = P from TMEM: 8 columns (16 fp16) per K16 step
          const uint64_t bdesc = umma_desc_sw128(sv + x * (ATT_V_BYTES / 2));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16_ts(tmem_o + x * ATT_D, tmem_s + x * 64 + 8 * ks, bdesc + 2 * ks, idesc, (j | ks) != 0 ? 1u : 0u);
          if (x == 1) umma_commit(v_empty + (j % ATT_STAGES));
          if (last) umma_commit(o_full + x);
          else issue_s(x, j + 1);       // overwrites P_x(j): executes behind the P V just issued
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
    // 32-bit shared-window addresses (one register each): the softmax loop is latency-bound, every instruction of
    // address arithmetic or pointer conversion in it costs throughput
    const uint32_t sb = smem_u32(smem);
    const uint32_t a_sfull = sb + ATT_OFF_BAR + (BAR_SF + half) * 8, a_pfull = sb + ATT_OFF_BAR + (BAR_PF + half) * 8;
    const uint32_t t_s = tmem_s + lane_addr + half * 64;
    const uint32_t t_o = tmem_o + lane_addr + half * ATT_D;

    // Warps whose 32 query rows all lie beyond the sequence (last query tile) keep the barrier protocol going but do
    // no softmax work: their P rows (left as whatever S held) only feed output rows that are never stored.
    const bool rows_dead = q0 + sub * 32 >= p.seq;
#ifdef LEMAS_ATT_TRACE
    if (cta_rec && threadIdx.x == 64) cta_rec[2] = gtime();
#endif
    for (int j = 0; j < n_blocks; ++j) {
      if (rows_dead) {
        ATT_WAIT_A(a_sfull, j & 1, 12 + half, j);
        // every lane polls on its own: without this reconvergence lane 0 (the only one that arrives) can run ahead,
        // the barrier laps the other 31 lanes by two phases and their parity wait never succeeds
        __syncwarp();
        if (lane == 0) mbar_arrive_s(a_pfull);
        continue;
      }
      const int valid = min(max(kvl - j * ATT_BN - half * 64, 0), 64);  // keys of this half-block that exist
#ifdef LEMAS_ATT_TRACE  // clock64 stamps of one CTA (tools/trace_att.py); costs a few % of the kernel, off by default
      // traced CTA = linear id stored in the unused stamp slot [warp 0][block 0][7]
      const bool tr = p.trace != nullptr && lane == 0 && j < 32 &&
                      (long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x == p.trace[7];
      long long* tp = p.trace + ((warp - 2) * 32 + j) * 8;
#define ATT_STAMP(i) do { if (tr) tp[i] = clock64(); } while (0)
#else
#define ATT_STAMP(i) do { } while (0)
#endif
      ATT_STAMP(0);
      ATT_WAIT_A(a_sfull, j & 1, 8 + half, j);   // S_x(j) landed; P_x(j-1) V_x(j-1) retired before it (same issuing thread)
      ATT_STAMP(1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(t_s, s0);
      tmem_ld_32x32(t_s + 32, s1);
      tmem_ld_wait();
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
        if (j > 0) {  // O_half holds the sum of blocks < j (retired, see the s_full wait): rescale it in TMEM
#pragma unroll 1
          for (int cc = 0; cc < ATT_D; cc += 8) {  // narrow chunks: S_j (64 registers) stays live across this
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
      ATT_STAMP(3);

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
          if (kFull && ATT_POLY_EVERY > 0 && (i % (ATT_POLY_EVERY > 0 ? ATT_POLY_EVERY : 1)) == 0) {
            // exp2 on the FMA / ALU pipes for one key pair in ATT_POLY_EVERY (the SFU is the contended unit):
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
      ATT_STAMP(4);
      // P_x(j) -> TMEM, over the first 32 of the 64 columns S_x(j) was read from: column k holds keys (2k, 2k+1)
      tmem_st_32x32(t_s, pk);
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
      ATT_STAMP(6);
    }

    // ---- merge the two key halves and normalise
#ifdef LEMAS_ATT_TRACE
    if (cta_rec && threadIdx.x == 64) cta_rec[3] = gtime();
#endif
    ATT_WAIT_A(sb + ATT_OFF_BAR + BAR_OF * 8, 0, 10, 0);
    ATT_WAIT_A(sb + ATT_OFF_BAR + (BAR_OF + 1) * 8, 0, 11, 0);
    tc_fence_after();
    float2* xch = reinterpret_cast<float2*>(smem + ATT_OFF_XCH);  // ring slot 0: free now that every MMA has retired
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

// v3 launcher (one 128-query tile per CTA, two CTAs per SM); lemas_attention_f16 below dispatches.
int lemas::attention_v3_launch(const void* qk, int32_t ld_qk, const void* vt, int32_t vt_ld, const int32_t* kv_len,
                               void* out16, int32_t batch, int32_t seq, int32_t heads, long long* trace, void* stream) {
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
  AttnParams p = {};
  p.trace = trace;
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

static long long* g_att_trace = nullptr;
static int g_att_variant = -1;   // -1: default (LEMAS_ATT_VARIANT or built-in choice)
// debug aids (not part of the public header)
extern "C" void lemas_debug_attention_trace(void* buf) { g_att_trace = static_cast<long long*>(buf); }
extern "C" void lemas_debug_attention_variant(int v) { g_att_variant = v; }

namespace {
constexpr int kAutoVariant = 100;    // choose between v3 and v9 by shape (see lemas_attention_f16)
constexpr int kDefaultVariant = kAutoVariant;   // v3 is the fastest kernel on long sequences (C2: 59 us; v5 63, v6 / v7 71,
                                                // v8 63 us), v9 on short and ragged ones
int attention_variant() {
  if (g_att_variant >= 0) return g_att_variant;
  static int env = -2;
  if (env == -2) {
    const char* e = getenv("LEMAS_ATT_VARIANT");
    env = e ? atoi(e) : -1;
  }
  return env >= 0 ? env : kDefaultVariant;
}

}  // namespace

// LEMAS_ATT_VARIANT / lemas_debug_attention_variant: 0 = v3 (this file, production); 7 / 8 / 9 = v7 (attention7.cu,
// experimental: persistent CTA, double-buffered scores, four key parts) with 0, 1/4, 3/8 of the exp2 on the FMA pipe;
// 18-21 = v7 timing ablations (wrong results); 30 = v8 (attention8.cu, experimental: v3's pipeline in a persistent CTA).
// LEMAS_A7_DEPHASE (clocks) sets v7's start-up stagger.
extern "C" int lemas_attention_f16(const void* qk, int32_t ld_qk, const void* vt, int32_t vt_ld, const int32_t* kv_len,
                                   void* out16, int32_t batch, int32_t seq, int32_t heads, void* stream) {
  LEMAS_REQUIRE(qk && vt && out16, "lemas_attention_f16: null pointer");
  LEMAS_REQUIRE(ld_qk % 8 == 0 && vt_ld % 8 == 0 && vt_ld >= seq, "lemas_attention_f16: ld_qk/vt_ld must be multiples of 8");
  LEMAS_REQUIRE(batch >= 1 && seq >= 1 && heads >= 1, "lemas_attention_f16: bad shape");
  int variant = attention_variant();
  // Default choice by shape (LEMAS_ATT_VARIANT = 0 forces v3, 40 forces v9): BATCHES of short or ragged sequences go to v9
  // (attention9.cu: one 64-key pipeline per CTA, four CTAs per SM — twice the blocks per CTA, no merge: C4's 768 keys
  // 229 vs 238 us, the ragged C3 mix 227 vs 246 us); long uniform sequences stay on v3 (C2 59.5 vs 60.2 us, C5 88 vs 100:
  // v9's 592 CTA slots quantise 704 tiles into two waves).  profiles/r02ah_attention_v9.log
  // v9 needs enough tiles to fill its 4 x 148 CTA slots more than once: with few tiles (one short utterance: 2 x 1024 keys,
  // 256 tiles) v3's two pipelines per tile finish in 16 us where v9's single one needs 28.
  if (variant == kAutoVariant) {
    const long tiles = (long)((seq + 127) / 128) * heads * batch;
    variant = ((kv_len != nullptr || seq <= 1024) && tiles >= 5L * sm_count()) ? 40 : 0;   // crossover measured at ~750 tiles
  }
  if (variant < 7)
    return attention_v3_launch(qk, ld_qk, vt, vt_ld, kv_len, out16, batch, seq, heads, g_att_trace, stream);
  const int inner = heads * 64;
  CUtensorMap tmQK, tmVT;
  {
    uint64_t dims[3] = {(uint64_t)2 * inner, (uint64_t)seq, (uint64_t)batch};
    uint64_t strides[2] = {(uint64_t)ld_qk * 2, (uint64_t)seq * ld_qk * 2};
    uint32_t box[3] = {64, 128, 1};
    LEMAS_TRY(make_tensor_map_f16(&tmQK, qk, 3, dims, strides, box));
  }
  {
    uint64_t dims[3] = {(uint64_t)seq, 64, (uint64_t)batch * heads};
    uint64_t strides[2] = {(uint64_t)vt_ld * 2, (uint64_t)64 * vt_ld * 2};
    uint32_t box[3] = {64, 64, 1};
    LEMAS_TRY(make_tensor_map_f16(&tmVT, vt, 3, dims, strides, box));
  }
  AttnParams p = {};
  p.trace = g_att_trace;
  p.kv_len = kv_len;
  p.out = static_cast<__half*>(out16);
  p.seq = seq;
  p.heads = heads;
  p.inner = inner;
  static int dephase_half = -1, dephase_tile = -1;
  if (dephase_half < 0) {
    const char* e = getenv("LEMAS_A7_DEPHASE");
    dephase_half = e ? atoi(e) : 150;
    dephase_tile = 0;
  }
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("LEMAS_A7_DEBUG"); dbg = e ? atoi(e) : 0; }
  p.debug = dbg;
  p.dephase_half = dephase_half;
  p.dephase_tile = dephase_tile;
  p.n_pairs = (seq + 127) / 128;   // 128-query tiles per (batch, head)
  p.n_items = p.n_pairs * heads * batch;
  if (variant == 40) return attention_v9_launch(qk, ld_qk, vt, vt_ld, p, batch, stream);   // v9: one 64-key pipeline per CTA
  if (variant == 30) return attention_v8_launch(tmQK, tmVT, p, stream);   // v8: v3's pipeline in a persistent CTA
  return attention_v7_launch(tmQK, tmVT, p, variant - 7, stream);
}
