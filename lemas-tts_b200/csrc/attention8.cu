// v8: the block pipeline of attention.cu (v3) inside a PERSISTENT CTA.
//
// v3 launches one CTA per 128-query tile; at C2 every CTA lives for 18 KV blocks and pays ~2 us of set-up and tear-down
// around ~25 us of work (barrier init, TMEM allocation, the first Q / K round trip, the drain before the merge, the
// exit / relaunch gap), and at C4 (6 blocks per tile) that share is ~25 % (stand-alone: 748 TFLOP/s at 18 blocks per
// CTA, 833 at 37 — same kernel, same steady state).  Here 2 CTAs per SM stay resident and walk a strided list of
// (batch, head, query tile) items: barriers and TMEM are set up once, the K / V rings run across item boundaries (the
// producer is already loading the next item's blocks while the last ones of the current item are in flight), Q is
// re-loaded the moment the last S of an item has retired, and the first S of the next item is issued right behind the
// last P V of the current one — the merge of item i overlaps the first block of item i + 1.
//
// Shared memory: Q 16 KB | K ring 2 x 16 KB | V^T ring 3 x 16 KB | merge exchange 2 KB | barriers = 98.25 KB (a K slot is
// free as soon as both halves of S_j have retired, so two suffice; the exchange area can no longer alias a ring slot).
// Everything else — the two independent 64-key half pipelines, P written back over the scores in TMEM, TS-form P V,
// lazy rescaling, a quarter of the exponentials on the FMA pipe — is v3's, see attention.cu.
#include <type_traits>

#include "att_common.cuh"

namespace lemas {

constexpr int A8_THREADS = 320;
constexpr int A8_BM = 128;
constexpr int A8_BN = 128;
constexpr int A8_D = 64;
constexpr int A8_KS = 2;   // K ring slots
constexpr int A8_VS = 3;   // V^T ring slots
constexpr int A8_TILE = A8_BN * A8_D * 2;                // 16 KB: Q, one K block, one V^T block
constexpr int A8_OFF_K = A8_TILE;
constexpr int A8_OFF_V = A8_OFF_K + A8_KS * A8_TILE;
constexpr int A8_OFF_XCH = A8_OFF_V + A8_VS * A8_TILE;   // float2 [2 halves][128 rows]
constexpr int A8_OFF_BAR = A8_OFF_XCH + 2 * A8_BM * 8;
constexpr int A8_SMEM = A8_OFF_BAR + 256;                // 98.25 KB

constexpr int B8_QF = 0, B8_QE = 1, B8_KF = 2, B8_KE = 4, B8_VF = 6, B8_VE = 9, B8_SF = 12, B8_PF = 14, B8_OF = 16,
              B8_COUNT = 18;
constexpr float A8_RESCALE_LOG2 = 8.0f;
#ifndef A8_POLY_EVERY
#define A8_POLY_EVERY 4
#endif
#ifndef A8_DEPHASE_CLK
#define A8_DEPHASE_CLK 1000
#endif

struct Item8 { int b, h, q0, kvl, n_blocks; };

// item -> (batch, head, query tile) in v3's launch order (query tile fastest): consecutive items share K / V in L2
DEVI bool item8(const AttnParams& p, int item, Item8& it) {
  const int qt = item % p.n_pairs;
  const int bh = item / p.n_pairs;
  it.h = bh % p.heads;
  it.b = bh / p.heads;
  it.q0 = qt * A8_BM;
  it.kvl = p.kv_len ? min(__ldg(p.kv_len + it.b), p.seq) : p.seq;
  it.n_blocks = (it.kvl + A8_BN - 1) / A8_BN;
  return it.q0 < it.kvl;   // tiles made only of padding rows are skipped by every role alike (see attention.cu)
}

__global__ void __launch_bounds__(A8_THREADS, 2)
attention8_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVT,
                  const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A8_OFF_BAR);
  uint64_t* q_full = bars + B8_QF;
  uint64_t* q_empty = bars + B8_QE;    // every S of the item has retired: the Q tile may be replaced
  uint64_t* k_full = bars + B8_KF;     // [2]
  uint64_t* k_empty = bars + B8_KE;    // [2]
  uint64_t* v_full = bars + B8_VF;     // [3]
  uint64_t* v_empty = bars + B8_VE;    // [3]
  uint64_t* s_full = bars + B8_SF;     // [2] per key half
  uint64_t* p_full = bars + B8_PF;     // [2] per key half
  uint64_t* o_full = bars + B8_OF;     // [2] per key half: the item's last P V has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B8_COUNT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("lemas attention: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmVT);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < A8_KS; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
    for (int s = 0; s < A8_VS; ++s) { mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1); }
    for (int x = 0; x < 2; ++x) {
      mbar_init(s_full + x, 1);
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
  pdl_wait();  // set-up overlapped the previous kernel's tail; q / k / v are visible from here on
  const uint32_t tmem_s = tmem_base;            // + 64 * half : S_x / P_x
  const uint32_t tmem_o = tmem_base + A8_BN;    // + 64 * half : O_x

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------- TMA producer
    if (elect_one()) {
      int kb = 0, vb = 0, n_it = 0;   // running K blocks, V blocks, items of this CTA
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        Item8 it;
        if (!item8(p, item, it)) continue;
        ATT_WAIT_P(q_empty, (n_it & 1) ^ 1, 0, n_it);
        mbar_arrive_expect_tx(q_full, A8_TILE);
        tma_load_3d(smem, &tmQK, q_full, it.h * A8_D, it.q0, it.b);
        for (int j = 0; j < it.n_blocks; ++j) {
          const int ks = kb % A8_KS, vs = vb % A8_VS;
          uint8_t* sk = smem + A8_OFF_K + ks * A8_TILE;
          uint8_t* sv = smem + A8_OFF_V + vs * A8_TILE;
          ATT_WAIT_P(k_empty + ks, ((kb / A8_KS) & 1) ^ 1, 1, j);
          mbar_arrive_expect_tx(k_full + ks, A8_TILE);
          tma_load_3d(sk, &tmQK, k_full + ks, p.inner + it.h * A8_D, j * A8_BN, it.b);
          ATT_WAIT_P(v_empty + vs, ((vb / A8_VS) & 1) ^ 1, 2, j);
          mbar_arrive_expect_tx(v_full + vs, A8_TILE);
          tma_load_3d(sv, &tmVT, v_full + vs, j * A8_BN, 0, it.b * p.heads + it.h);
          tma_load_3d(sv + A8_TILE / 2, &tmVT, v_full + vs, j * A8_BN + 64, 0, it.b * p.heads + it.h);
          ++kb;
          ++vb;
        }
        ++n_it;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(A8_BM, 64);   // both MMA shapes are M128 N64 K16
    const uint32_t sq = smem_u32(smem);
    int kb = 0, vb = 0, pb = 0, n_it = 0;   // running K blocks (S issued), V blocks (P V issued), p_full phases, items
    auto issue_s = [&](int x, bool last_s_of_item) {  // S_x = Q K[64x : 64x+64]^T for K block number kb (running)
      const int ks = kb % A8_KS;
      const uint32_t sk = smem_u32(smem + A8_OFF_K + ks * A8_TILE) + x * (A8_TILE / 2);
      const uint64_t adesc = umma_desc_sw128(sq), bdesc = umma_desc_sw128(sk);
#pragma unroll
      for (int k = 0; k < A8_D / 16; ++k) umma_f16_ss(tmem_s + x * 64, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
      umma_commit(s_full + x);
      if (x == 1) {
        umma_commit(k_empty + ks);
        if (last_s_of_item) umma_commit(q_empty);
      }
    };
    // first real item of this CTA
    int item = blockIdx.x;
    Item8 it;
    bool have = false;
    for (; item < p.n_items; item += gridDim.x)
      if (item8(p, item, it)) { have = true; break; }
    if (have) {
      ATT_WAIT_P(q_full, 0, 3, 0);
      ATT_WAIT_P(k_full + 0, 0, 4, 0);
      tc_fence_after();
      if (elect_one()) issue_s(0, false);
      __syncwarp();
      if (A8_DEPHASE_CLK > 0 && it.n_blocks > 2) {  // head start for key half A (see attention.cu)
        const long long t_go = clock64() + A8_DEPHASE_CLK;
        while (clock64() < t_go) { }
      }
      if (elect_one()) issue_s(1, it.n_blocks == 1);
      __syncwarp();
      ++kb;
    }
    while (have) {
      // next real item (its first S is issued from inside the last block of this one)
      Item8 nx = it;
      int nitem = item + gridDim.x;
      bool have_next = false;
      for (; nitem < p.n_items; nitem += gridDim.x)
        if (item8(p, nitem, nx)) { have_next = true; break; }
      for (int j = 0; j < it.n_blocks; ++j) {
        const bool last = j + 1 == it.n_blocks;
        const int vs = vb % A8_VS;
        const uint32_t sv = smem_u32(smem + A8_OFF_V + vs * A8_TILE);
        ATT_WAIT_P(v_full + vs, (vb / A8_VS) & 1, 5, j);
        if (!last) ATT_WAIT_P(k_full + (kb % A8_KS), (kb / A8_KS) & 1, 4, j + 1);
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          ATT_WAIT_P(p_full + x, pb & 1, 6 + x, j);
          tc_fence_after();
          if (elect_one()) {
            // O_x (+)= P_x(j) V_j[64x : 64x+64]; A = P from TMEM: 8 columns (16 fp16) per K16 step
            const uint64_t bdesc = umma_desc_sw128(sv + x * (A8_TILE / 2));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16_ts(tmem_o + x * A8_D, tmem_s + x * 64 + 8 * ks, bdesc + 2 * ks, idesc, (j | ks) != 0 ? 1u : 0u);
            if (x == 1) umma_commit(v_empty + vs);
            if (last) umma_commit(o_full + x);
            else issue_s(x, j + 2 == it.n_blocks);   // overwrites P_x(j): executes behind the P V just issued
          }
          __syncwarp();
          if (last && have_next) {
            // the FIRST scores of the next item go right behind the last P V of this one (its Q tile was loaded after
            // q_empty of this item, its first K block travels through the same ring)
            if (x == 0) {
              ATT_WAIT_P(q_full, (n_it + 1) & 1, 3, n_it + 1);
              ATT_WAIT_P(k_full + (kb % A8_KS), (kb / A8_KS) & 1, 4, 0);
              tc_fence_after();
            }
            if (elect_one()) issue_s(x, nx.n_blocks == 1);
            __syncwarp();
          }
        }
        ++vb;
        ++pb;
        if (!last || have_next) ++kb;
      }
      ++n_it;
      have = have_next;
      item = nitem;
      it = nx;
    }
  } else {
    // ------------------------------------------------------------------------------------------- softmax warps
    const int sub = warp & 3;          // TMEM sub-partition: lanes [32*sub, 32*sub+32)
    const int half = (warp - 2) >> 2;  // key half of every KV block this warp owns
    const int r = sub * 32 + lane;     // query row inside the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(sub * 32) << 16;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint32_t sb = smem_u32(smem);
    const uint32_t a_sfull = sb + A8_OFF_BAR + (B8_SF + half) * 8, a_pfull = sb + A8_OFF_BAR + (B8_PF + half) * 8;
    const uint32_t t_s = tmem_s + lane_addr + half * 64;
    const uint32_t t_o = tmem_o + lane_addr + half * A8_D;
    int sb_run = 0, n_it = 0;          // running blocks (s_full / p_full phases) and items of this CTA
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      Item8 it;
      if (!item8(p, item, it)) continue;
      const int q0 = it.q0, kvl = it.kvl, n_blocks = it.n_blocks, b = it.b, h = it.h;
      float m_ref = -INFINITY;         // max the accumulators O_half / l are currently scaled by
      float l_run = 0.f;
      const bool rows_dead = q0 + sub * 32 >= p.seq;   // see attention.cu
    for (int j = 0; j < n_blocks; ++j) {
      if (rows_dead) {
        ATT_WAIT_A(a_sfull, sb_run & 1, 12 + half, j);
        __syncwarp();
        if (lane == 0) mbar_arrive_s(a_pfull);
        ++sb_run;
        continue;
      }
      const int valid = min(max(kvl - j * A8_BN - half * 64, 0), 64);  // keys of this half-block that exist
      ATT_WAIT_A(a_sfull, sb_run & 1, 8 + half, j);   // S_x(j) landed; P_x(j-1) V_x(j-1) retired before it (same issuing thread)
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
      const bool grow = (mx - m_ref) * c > A8_RESCALE_LOG2;  // also true for the first finite max (m_ref = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mx : m_ref;
        const float alpha = (m_ref == -INFINITY) ? 0.f : ex2f((m_ref - m_new) * c);
        l_run *= alpha;
        if (j > 0) {  // O_half holds the sum of blocks < j (retired, see the s_full wait): rescale it in TMEM
#pragma unroll 1
          for (int cc = 0; cc < A8_D; cc += 8) {  // narrow chunks: S_j (64 registers) stays live across this
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
          if (kFull && A8_POLY_EVERY > 0 && (i % (A8_POLY_EVERY > 0 ? A8_POLY_EVERY : 1)) == 0) {
            // exp2 on the FMA / ALU pipes for one key pair in A8_POLY_EVERY (the SFU is the contended unit):
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
      ++sb_run;
    }

    // ---- merge the two key halves and normalise
    ATT_WAIT_A(sb + A8_OFF_BAR + B8_OF * 8, n_it & 1, 10, 0);
    ATT_WAIT_A(sb + A8_OFF_BAR + (B8_OF + 1) * 8, n_it & 1, 11, 0);
    tc_fence_after();
    float2* xch = reinterpret_cast<float2*>(smem + A8_OFF_XCH);
    xch[half * A8_BM + r] = make_float2(m_ref, l_run);
    named_bar_sync(1 + sub, 64);  // the two warps that share these 32 rows
    const float2 other = xch[(half ^ 1) * A8_BM + r];
    const float m_all = fmaxf(m_ref, other.x);
    const float w_me = (m_ref == -INFINITY) ? 0.f : ex2f((m_ref - m_all) * c);
    const float w_ot = (other.x == -INFINITY) ? 0.f : ex2f((other.x - m_all) * c);
    const float inv = 1.0f / (w_me * l_run + w_ot * other.y);
    const float wa = (half == 0 ? w_me : w_ot) * inv, wb = (half == 0 ? w_ot : w_me) * inv;
    uint32_t oa[32], ob[32];  // this warp outputs head-dim columns [32*half, 32*half+32)
    tmem_ld_32x32(tmem_o + lane_addr + half * 32, oa);
    tmem_ld_32x32(tmem_o + lane_addr + A8_D + half * 32, ob);
    tmem_ld_wait();
    // Both warps of the pair have read O_A, O_B and the exchange slots of these rows before either goes on: the next
    // item's first P (from this warp) releases the P V that OVERWRITES O_x, and its merge rewrites the exchange slots.
    tc_fence_before();
    named_bar_sync(1 + sub, 64);
    const int row = q0 + r;
    if (row < p.seq) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((long)b * p.seq + row) * p.inner + h * A8_D + half * 32);
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
      ++n_it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

int attention_v8_launch(const CUtensorMap& tmQK, const CUtensorMap& tmVT, const AttnParams& p, void* stream) {
  static unsigned long long configured = 0;
  LEMAS_CUDA_OK(ensure_dynamic_smem(attention8_kernel, A8_SMEM, configured));
  int grid = 2 * sm_count();
  if (grid > p.n_items) grid = p.n_items;
  LEMAS_CUDA_OK(launch_pdl(attention8_kernel, dim3(grid), dim3(A8_THREADS), A8_SMEM, (cudaStream_t)stream, tmQK, tmVT, p));
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

}  // namespace lemas
