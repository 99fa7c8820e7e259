// Attention v5: ONE persistent CTA per SM that works on TWO 128-query tiles of one (batch, head) at a time and walks
// a static list of work items.  Same arithmetic as attention.cu (v3) — softmax(Q K^T / 8 + keymask) V for head_dim 64,
// modules.py:483-491 — restructured around what v3's traces showed:
//   * two co-resident v3 CTAs progressed at 2200 and 3250 clk per KV block (the older CTA's warps win the
//     arbitration), CTAs finished 18-33 us apart and the second wave of the 576-CTA grid left SMs idle for up to 18 us;
//     here both tiles belong to one CTA (equal-aged warps) and every SM runs the same number of tile pairs;
//   * K / V tiles are loaded once per KV block for 256 queries (one TMA ring shared by both tiles: 32 KB of
//     shared-memory writes per block pair instead of 64 KB);
//   * TMEM allocation, barrier set-up and the tensor-map prefetch happen once per CTA, the Q tiles of the next item
//     are prefetched while the current one runs, and the first S of the next item is issued behind the last P V of
//     the current one, so the per-tile overhead (~2.5 us in v3, ~20 % of a tile at 6 KV blocks) is mostly hidden;
//
// Roles (640 threads):
//   warp 0      TMA producer: Q tiles of an item into a 2-slot buffer, K / V^T blocks through two 4-slot rings
//   warp 1, 2   MMA issuer of query tile 0 / 1.  Per tile exactly v3's two independent 64-key half pipelines:
//                 S_x = Q K_x^T (SS, M128 N64 K16 x4) -> TMEM; P_x overwrites S_x in place (fp16 pairs, tcgen05.st by
//                 the softmax warps); O_x += P_x V_x (TS form, A from TMEM); S_x(j+1) issued right behind P_x(j) V_x(j).
//               A ring slot is released by tcgen05.commit from BOTH issuers (mbarrier count 2).
//   warp 3      idle (keeps warp index % 4 == TMEM sub-partition for the softmax warps)
//   warps 4-19  softmax: tile = (w-4)/8, key half = ((w-4)/4)%2, sub-partition = w%4; thread == query row.
// TMEM (512 columns): tile t at column 256 t: S/P half x at +64x (64 columns), O half x at +128+64x.
// Phase parities come from running counters (KV blocks / items processed so far); every wait is deadline-bounded.
#include <type_traits>

#include "att_common.cuh"

namespace lemas {

constexpr int A5_THREADS = 640;
constexpr int A5_STAGES = 4;
constexpr int A5_TILE_BYTES = 128 * 64 * 2;                    // 16 KB: Q tile, K block, V^T block
constexpr int A5_OFF_Q = 0;                                    // [2 slots][2 tiles]
constexpr int A5_OFF_KV = 4 * A5_TILE_BYTES;                   // [stages][K | V^T]
constexpr int A5_OFF_XCH = A5_OFF_KV + A5_STAGES * 2 * A5_TILE_BYTES;  // float2 [2 parities][2 tiles][2 halves][128]
constexpr int A5_OFF_BAR = A5_OFF_XCH + 2 * 2 * 2 * 128 * 8;
constexpr int A5_SMEM = A5_OFF_BAR + 512;                      // 200.5 KB: one CTA per SM

constexpr int B5_QF = 0, B5_QE = 2, B5_KF = 4, B5_KE = B5_KF + A5_STAGES, B5_VF = B5_KE + A5_STAGES,
              B5_VE = B5_VF + A5_STAGES, B5_SF = B5_VE + A5_STAGES, B5_PF = B5_SF + 4, B5_OF = B5_PF + 4,
              B5_OE = B5_OF + 4, B5_COUNT = B5_OE + 2;
static_assert(B5_COUNT * 8 + 8 <= 512, "barrier block");

constexpr float A5_RESCALE_LOG2 = 8.0f;

struct A5Item {
  int b, h, q0, kvl, n_blocks;
  bool live0, live1;
};
DEVI A5Item a5_item(const AttnParams& p, int it) {
  A5Item w;
  const int pair = it % p.n_pairs;
  const int hb = it / p.n_pairs;
  w.h = hb % p.heads;
  w.b = hb / p.heads;
  w.q0 = pair * 256;
  w.kvl = p.kv_len ? min(__ldg(p.kv_len + w.b), p.seq) : p.seq;
  w.n_blocks = (w.kvl + 127) / 128;
  // query tiles made only of padding rows are skipped: the to_out epilogue zeroes those rows (modules.py:499-501)
  w.live0 = w.q0 < w.kvl;
  w.live1 = w.q0 + 128 < w.kvl;
  return w;
}

// kCtrlLast: the four control warps take the HIGHEST warp indices (the softmax warps are warps 0-15) instead of the
// lowest — the issue arbiter of a sub-partition is not age-neutral, and the MMA issuer's few instructions are the
// latency-critical ones.
template <uint32_t kPolyMask, bool kCtrlLast>
__global__ void __launch_bounds__(A5_THREADS, 1)
attention5_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVT,
                  const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A5_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B5_COUNT);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ctrl = kCtrlLast ? warp - 16 : warp;   // 0 TMA, 1 / 2 MMA issuer of tile 0 / 1, 3 idle; < 0 or > 3: softmax
  const int sw = kCtrlLast ? warp : warp - 4;      // softmax warp index 0..15 (warp % 4 == sw % 4 either way)

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("lemas attention5: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmVT);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bars + B5_QF + s, 1);
      mbar_init(bars + B5_QE + s, 2);       // both MMA issuers
      mbar_init(bars + B5_OE + s, 8);       // the 8 softmax warps of a tile
    }
    for (int s = 0; s < A5_STAGES; ++s) {
      mbar_init(bars + B5_KF + s, 1);
      mbar_init(bars + B5_KE + s, 2);
      mbar_init(bars + B5_VF + s, 1);
      mbar_init(bars + B5_VE + s, 2);
    }
    for (int i = 0; i < 4; ++i) {           // index = tile * 2 + key half
      mbar_init(bars + B5_SF + i, 1);
      mbar_init(bars + B5_PF + i, 4);       // one arrival per softmax warp of the half
      mbar_init(bars + B5_OF + i, 1);
    }
    fence_barrier_init();
  }
  if (ctrl == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();  // set-up overlapped the previous kernel's tail; q / k / v are visible from here on

  const int n_items = p.n_items;
  const int stride = gridDim.x;

  if (ctrl == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int qn = 0, r = 0;
      for (int it = blockIdx.x; it < n_items; it += stride) {
        const A5Item w = a5_item(p, it);
        if (!w.live0) continue;
        const int slot = qn & 1;
        ATT_WAIT_P(bars + B5_QE + slot, ((qn >> 1) & 1) ^ 1, 1, it);
        mbar_arrive_expect_tx(bars + B5_QF + slot, w.live1 ? 2 * A5_TILE_BYTES : A5_TILE_BYTES);
        tma_load_3d(smem + A5_OFF_Q + (slot * 2) * A5_TILE_BYTES, &tmQK, bars + B5_QF + slot, w.h * 64, w.q0, w.b);
        if (w.live1)
          tma_load_3d(smem + A5_OFF_Q + (slot * 2 + 1) * A5_TILE_BYTES, &tmQK, bars + B5_QF + slot, w.h * 64,
                      w.q0 + 128, w.b);
        for (int j = 0; j < w.n_blocks; ++j, ++r) {
          const int s = r % A5_STAGES;
          const uint32_t ph = ((r / A5_STAGES) & 1) ^ 1;
          uint8_t* sk = smem + A5_OFF_KV + s * 2 * A5_TILE_BYTES;
          ATT_WAIT_P(bars + B5_KE + s, ph, 2, j);
          mbar_arrive_expect_tx(bars + B5_KF + s, A5_TILE_BYTES);
          tma_load_3d(sk, &tmQK, bars + B5_KF + s, p.inner + w.h * 64, j * 128, w.b);
          ATT_WAIT_P(bars + B5_VE + s, ph, 3, j);
          mbar_arrive_expect_tx(bars + B5_VF + s, A5_TILE_BYTES);
          tma_load_3d(sk + A5_TILE_BYTES, &tmVT, bars + B5_VF + s, j * 128, 0, w.b * p.heads + w.h);
          tma_load_3d(sk + A5_TILE_BYTES + A5_TILE_BYTES / 2, &tmVT, bars + B5_VF + s, j * 128 + 64, 0,
                      w.b * p.heads + w.h);
        }
        ++qn;
      }
    }
  } else if (ctrl == 1 || ctrl == 2) {
    // ------------------------------------------------------------------ MMA issuer of tile t
    const int t = ctrl - 1;
    constexpr uint32_t idesc = umma_idesc_f16(128, 64);   // both MMA shapes are M128 N64 K16
    const uint32_t tmem_s = tmem_base + t * 256;          // + 64 * half
    const uint32_t tmem_o = tmem_s + 128;                 // + 64 * half
    uint64_t* s_full = bars + B5_SF + t * 2;
    uint64_t* p_full = bars + B5_PF + t * 2;
    uint64_t* o_full = bars + B5_OF + t * 2;
    int qn = 0, r = 0, g = 0, on = 0;
    for (int it = blockIdx.x; it < n_items; it += stride) {
      const A5Item w = a5_item(p, it);
      if (!w.live0) continue;
      const int slot = qn & 1;
      const int nb = w.n_blocks;
      if (t == 1 && !w.live1) {
        // this tile is all padding: only keep the shared rings moving (pace on the full barriers so that an arrival
        // can never land in a later phase of the empty barriers)
        for (int j = 0; j < nb; ++j) {
          const int s = (r + j) % A5_STAGES;
          const uint32_t ph = ((r + j) / A5_STAGES) & 1;
          ATT_WAIT_P(bars + B5_KF + s, ph, 4, j);
          ATT_WAIT_P(bars + B5_VF + s, ph, 5, j);
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bars + B5_KE + s);
            mbar_arrive(bars + B5_VE + s);
          }
        }
        ATT_WAIT_P(bars + B5_QF + slot, (qn >> 1) & 1, 6, it);
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + B5_QE + slot);
        r += nb;
        ++qn;
        continue;
      }
      const uint32_t sq = smem_u32(smem + A5_OFF_Q + (slot * 2 + t) * A5_TILE_BYTES);
      auto issue_s = [&](int x, int rr) {  // S_x = Q_t K[64x : 64x+64]^T of ring entry rr
        const int s = rr % A5_STAGES;
        const uint32_t sk = smem_u32(smem + A5_OFF_KV + s * 2 * A5_TILE_BYTES) + x * (A5_TILE_BYTES / 2);
        const uint64_t adesc = umma_desc_sw128(sq), bdesc = umma_desc_sw128(sk);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_s + x * 64, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
        umma_commit(s_full + x);
        if (x == 1) umma_commit(bars + B5_KE + s);
      };
      ATT_WAIT_P(bars + B5_QF + slot, (qn >> 1) & 1, 6, it);
      ATT_WAIT_P(bars + B5_KF + r % A5_STAGES, (r / A5_STAGES) & 1, 4, 0);
      tc_fence_after();
      if (t == 1 && p.dephase_tile > 0 && nb > 2) {  // tile 1 starts a fraction of a block period behind tile 0
        const long long t_go = clock64() + p.dephase_tile;
        while (clock64() < t_go) { }
      }
      if (elect_one()) issue_s(0, r);
      __syncwarp();
      if (p.dephase_half > 0 && nb > 2) {  // head start for key half A (the two warps of a sub-partition and tile
        const long long t_go = clock64() + p.dephase_half;  // should not be in their exponential phase together)
        while (clock64() < t_go) { }
      }
      if (elect_one()) issue_s(1, r);
      __syncwarp();
      for (int j = 0; j < nb; ++j) {
        const int rr = r + j;
        const int s = rr % A5_STAGES;
        const bool last = j + 1 == nb;
        const uint32_t sv = smem_u32(smem + A5_OFF_KV + s * 2 * A5_TILE_BYTES + A5_TILE_BYTES);
        ATT_WAIT_P(bars + B5_VF + s, (rr / A5_STAGES) & 1, 5, j);
        if (!last) ATT_WAIT_P(bars + B5_KF + (rr + 1) % A5_STAGES, ((rr + 1) / A5_STAGES) & 1, 4, j + 1);
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          ATT_WAIT_P(p_full + x, (g + j) & 1, 7 + x, j);
#ifdef LEMAS_ATT_TRACE
          if (p.trace && lane == 0 && it == (int)blockIdx.x && j < 32 && (long long)blockIdx.x == p.trace[7])
            p.trace[4096 + ((t * 32 + j) * 2 + x) * 2] = clock64();
#endif
          // the first P V of an item overwrites O: the merge of the previous item must have read it
          if (j == 0 && x == 0) ATT_WAIT_P(bars + B5_OE + t, (on & 1) ^ 1, 9, it);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t bdesc = umma_desc_sw128(sv + x * (A5_TILE_BYTES / 2));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16_ts(tmem_o + x * 64, tmem_s + x * 64 + 8 * ks, bdesc + 2 * ks, idesc, (j | ks) != 0 ? 1u : 0u);
            if (x == 1) umma_commit(bars + B5_VE + s);
            if (last) {
              umma_commit(o_full + x);
              if (x == 1) umma_commit(bars + B5_QE + slot);
            } else {
              issue_s(x, rr + 1);  // overwrites P_x(j): executes behind the P V just issued
            }
          }
#ifdef LEMAS_ATT_TRACE
          if (p.trace && lane == 0 && it == (int)blockIdx.x && j < 32 && (long long)blockIdx.x == p.trace[7])
            p.trace[4096 + ((t * 32 + j) * 2 + x) * 2 + 1] = clock64();
#endif
          __syncwarp();
        }
      }
      r += nb;
      g += nb;
      ++on;
      ++qn;
    }
  } else if (sw >= 0 && sw < 16) {
    // ------------------------------------------------------------------ softmax warps
    const int t = sw >> 3;
    const int half = (sw >> 2) & 1;
    const int sub = warp & 3;          // TMEM sub-partition: lanes [32*sub, 32*sub+32)
    const int r = sub * 32 + lane;     // query row inside the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(sub * 32) << 16;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint32_t sb = smem_u32(smem);
    const uint32_t a_sfull = sb + A5_OFF_BAR + (B5_SF + t * 2 + half) * 8;
    const uint32_t a_pfull = sb + A5_OFF_BAR + (B5_PF + t * 2 + half) * 8;
    const uint32_t a_ofull = sb + A5_OFF_BAR + (B5_OF + t * 2) * 8;
    const uint32_t a_oempty = sb + A5_OFF_BAR + (B5_OE + t) * 8;
    const uint32_t tmem_s = tmem_base + t * 256;
    const uint32_t tmem_o = tmem_s + 128;
    const uint32_t t_s = tmem_s + lane_addr + half * 64;
    const uint32_t t_o = tmem_o + lane_addr + half * 64;
    int g = 0, on = 0;
    for (int it = blockIdx.x; it < n_items; it += stride) {
      const A5Item w = a5_item(p, it);
      if (!w.live0 || (t == 1 && !w.live1)) continue;
      const int q0 = w.q0 + t * 128;
      const int kvl = w.kvl;
      const int n_blocks = w.n_blocks;
      float m_ref = -INFINITY;           // max the accumulators O_half / l are currently scaled by
      float l_run = 0.f;
      // Warps whose 32 query rows all lie beyond the sequence keep the barrier protocol going but do no softmax
      // work: their P rows (left as whatever S held) only feed output rows that are never stored.
      const bool rows_dead = q0 + sub * 32 >= p.seq;
#ifdef LEMAS_ATT_TRACE  // clock64 stamps of the FIRST item of one CTA (tools/trace_att.py)
      const bool tr_item = p.trace != nullptr && lane == 0 && it == blockIdx.x && (long long)blockIdx.x == p.trace[7];
#endif
      for (int j = 0; j < n_blocks; ++j) {
        const uint32_t par = (g + j) & 1;
        if (rows_dead) {
          ATT_WAIT_A(a_sfull, par, 12 + half, j);
          __syncwarp();  // lanes poll independently: reconverge before the single arrival (see attention.cu)
          if (lane == 0) mbar_arrive_s(a_pfull);
          continue;
        }
        const int valid = min(max(kvl - j * 128 - half * 64, 0), 64);  // keys of this half-block that exist
#ifdef LEMAS_ATT_TRACE
        const bool tr = tr_item && j < 32;
        long long* tp = p.trace + (sw * 32 + j) * 8;
#define A5_STAMP(i) do { if (tr) tp[i] = clock64(); } while (0)
#else
#define A5_STAMP(i) do { } while (0)
#endif
        A5_STAMP(0);
        ATT_WAIT_A(a_sfull, par, 10 + half, j);   // S_x(j) landed; P_x(j-1) V_x(j-1) retired before it
        A5_STAMP(1);
        tc_fence_after();
        uint32_t s0[32], s1[32];
        tmem_ld_32x32(t_s, s0);
        tmem_ld_32x32(t_s + 32, s1);
        tmem_ld_wait();
        A5_STAMP(2);

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
        const bool grow = (mx - m_ref) * c > A5_RESCALE_LOG2;  // also true for the first finite max (m_ref = -inf)
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mx : m_ref;
          const float alpha = (m_ref == -INFINITY) ? 0.f : ex2f((m_ref - m_new) * c);
          l_run *= alpha;
          if (j > 0) {  // O_half holds the sum of blocks < j (retired, see the s_full wait): rescale it in TMEM
#pragma unroll 1
            for (int cc = 0; cc < 64; cc += 8) {
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
        A5_STAMP(3);

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
            if (kFull && ((kPolyMask >> i) & 1u)) {
              // exp2 on the FMA / ALU pipes (the SFU is the contended unit): x = n + f, n = round(x) via the
              // 1.5 * 2^23 magic constant, f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial (max relative error
              // 7.5e-5, below the fp16 rounding of P); 2^n added into the exponent field.  x <= 8 by the lazy-rescale
              // bound; the clamp keeps n inside the exponent range (result < 2^-125 ~ 0).
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
        A5_STAMP(4);
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
        A5_STAMP(6);
      }
      g += n_blocks;

      // ---- merge the two key halves of the tile, normalise, store; then hand O back to the MMA issuer
      ATT_WAIT_A(a_ofull, on & 1, 14, it);
      ATT_WAIT_A(a_ofull + 8, on & 1, 15, it);
      tc_fence_after();
      float2* xch = reinterpret_cast<float2*>(smem + A5_OFF_XCH) + ((on & 1) * 2 + t) * 256;
      xch[half * 128 + r] = make_float2(m_ref, l_run);
      named_bar_sync(1 + t * 4 + sub, 64);  // the two warps that share these 32 rows
      const float2 other = xch[(half ^ 1) * 128 + r];
      const float m_all = fmaxf(m_ref, other.x);
      const float w_me = (m_ref == -INFINITY) ? 0.f : ex2f((m_ref - m_all) * c);
      const float w_ot = (other.x == -INFINITY) ? 0.f : ex2f((other.x - m_all) * c);
      const float inv = 1.0f / (w_me * l_run + w_ot * other.y);
      const float wa = (half == 0 ? w_me : w_ot) * inv, wb = (half == 0 ? w_ot : w_me) * inv;
      uint32_t oa[32], ob[32];  // this warp outputs head-dim columns [32*half, 32*half+32)
      tmem_ld_32x32(tmem_o + lane_addr + half * 32, oa);
      tmem_ld_32x32(tmem_o + lane_addr + 64 + half * 32, ob);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_s(a_oempty);   // O is in registers: the next item's first P V may overwrite it
      const int row = q0 + r;
      if (row < p.seq) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((long)w.b * p.seq + row) * p.inner + w.h * 64 + half * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o[i] = __uint_as_float(oa[8 * u + i]) * wa + __uint_as_float(ob[8 * u + i]) * wb;
          uint4 v;
          v.x = pack_half2(o[0], o[1]);
          v.y = pack_half2(o[2], o[3]);
          v.z = pack_half2(o[4], o[5]);
          v.w = pack_half2(o[6], o[7]);
          dst[u] = v;
        }
      }
      ++on;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (ctrl == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace lemas

using namespace lemas;

static long long* g_att_trace = nullptr;
static int g_att_variant = -1;   // -1: default (LEMAS_ATT_VARIANT or built-in choice)
// debug aids (not part of the public header)
extern "C" void lemas_debug_attention_trace(void* buf) { g_att_trace = static_cast<long long*>(buf); }
extern "C" void lemas_debug_attention_variant(int v) { g_att_variant = v; }

namespace {
constexpr int kDefaultVariant = 0;   // v3 (attention.cu) is the fastest measured kernel (C2: 59 us; v5 63, v6 / v7 71)
int attention_variant() {
  if (g_att_variant >= 0) return g_att_variant;
  static int env = -2;
  if (env == -2) {
    const char* e = getenv("LEMAS_ATT_VARIANT");
    env = e ? atoi(e) : -1;
  }
  return env >= 0 ? env : kDefaultVariant;
}

template <uint32_t kPolyMask, bool kCtrlLast>
int launch_v5(const CUtensorMap& tmQK, const CUtensorMap& tmVT, const AttnParams& p, void* stream) {
  static unsigned long long configured = 0;
  LEMAS_CUDA_OK(ensure_dynamic_smem(attention5_kernel<kPolyMask, kCtrlLast>, A5_SMEM, configured));
  const int grid = p.n_items < sm_count() ? p.n_items : sm_count();
  LEMAS_CUDA_OK(launch_pdl(attention5_kernel<kPolyMask, kCtrlLast>, dim3(grid), dim3(A5_THREADS), A5_SMEM, (cudaStream_t)stream,
                           tmQK, tmVT, p));
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}
}  // namespace

// variants: 0 = v3 (attention.cu); v5: 1 = control warps first, 1/4 of the exp2 on the FMA pipe; 2 = control warps
// last, 1/4; 3 = control warps last, all exp2 on the SFU.  LEMAS_A5_DEPHASE_HALF / LEMAS_A5_DEPHASE_TILE (clocks)
// override the pipeline stagger (experiments).  v6 (attention6.cu): 4 = all exp2 on the SFU, 5 = 1/4 on the FMA pipe,
// 6 = 3/8.  v7 (attention7.cu, four key parts): 7 = all on the SFU, 8 = 1/4 on the FMA pipe, 9 = 3/8.
extern "C" int lemas_attention_f16(const void* qk, int32_t ld_qk, const void* vt, int32_t vt_ld, const int32_t* kv_len,
                                   void* out16, int32_t batch, int32_t seq, int32_t heads, void* stream) {
  LEMAS_REQUIRE(qk && vt && out16, "lemas_attention_f16: null pointer");
  LEMAS_REQUIRE(ld_qk % 8 == 0 && vt_ld % 8 == 0 && vt_ld >= seq, "lemas_attention_f16: ld_qk/vt_ld must be multiples of 8");
  LEMAS_REQUIRE(batch >= 1 && seq >= 1 && heads >= 1, "lemas_attention_f16: bad shape");
  const int variant = attention_variant();
  if (variant == 0)
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
  p.n_pairs = (seq + 255) / 256;
  p.n_items = p.n_pairs * heads * batch;
  static int dephase_half = -1, dephase_tile = -1;
  if (dephase_half < 0) {
    const char* e = getenv("LEMAS_A5_DEPHASE_HALF");
    dephase_half = e ? atoi(e) : (variant >= 4 ? 400 : 600);
    e = getenv("LEMAS_A5_DEPHASE_TILE");
    dephase_tile = e ? atoi(e) : 300;
  }
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("LEMAS_A7_DEBUG"); dbg = e ? atoi(e) : 0; }
  p.debug = dbg;
  p.dephase_half = dephase_half;
  p.dephase_tile = dephase_tile;
  if (variant >= 4) {  // v6 / v7 (attention6.cu / attention7.cu): one tile per item, double-buffered scores
    p.n_pairs = (seq + 127) / 128;
    p.n_items = p.n_pairs * heads * batch;
    if (variant >= 7) return attention_v7_launch(tmQK, tmVT, p, variant - 7, stream);
    return attention_v6_launch(tmQK, tmVT, p, variant - 4, stream);
  }
  switch (variant) {
    case 2: return launch_v5<0x11111111u, true>(tmQK, tmVT, p, stream);
    case 3: return launch_v5<0u, true>(tmQK, tmVT, p, stream);
    default: return launch_v5<0x11111111u, false>(tmQK, tmVT, p, stream);
  }
}
