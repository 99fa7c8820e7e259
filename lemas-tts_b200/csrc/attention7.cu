// Attention v7: v6's double-buffered scores with FOUR 32-key parts per KV block instead of two 64-key halves — 16
// softmax warps (four per SM sub-partition) that never wait for the tensor pipe.  Arithmetic as attention.cu (v3):
// softmax(Q K^T / 8 + keymask) V, head_dim 64 (modules.py:483-491); every key part keeps its own online-softmax state
// (reference max, row sum) and its own accumulator O_q; the four partial results are merged once per query tile.
//
// Measured on the way here (profiles/r02*_): the softmax arithmetic of a 128 x 128 score block needs ~900 clk of a
// sub-partition's time (tools/micro/softmax_block_bench.cu: 450 clk per 64-key warp block with two or four warps per
// sub-partition, SFU 85 % busy), v3 / v5 run at 1350 / 1650 clk because every warp idles 600-1000 clk per block for
// "P stored -> P V -> next S -> scores visible" (a thread is held ~140 clk per dependent tcgen05.mma), and v6 (scores
// issued ahead, 8 softmax warps) removed that wait but cannot hide a warp's own latencies with two warps per
// sub-partition (1550 clk).  v7 = no wait AND four instruction streams per sub-partition.
//
// TMEM (all 512 columns): score buffer b at column 128 b, key part q at +32 q (32 fp32 columns; P_q as fp16 pairs over
// its first 16); O_q at 256 + 64 q.
// Roles (704 threads):
//   warp 0      TMA producer : Q tile of an item into a 2-slot buffer, K / V^T blocks through a 5-slot ring
//   warp 1      MMA issuer of S(g) = Q K(g)^T (M128 N128 K16 x4) into buffer g & 1, up to two blocks ahead
//   warps 2-5   MMA issuer of O_q += P_q V_q (M128 N64 K16 x2), q = warp - 2
//   warps 6-21  softmax: key part q = (warp - 6) / 4, TMEM sub-partition warp % 4; thread == query row
#include <type_traits>

#include "att_common.cuh"

namespace lemas {

constexpr int A7_THREADS = 704;
constexpr int A7_STAGES = 5;
constexpr int A7_TILE_BYTES = 128 * 64 * 2;                    // 16 KB: Q tile, K block, V^T block
constexpr int A7_OFF_Q = 0;                                    // [2 slots]
constexpr int A7_OFF_KV = 2 * A7_TILE_BYTES;                   // [stages][K | V^T]
constexpr int A7_OFF_XCH = A7_OFF_KV + A7_STAGES * 2 * A7_TILE_BYTES;  // float2 [2 parities][4 parts][128]
constexpr int A7_OFF_BAR = A7_OFF_XCH + 2 * 4 * 128 * 8;
constexpr int A7_SMEM = A7_OFF_BAR + 512;                      // 200.5 KB: one CTA per SM

// barriers: q_full[2] q_empty[2] k_full[S] k_empty[S] v_full[S] v_empty[S] s_full[2] p_full[4 parts][2 buffers]
// pv_done[4][2] o_full[4] o_empty
constexpr int B7_QF = 0, B7_QE = 2, B7_KF = 4, B7_KE = B7_KF + A7_STAGES, B7_VF = B7_KE + A7_STAGES,
              B7_VE = B7_VF + A7_STAGES, B7_SF = B7_VE + A7_STAGES, B7_PF = B7_SF + 2, B7_PD = B7_PF + 8,
              B7_OF = B7_PD + 8, B7_OE = B7_OF + 4, B7_COUNT = B7_OE + 1;
static_assert(B7_COUNT * 8 + 8 <= 512, "barrier block");
static_assert(A7_SMEM <= 227 * 1024, "shared memory budget");

constexpr float A7_RESCALE_LOG2 = 8.0f;
#ifndef A7_SINGLE_LANE_WAITS
#define A7_SINGLE_LANE_WAITS 0
#endif

struct A7Item {
  int b, h, q0, kvl, n_blocks;
  bool live;
};
DEVI A7Item a7_item(const AttnParams& p, int it) {  // n_pairs holds the number of 128-query tiles per (batch, head)
  A7Item w;
  const int tile = it % p.n_pairs;
  const int hb = it / p.n_pairs;
  w.h = hb % p.heads;
  w.b = hb / p.heads;
  w.q0 = tile * 128;
  w.kvl = p.kv_len ? min(__ldg(p.kv_len + w.b), p.seq) : p.seq;
  w.n_blocks = (w.kvl + 127) / 128;
  // query tiles made only of padding rows are skipped: the to_out epilogue zeroes those rows (modules.py:499-501)
  w.live = w.q0 < w.kvl;
  return w;
}

// Walks the live work items of this CTA block by block.
struct A7Cursor {
  int it, j, qn, r;       // item index, block inside the item, live items / KV blocks consumed before this one
  A7Item w;
  bool done;
  DEVI void seek(const AttnParams& p, int stride) {
    while (it < p.n_items) {
      w = a7_item(p, it);
      if (w.live) { done = false; return; }
      it += stride;
    }
    done = true;
  }
  DEVI void start(const AttnParams& p, int first, int stride) { it = first; j = 0; qn = 0; r = 0; seek(p, stride); }
  DEVI void next_block(const AttnParams& p, int stride) {
    ++r;
    if (++j == w.n_blocks) { j = 0; ++qn; it += stride; seek(p, stride); }
  }
};

// kAblate (timing experiments only, results are wrong): 1 = no max pass, 2 = no exponentials, 3 = neither,
// 4 = neither and no TMEM load / store of the scores
template <uint32_t kPolyMask, int kAblate = 0>   // bit i: key pair i of a 32-key part (16 pairs) gets its exp2 from the FMA pipe
__global__ void __launch_bounds__(A7_THREADS, 1)
attention7_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVT,
                  const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A7_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B7_COUNT);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("lemas attention7: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmVT);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bars + B7_QF + s, 1);
      mbar_init(bars + B7_QE + s, 1);
      mbar_init(bars + B7_SF + s, 1);       // per score buffer
    }
    for (int i = 0; i < 8; ++i) {           // index = key part * 2 + score buffer
      // P ready: four warps per part.  Per BUFFER, so that a fast warp's arrival for block g+1 cannot land in the
      // phase a slow warp still owes its arrival for g (it can only come back to this barrier for g+2, whose scores
      // are issued after P V(g), i.e. after all four arrivals for g were observed).
      mbar_init(bars + B7_PF + i, 4);
      mbar_init(bars + B7_PD + i, 1);       // P V retired: frees the buffer part for S(g+2), lets the lazy rescale touch O
    }
    for (int q = 0; q < 4; ++q) mbar_init(bars + B7_OF + q, 1);   // last P V of an item retired
    mbar_init(bars + B7_OE, 16);            // the 16 softmax warps have read O
    for (int s = 0; s < A7_STAGES; ++s) {
      mbar_init(bars + B7_KF + s, 1);
      mbar_init(bars + B7_KE + s, 1);
      mbar_init(bars + B7_VF + s, 1);
      mbar_init(bars + B7_VE + s, 4);       // the four P V issuers
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();  // set-up overlapped the previous kernel's tail; q / k / v are visible from here on

  const int stride = gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int qn = 0, r = 0;
      for (int it = blockIdx.x; it < p.n_items; it += stride) {
        const A7Item w = a7_item(p, it);
        if (!w.live) continue;
        const int slot = qn & 1;
        ATT_WAIT_P(bars + B7_QE + slot, ((qn >> 1) & 1) ^ 1, 1, it);
        mbar_arrive_expect_tx(bars + B7_QF + slot, A7_TILE_BYTES);
        tma_load_3d(smem + A7_OFF_Q + slot * A7_TILE_BYTES, &tmQK, bars + B7_QF + slot, w.h * 64, w.q0, w.b);
        for (int j = 0; j < w.n_blocks; ++j, ++r) {
          const int s = r % A7_STAGES;
          const uint32_t ph = ((r / A7_STAGES) & 1) ^ 1;
          uint8_t* sk = smem + A7_OFF_KV + s * 2 * A7_TILE_BYTES;
          ATT_WAIT_P(bars + B7_KE + s, ph, 2, j);
          mbar_arrive_expect_tx(bars + B7_KF + s, A7_TILE_BYTES);
          tma_load_3d(sk, &tmQK, bars + B7_KF + s, p.inner + w.h * 64, j * 128, w.b);
          ATT_WAIT_P(bars + B7_VE + s, ph, 3, j);
          mbar_arrive_expect_tx(bars + B7_VF + s, A7_TILE_BYTES);
          tma_load_3d(sk + A7_TILE_BYTES, &tmVT, bars + B7_VF + s, j * 128, 0, w.b * p.heads + w.h);
          tma_load_3d(sk + A7_TILE_BYTES + A7_TILE_BYTES / 2, &tmVT, bars + B7_VF + s, j * 128 + 64, 0,
                      w.b * p.heads + w.h);
        }
        ++qn;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer of the scores
    // The buffer g & 1 held P(g-2), which the four P V issuers consume — other threads, so their per-buffer pv_done
    // commits are awaited (the tensor pipe orders only the MMAs of one thread).
    // A7_SINGLE_LANE_WAITS = 1: ONE lane runs the whole role and one lane of a softmax warp polls for the warp.
    // Measured SLOWER (C2: 91 us against 71 us with all 32 lanes polling) — kept as a switch for the record.
    constexpr uint32_t idesc_s = umma_idesc_f16(128, 128);
    A7Cursor cs;
    cs.start(p, blockIdx.x, stride);
    if (A7_SINGLE_LANE_WAITS && lane != 0) cs.done = true;
#ifdef LEMAS_ATT_TRACE
    const bool str = p.trace != nullptr && lane == 0 && (long long)blockIdx.x == p.trace[7];
#define A7_SSTAMP(i) do { if (str && cs.it == (int)blockIdx.x && cs.j < 32) p.trace[4096 + cs.j * 4 + (i)] = clock64(); } while (0)
#else
#define A7_SSTAMP(i) do { } while (0)
#endif
    while (!cs.done) {
      const int g = cs.r, buf = g & 1;
      A7_SSTAMP(0);
      if (cs.j == 0) ATT_WAIT_P(bars + B7_QF + (cs.qn & 1), (cs.qn >> 1) & 1, 4, cs.it);
      const int s = g % A7_STAGES;
      ATT_WAIT_P(bars + B7_KF + s, (g / A7_STAGES) & 1, 5, cs.j);
      A7_SSTAMP(1);
      if (g >= 2 && !(p.debug & 1)) {  // completion (g-2)/2 of the buffer's pv_done barriers; the next one needs S(g): no lapping
#pragma unroll
        for (int q = 0; q < 4; ++q) ATT_WAIT_P(bars + B7_PD + 2 * q + buf, ((g - 2) >> 1) & 1, 6, cs.j);
      }
      A7_SSTAMP(2);
      tc_fence_after();
      if (A7_SINGLE_LANE_WAITS || elect_one()) {
        const uint64_t adesc = umma_desc_sw128(smem_u32(smem + A7_OFF_Q + (cs.qn & 1) * A7_TILE_BYTES));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + A7_OFF_KV + s * 2 * A7_TILE_BYTES));
        const uint32_t d = tmem_base + buf * 128;
        if (!(p.debug & 4)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        }
        umma_commit(bars + B7_SF + buf);
        umma_commit(bars + B7_KE + s);
        if (cs.j + 1 == cs.w.n_blocks) umma_commit(bars + B7_QE + (cs.qn & 1));  // last read of this Q tile
      }
      if (!A7_SINGLE_LANE_WAITS) __syncwarp();
      A7_SSTAMP(3);
      cs.next_block(p, stride);
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ MMA issuer of O_q += P_q V_q
    const int q = warp - 2;
    constexpr uint32_t idesc_o = umma_idesc_f16(128, 64);   // M128 N64 K16, two K steps (32 keys)
    const uint32_t tmem_o = tmem_base + 256 + q * 64;
    A7Cursor cp;
    cp.start(p, blockIdx.x, stride);
    if (A7_SINGLE_LANE_WAITS && lane != 0) cp.done = true;
#ifdef LEMAS_ATT_TRACE
    const bool ptr_ = p.trace != nullptr && lane == 0 && q == 0 && (long long)blockIdx.x == p.trace[7];
#define A7_PSTAMP(i) do { if (ptr_ && cp.it == (int)blockIdx.x && cp.j < 32) p.trace[4096 + 128 + cp.j * 4 + (i)] = clock64(); } while (0)
#else
#define A7_PSTAMP(i) do { } while (0)
#endif
    while (!cp.done) {
      const int g = cp.r, buf = g & 1;
      const int s = g % A7_STAGES;
      const bool last = cp.j + 1 == cp.w.n_blocks;
      A7_PSTAMP(0);
      // V^T tile: two 64-key halves of [64 dh rows x 128 B]; part q = keys [32 (q & 1), +32) of half q >> 1
      const uint32_t sv = smem_u32(smem + A7_OFF_KV + s * 2 * A7_TILE_BYTES + A7_TILE_BYTES) + (q >> 1) * (A7_TILE_BYTES / 2);
      const uint32_t tmem_p = tmem_base + buf * 128 + q * 32;
      ATT_WAIT_P(bars + B7_VF + s, (g / A7_STAGES) & 1, 7, cp.j);
      A7_PSTAMP(1);
      ATT_WAIT_P(bars + B7_PF + 2 * q + buf, (g >> 1) & 1, 8, cp.j);
      A7_PSTAMP(2);
      // the first P V of an item overwrites O: the merge of the previous item must have read it
      if (cp.j == 0) ATT_WAIT_P(bars + B7_OE, (cp.qn & 1) ^ 1, 9, cp.it);
      tc_fence_after();
      if (A7_SINGLE_LANE_WAITS || elect_one()) {
        const uint64_t bdesc = umma_desc_sw128(sv) + 4 * (q & 1);   // + 64 B along K
        if (!(p.debug & 2)) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_f16_ts(tmem_o, tmem_p + 8 * ks, bdesc + 2 * ks, idesc_o, (cp.j | ks) != 0 ? 1u : 0u);
        }
        umma_commit(bars + B7_PD + 2 * q + buf);
        umma_commit(bars + B7_VE + s);
        if (last) umma_commit(bars + B7_OF + q);
      }
      if (!A7_SINGLE_LANE_WAITS) __syncwarp();
      A7_PSTAMP(3);
      cp.next_block(p, stride);
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int sw = warp - 6;
    const int q = sw >> 2;             // key part
    const int sub = warp & 3;          // TMEM sub-partition: lanes [32*sub, 32*sub+32)
    const int r = sub * 32 + lane;     // query row inside the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(sub * 32) << 16;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint32_t sb = smem_u32(smem);
    const uint32_t a_sfull = sb + A7_OFF_BAR + B7_SF * 8;                  // + 8 * buffer
    const uint32_t a_pfull0 = sb + A7_OFF_BAR + (B7_PF + 2 * q) * 8;       // + 8 * buffer
    const uint32_t a_pdone0 = sb + A7_OFF_BAR + (B7_PD + 2 * q) * 8;       // + 8 * buffer
    const uint32_t a_ofull = sb + A7_OFF_BAR + B7_OF * 8;                  // + 8 * part
    const uint32_t a_oempty = sb + A7_OFF_BAR + B7_OE * 8;
    const uint32_t tmem_o = tmem_base + 256;
    const uint32_t t_o = tmem_o + lane_addr + q * 64;
    int g = 0, on = 0;                 // KV blocks / live items processed so far
    if (p.dephase_half > 0 && q > 0) {  // the four warps of a sub-partition should not exponentiate in lockstep
      const long long t_go = clock64() + (long long)q * p.dephase_half;
      while (clock64() < t_go) { }
    }
    for (int it = blockIdx.x; it < p.n_items; it += stride) {
      const A7Item w = a7_item(p, it);
      if (!w.live) continue;
      const int q0 = w.q0;
      const int kvl = w.kvl;
      const int n_blocks = w.n_blocks;
      float m_ref = -INFINITY;           // max the accumulator O_q / l are currently scaled by
      float l_run = 0.f;
      // Warps whose 32 query rows all lie beyond the sequence keep the barrier protocol going but do no softmax
      // work: their P rows (left as whatever S held) only feed output rows that are never stored.
      const bool rows_dead = q0 + sub * 32 >= p.seq;
#ifdef LEMAS_ATT_TRACE  // clock64 stamps of the FIRST item of one CTA (tools/trace_att5.py)
      const bool tr_item = p.trace != nullptr && lane == 0 && it == blockIdx.x && (long long)blockIdx.x == p.trace[7];
#endif
      for (int j = 0; j < n_blocks; ++j, ++g) {
        const uint32_t buf = g & 1;
        const uint32_t t_s = tmem_base + buf * 128 + lane_addr + q * 32;
        if (rows_dead) {
          if (!A7_SINGLE_LANE_WAITS || lane == 0) ATT_WAIT_A(a_sfull + 8 * buf, (g >> 1) & 1, 12, j);
          __syncwarp();  // lanes poll independently: reconverge before the single arrival (see attention.cu)
          if (lane == 0) mbar_arrive_s(a_pfull0 + 8 * buf);
          continue;
        }
        const int valid = min(max(kvl - j * 128 - q * 32, 0), 32);  // keys of this part that exist
#ifdef LEMAS_ATT_TRACE
        const bool tr = tr_item && j < 32;
        long long* tp = p.trace + (sw * 32 + j) * 8;
#define A7_STAMP(i) do { if (tr) tp[i] = clock64(); } while (0)
#else
#define A7_STAMP(i) do { } while (0)
#endif
        A7_STAMP(0);
        if (!A7_SINGLE_LANE_WAITS || lane == 0) ATT_WAIT_A(a_sfull + 8 * buf, (g >> 1) & 1, 10, j);   // S(g) landed
        if (A7_SINGLE_LANE_WAITS) __syncwarp();
        A7_STAMP(1);
        tc_fence_after();
        uint32_t sc[32];
        if (kAblate < 4) {
          tmem_ld_32x32(t_s, sc);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) sc[i] = __float_as_uint(-0.01f * (float)(i + lane + g));
        }
        A7_STAMP(2);

        float mx = -INFINITY;
        if (kAblate & 1) {
          mx = __uint_as_float(sc[0]);
        } else if (valid == 32) {  // four independent FMNMX3 chains
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 16; ++i) m4[i & 3] = fmax3f(m4[i & 3], __uint_as_float(sc[i]), __uint_as_float(sc[16 + i]));
          mx = fmaxf(fmax3f(m4[0], m4[1], m4[2]), m4[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < valid) mx = fmaxf(mx, __uint_as_float(sc[i]));
        }
        // lazy rescale: advance the reference max only when this block exceeds it by more than 2^8 (warp-uniform
        // decision, tcgen05.ld/st are warp-collective)
        const bool grow = (mx - m_ref) * c > A7_RESCALE_LOG2;  // also true for the first finite max (m_ref = -inf)
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mx : m_ref;
          const float alpha = (m_ref == -INFINITY) ? 0.f : ex2f((m_ref - m_new) * c);
          l_run *= alpha;
          if (j > 0) {
            // O_q holds the sum of blocks < j only once P V(g-1) has RETIRED; nothing on this warp's path has waited
            // for that (the scores are issued ahead by another thread), so this rare path waits for the commit of
            // P V(g-1) on its buffer's barrier; the next completion there is P V(g+1), which needs this warp: no lapping.
            if (!A7_SINGLE_LANE_WAITS || lane == 0) ATT_WAIT_A(a_pdone0 + 8 * ((g - 1) & 1), ((g - 1) >> 1) & 1, 14, j);
            __syncwarp();
            tc_fence_after();
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
        A7_STAMP(3);

        uint64_t rs2[4] = {0ull, 0ull, 0ull, 0ull};   // bit pattern of (0.f, 0.f)
        uint32_t pk[16];
        const uint64_t c2 = f32x2(c, c), nmc2 = f32x2(-mc, -mc);
        auto exp_block = [&](auto full_tag) {
          constexpr bool kFull = decltype(full_tag)::value;  // full part: no per-element masking code at all
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = 2 * i;
            if (!kFull && col >= valid) {  // warp-uniform: masked key pairs cost no SFU work
              pk[i] = 0u;
              continue;
            }
            float x0, x1;
            f32x2_split(ffma2(f32x2(__uint_as_float(sc[col]), __uint_as_float(sc[col + 1])), c2, nmc2), x0, x1);
            float e0, e1;
            if (kAblate & 2) {
              e0 = x0; e1 = x1;
            } else if (kFull && ((kPolyMask >> i) & 1u)) {
              // exp2 on the FMA / ALU pipes (see attention.cu): Cody-Waite + degree-3 minimax polynomial
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
        if (valid == 32) exp_block(std::true_type{}); else exp_block(std::false_type{});
        A7_STAMP(4);
        // P_q(g) -> TMEM, over the first 16 of the 32 columns its scores were read from: column k = keys (2k, 2k+1)
        if (kAblate < 4) tmem_st_32x32_x16(t_s, pk);
        else if (pk[3] == 0x12345u) l_run += 1.f;
        {
          float lo, hi, lo2, hi2;
          f32x2_split(fadd2(rs2[0], rs2[1]), lo, hi);
          f32x2_split(fadd2(rs2[2], rs2[3]), lo2, hi2);
          l_run += (lo + hi) + (lo2 + hi2);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_s(a_pfull0 + 8 * buf);
        A7_STAMP(6);
      }

      // ---- merge the four key parts of the tile, normalise, store; then hand O back to the MMA issuers
      if (!A7_SINGLE_LANE_WAITS || lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ATT_WAIT_A(a_ofull + 8 * k, on & 1, 15, it);
      }
      __syncwarp();
      tc_fence_after();
      float2* xch = reinterpret_cast<float2*>(smem + A7_OFF_XCH) + (on & 1) * 512;
      xch[q * 128 + r] = make_float2(m_ref, l_run);
      named_bar_sync(1 + sub, 128);  // the four warps that share these 32 rows
      float2 ml[4];
      float m_all = -INFINITY;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ml[k] = xch[k * 128 + r];
        m_all = fmaxf(m_all, ml[k].x);
      }
      float wgt[4], den = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        wgt[k] = (ml[k].x == -INFINITY) ? 0.f : ex2f((ml[k].x - m_all) * c);
        den += wgt[k] * ml[k].y;
      }
      const float inv = 1.0f / den;
      float o[16];  // this warp outputs head-dim columns [16 q, 16 q + 16)
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t ok[16];
        tmem_ld_32x32_x16(tmem_o + lane_addr + k * 64 + q * 16, ok);
        tmem_ld_wait();
        const float wk = wgt[k] * inv;
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] += __uint_as_float(ok[i]) * wk;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_s(a_oempty);   // O is in registers: the next item's first P V may overwrite it
      const int row = q0 + r;
      if (row < p.seq) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((long)w.b * p.seq + row) * p.inner + w.h * 64 + q * 16);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint4 v;
          v.x = pack_half2(o[8 * u + 0], o[8 * u + 1]);
          v.y = pack_half2(o[8 * u + 2], o[8 * u + 3]);
          v.z = pack_half2(o[8 * u + 4], o[8 * u + 5]);
          v.w = pack_half2(o[8 * u + 6], o[8 * u + 7]);
          dst[u] = v;
        }
      }
      ++on;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

template <uint32_t kPolyMask, int kAblate = 0>
static int launch_v7(const CUtensorMap& tmQK, const CUtensorMap& tmVT, const AttnParams& p, void* stream) {
  static unsigned long long configured = 0;
  LEMAS_CUDA_OK(ensure_dynamic_smem(attention7_kernel<kPolyMask, kAblate>, A7_SMEM, configured));
  const int grid = p.n_items < sm_count() ? p.n_items : sm_count();
  LEMAS_CUDA_OK(launch_pdl(attention7_kernel<kPolyMask, kAblate>, dim3(grid), dim3(A7_THREADS), A7_SMEM, (cudaStream_t)stream,
                           tmQK, tmVT, p));
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

// poly: 0 = all exp2 on the SFU, 1 = 1/4 of the key pairs on the FMA pipe, 2 = 3/8
int attention_v7_launch(const CUtensorMap& tmQK, const CUtensorMap& tmVT, const AttnParams& p, int poly, void* stream) {
  switch (poly) {
    case 0: return launch_v7<0u>(tmQK, tmVT, p, stream);
    case 2: return launch_v7<0x2929u>(tmQK, tmVT, p, stream);
    case 11: return launch_v7<0x1111u, 1>(tmQK, tmVT, p, stream);
    case 12: return launch_v7<0x1111u, 2>(tmQK, tmVT, p, stream);
    case 13: return launch_v7<0x1111u, 3>(tmQK, tmVT, p, stream);
    case 14: return launch_v7<0x1111u, 4>(tmQK, tmVT, p, stream);
    default: return launch_v7<0x1111u>(tmQK, tmVT, p, stream);
  }
}

}  // namespace lemas
