// Attention v6: persistent CTA (one per SM), one 128-query tile at a time, the scores DOUBLE-BUFFERED in TMEM so that
// S(j+1) = Q K(j+1)^T is issued BEFORE the softmax of block j has finished.  Same arithmetic as attention.cu (v3):
// softmax(Q K^T / 8 + keymask) V for head_dim 64 (modules.py:483-491), two 64-key halves per KV block with their own
// online-softmax state and accumulator, P written back into TMEM over the scores it came from, TS-form P V.
//
// Why (measured, tools/micro/umma_bench.cu + tools/trace_att5.py, profiles/r02c_*): one thread issues dependent
// tcgen05.mma instructions at ~54 clk per M128 N64 K16 instruction (the tensor pipe itself: 40 clk with several
// issuers, 32 nominal), so "P_x(j) stored -> P V -> next S -> scores visible" costs 600-1000 clk per block in v3 / v5 —
// a third of every softmax warp's block period, during which it cannot touch the SFU.  The scores of the next block
// do not depend on the softmax of this one; only the aliasing of P over S made them wait.  With two score buffers
// (2 x 128 columns) + two accumulators (2 x 64 columns) = 384 of the 512 TMEM columns, the MMA issuer runs one block
// ahead (across work items too) and the softmax warps never wait for the tensor pipe; S is one M128 N128 chain (the
// tensor pipe runs it at its nominal rate) instead of two N64 chains.
//
// Three issuing threads instead of one: a thread is held ~140 clk per dependent tcgen05.mma it issues (a chain of four
// K16 steps ~570 clk whatever N is), and the chains S(g+2), P V_A(g), P V_B(g) are independent of each other.
// Roles (384 threads, up to 168 registers each):
//   warp 0     TMA producer : Q tile of an item into a 2-slot buffer, K / V^T blocks through a 5-slot ring
//   warp 1     MMA issuer of the scores, up to two blocks ahead of the softmax
//   warp 2, 3  MMA issuer of O_A += P_A V_A / O_B += P_B V_B
//   warps 4-11 softmax      : warps 4-7 key half A, 8-11 key half B; thread == query row
// TMEM: score buffer b at column 128 b (half x at +64 x; P_x over its first 32 columns), O_x at 256 + 64 x.
#include <type_traits>

#include "att_common.cuh"

namespace lemas {

constexpr int A6_THREADS = 384;
constexpr int A6_STAGES = 5;
constexpr int A6_TILE_BYTES = 128 * 64 * 2;                    // 16 KB: Q tile, K block, V^T block
constexpr int A6_OFF_Q = 0;                                    // [2 slots]
constexpr int A6_OFF_KV = 2 * A6_TILE_BYTES;                   // [stages][K | V^T]
constexpr int A6_OFF_XCH = A6_OFF_KV + A6_STAGES * 2 * A6_TILE_BYTES;  // float2 [2 parities][2 halves][128]
constexpr int A6_OFF_BAR = A6_OFF_XCH + 2 * 2 * 128 * 8;
constexpr int A6_SMEM = A6_OFF_BAR + 512;                      // 196.5 KB: one CTA per SM

constexpr int B6_QF = 0, B6_QE = 2, B6_KF = 4, B6_KE = B6_KF + A6_STAGES, B6_VF = B6_KE + A6_STAGES,
              B6_VE = B6_VF + A6_STAGES, B6_SF = B6_VE + A6_STAGES, B6_PF = B6_SF + 2, B6_OF = B6_PF + 4,
              B6_OE = B6_OF + 2, B6_PD = B6_OE + 1, B6_COUNT = B6_PD + 4;
static_assert(B6_COUNT * 8 + 8 <= 512, "barrier block");
static_assert(A6_SMEM <= 227 * 1024, "shared memory budget");

constexpr float A6_RESCALE_LOG2 = 8.0f;

struct A6Item {
  int b, h, q0, kvl, n_blocks;
  bool live;
};
DEVI A6Item a6_item(const AttnParams& p, int it) {  // n_pairs holds the number of 128-query tiles per (batch, head)
  A6Item w;
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

// Walks the live work items of this CTA block by block (the MMA issuer keeps two of these: one for the scores it
// issues ahead, one for the P V it is serving).
struct A6Cursor {
  int it, j, qn, r;       // item index, block inside the item, live items / ring entries consumed before this one
  A6Item w;
  bool done;
  DEVI void seek(const AttnParams& p, int stride) {  // advance `it` to the next live item (or the end)
    while (it < p.n_items) {
      w = a6_item(p, it);
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

template <uint32_t kPolyMask>
__global__ void __launch_bounds__(A6_THREADS, 1)
attention6_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVT,
                  const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A6_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B6_COUNT);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("lemas attention6: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmVT);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bars + B6_QF + s, 1);
      mbar_init(bars + B6_QE + s, 1);
      mbar_init(bars + B6_SF + s, 1);       // per score buffer
      // P ready, per (key half, score buffer).  One barrier per half is NOT enough here: S(g+1) is available before
      // every warp has finished block g, so a fast warp would arrive for g+1 in the phase a slow warp still owes its
      // arrival for g.  With one barrier per buffer a warp can only come back to the same barrier for block g+2, whose
      // scores are issued after P V(g), i.e. after all four arrivals for g were observed.
      mbar_init(bars + B6_PF + 2 * s, 4);
      mbar_init(bars + B6_PF + 2 * s + 1, 4);
      mbar_init(bars + B6_OF + s, 1);       // per key half: last P V of an item retired
      mbar_init(bars + B6_PD + 2 * s, 1);   // P V retired, per (key half, score buffer): frees the buffer for the
      mbar_init(bars + B6_PD + 2 * s + 1, 1);  // scores of block g+2 and lets the lazy rescale touch O
    }
    mbar_init(bars + B6_OE, 8);             // the 8 softmax warps
    for (int s = 0; s < A6_STAGES; ++s) {
      mbar_init(bars + B6_KF + s, 1);
      mbar_init(bars + B6_KE + s, 1);
      mbar_init(bars + B6_VF + s, 1);
      mbar_init(bars + B6_VE + s, 2);       // both P V issuers
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
        const A6Item w = a6_item(p, it);
        if (!w.live) continue;
        const int slot = qn & 1;
        ATT_WAIT_P(bars + B6_QE + slot, ((qn >> 1) & 1) ^ 1, 1, it);
        mbar_arrive_expect_tx(bars + B6_QF + slot, A6_TILE_BYTES);
        tma_load_3d(smem + A6_OFF_Q + slot * A6_TILE_BYTES, &tmQK, bars + B6_QF + slot, w.h * 64, w.q0, w.b);
        for (int j = 0; j < w.n_blocks; ++j, ++r) {
          const int s = r % A6_STAGES;
          const uint32_t ph = ((r / A6_STAGES) & 1) ^ 1;
          uint8_t* sk = smem + A6_OFF_KV + s * 2 * A6_TILE_BYTES;
          ATT_WAIT_P(bars + B6_KE + s, ph, 2, j);
          mbar_arrive_expect_tx(bars + B6_KF + s, A6_TILE_BYTES);
          tma_load_3d(sk, &tmQK, bars + B6_KF + s, p.inner + w.h * 64, j * 128, w.b);
          ATT_WAIT_P(bars + B6_VE + s, ph, 3, j);
          mbar_arrive_expect_tx(bars + B6_VF + s, A6_TILE_BYTES);
          tma_load_3d(sk + A6_TILE_BYTES, &tmVT, bars + B6_VF + s, j * 128, 0, w.b * p.heads + w.h);
          tma_load_3d(sk + A6_TILE_BYTES + A6_TILE_BYTES / 2, &tmVT, bars + B6_VF + s, j * 128 + 64, 0,
                      w.b * p.heads + w.h);
        }
        ++qn;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer of the scores
    // S(g) = Q K(g)^T -> score buffer g & 1, as early as the buffer is free: it held P(g-2), which the two P V issuers
    // consume — they are other threads, so their per-buffer pv_done commits are awaited (no cross-thread ordering in
    // the tensor pipe).  Runs ahead of the softmax by up to two blocks, across work items.
    constexpr uint32_t idesc_s = umma_idesc_f16(128, 128);  // M128 N128 (both key halves) K16 x4
    A6Cursor cs;
    cs.start(p, blockIdx.x, stride);
    while (!cs.done) {
      const int g = cs.r, buf = g & 1;
      if (cs.j == 0) ATT_WAIT_P(bars + B6_QF + (cs.qn & 1), (cs.qn >> 1) & 1, 4, cs.it);
      const int s = g % A6_STAGES;
      ATT_WAIT_P(bars + B6_KF + s, (g / A6_STAGES) & 1, 5, cs.j);
      if (g >= 2) {  // completion (g-2)/2 of the buffer's pv_done barriers; the next one needs S(g): no lapping
        ATT_WAIT_P(bars + B6_PD + buf, ((g - 2) >> 1) & 1, 6, cs.j);
        ATT_WAIT_P(bars + B6_PD + 2 + buf, ((g - 2) >> 1) & 1, 6, cs.j);
      }
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = umma_desc_sw128(smem_u32(smem + A6_OFF_Q + (cs.qn & 1) * A6_TILE_BYTES));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + A6_OFF_KV + s * 2 * A6_TILE_BYTES));
        const uint32_t d = tmem_base + buf * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        umma_commit(bars + B6_SF + buf);
        umma_commit(bars + B6_KE + s);
        if (cs.j + 1 == cs.w.n_blocks) umma_commit(bars + B6_QE + (cs.qn & 1));  // last read of this Q tile
      }
      __syncwarp();
      cs.next_block(p, stride);
    }
  } else if (warp == 2 || warp == 3) {
    // ------------------------------------------------------------------ MMA issuer of O_x += P_x V_x, x = warp - 2
    const int x = warp - 2;
    constexpr uint32_t idesc_o = umma_idesc_f16(128, 64);   // M128 N64 K16 x4
    const uint32_t tmem_o = tmem_base + 256 + x * 64;
    A6Cursor cp;
    cp.start(p, blockIdx.x, stride);
#ifdef LEMAS_ATT_TRACE
    const bool mtr = p.trace != nullptr && lane == 0 && (long long)blockIdx.x == p.trace[7];
#define A6_MSTAMP(i) do { if (mtr && cp.it == (int)blockIdx.x && cp.j < 32) p.trace[4096 + (cp.j * 2 + x) * 2 + (i)] = clock64(); } while (0)
#else
#define A6_MSTAMP(i) do { } while (0)
#endif
    while (!cp.done) {
      const int g = cp.r, buf = g & 1;
      const int s = g % A6_STAGES;
      const bool last = cp.j + 1 == cp.w.n_blocks;
      const uint32_t sv = smem_u32(smem + A6_OFF_KV + s * 2 * A6_TILE_BYTES + A6_TILE_BYTES) + x * (A6_TILE_BYTES / 2);
      const uint32_t tmem_p = tmem_base + buf * 128 + x * 64;
      ATT_WAIT_P(bars + B6_VF + s, (g / A6_STAGES) & 1, 7, cp.j);
      ATT_WAIT_P(bars + B6_PF + 2 * x + buf, (g >> 1) & 1, 8 + x, cp.j);
      // the first P V of an item overwrites O: the merge of the previous item must have read it
      if (cp.j == 0) ATT_WAIT_P(bars + B6_OE, (cp.qn & 1) ^ 1, 10, cp.it);
      A6_MSTAMP(0);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bdesc = umma_desc_sw128(sv);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16_ts(tmem_o, tmem_p + 8 * ks, bdesc + 2 * ks, idesc_o, (cp.j | ks) != 0 ? 1u : 0u);
        umma_commit(bars + B6_PD + 2 * x + buf);
        umma_commit(bars + B6_VE + s);
        if (last) umma_commit(bars + B6_OF + x);
      }
      __syncwarp();
      A6_MSTAMP(1);
      cp.next_block(p, stride);
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int sw = warp - 4;
    const int half = sw >> 2;
    const int sub = warp & 3;          // TMEM sub-partition: lanes [32*sub, 32*sub+32)
    const int r = sub * 32 + lane;     // query row inside the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(sub * 32) << 16;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint32_t sb = smem_u32(smem);
    const uint32_t a_sfull = sb + A6_OFF_BAR + B6_SF * 8;             // + 8 * buffer
    const uint32_t a_pfull0 = sb + A6_OFF_BAR + (B6_PF + 2 * half) * 8;   // + 8 * buffer
    const uint32_t a_ofull = sb + A6_OFF_BAR + B6_OF * 8;             // + 8 * half
    const uint32_t a_oempty = sb + A6_OFF_BAR + B6_OE * 8;
    const uint32_t tmem_o = tmem_base + 256;
    const uint32_t t_o = tmem_o + lane_addr + half * 64;
    int g = 0, on = 0;                 // KV blocks / live items processed so far
    if (half == 1 && p.dephase_half > 0) {  // the two warps of a sub-partition should not exponentiate in lockstep
      const long long t_go = clock64() + p.dephase_half;
      while (clock64() < t_go) { }
    }
    for (int it = blockIdx.x; it < p.n_items; it += stride) {
      const A6Item w = a6_item(p, it);
      if (!w.live) continue;
      const int q0 = w.q0;
      const int kvl = w.kvl;
      const int n_blocks = w.n_blocks;
      float m_ref = -INFINITY;           // max the accumulators O_half / l are currently scaled by
      float l_run = 0.f;
      // Warps whose 32 query rows all lie beyond the sequence keep the barrier protocol going but do no softmax
      // work: their P rows (left as whatever S held) only feed output rows that are never stored.
      const bool rows_dead = q0 + sub * 32 >= p.seq;
#ifdef LEMAS_ATT_TRACE  // clock64 stamps of the FIRST item of one CTA (tools/trace_att5.py)
      const bool tr_item = p.trace != nullptr && lane == 0 && it == blockIdx.x && (long long)blockIdx.x == p.trace[7];
#endif
      for (int j = 0; j < n_blocks; ++j, ++g) {
        const uint32_t buf = g & 1;
        const uint32_t t_s = tmem_base + buf * 128 + lane_addr + half * 64;
        if (rows_dead) {
          ATT_WAIT_A(a_sfull + 8 * buf, (g >> 1) & 1, 12 + half, j);
          __syncwarp();  // lanes poll independently: reconverge before the single arrival (see attention.cu)
          if (lane == 0) mbar_arrive_s(a_pfull0 + 8 * buf);
          continue;
        }
        const int valid = min(max(kvl - j * 128 - half * 64, 0), 64);  // keys of this half-block that exist
#ifdef LEMAS_ATT_TRACE
        const bool tr = tr_item && j < 32;
        long long* tp = p.trace + (sw * 32 + j) * 8;
#define A6_STAMP(i) do { if (tr) tp[i] = clock64(); } while (0)
#else
#define A6_STAMP(i) do { } while (0)
#endif
        A6_STAMP(0);
        ATT_WAIT_A(a_sfull + 8 * buf, (g >> 1) & 1, 10 + half, j);   // S(j) landed (normally long ago)
        A6_STAMP(1);
        tc_fence_after();
        uint32_t s0[32], s1[32];
        tmem_ld_32x32(t_s, s0);
        tmem_ld_32x32(t_s + 32, s1);
        tmem_ld_wait();
        A6_STAMP(2);

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
        const bool grow = (mx - m_ref) * c > A6_RESCALE_LOG2;  // also true for the first finite max (m_ref = -inf)
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mx : m_ref;
          const float alpha = (m_ref == -INFINITY) ? 0.f : ex2f((m_ref - m_new) * c);
          l_run *= alpha;
          if (j > 0) {
            // O_half holds the sum of blocks < j only once P V(j-1) has RETIRED; nothing on this warp's path has waited
            // for that (the scores are issued ahead by another thread), so this rare path waits for the commit of
            // P V(g-1) on its buffer's barrier; the next completion there is P V(g+1), which needs this warp: no lapping.
            ATT_WAIT_A(sb + A6_OFF_BAR + (B6_PD + 2 * half + ((g - 1) & 1)) * 8, ((g - 1) >> 1) & 1, 14, j);
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
        A6_STAMP(3);

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
        if (valid == 64) exp_block(std::true_type{}); else exp_block(std::false_type{});
        A6_STAMP(4);
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
        if (lane == 0) mbar_arrive_s(a_pfull0 + 8 * buf);
        A6_STAMP(6);
      }

      // ---- merge the two key halves of the tile, normalise, store; then hand O back to the MMA issuer
      ATT_WAIT_A(a_ofull, on & 1, 15, it);
      ATT_WAIT_A(a_ofull + 8, on & 1, 16, it);
      tc_fence_after();
      float2* xch = reinterpret_cast<float2*>(smem + A6_OFF_XCH) + (on & 1) * 256;
      xch[half * 128 + r] = make_float2(m_ref, l_run);
      named_bar_sync(1 + sub, 64);  // the two warps that share these 32 rows
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
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

template <uint32_t kPolyMask>
static int launch_v6(const CUtensorMap& tmQK, const CUtensorMap& tmVT, const AttnParams& p, void* stream) {
  static unsigned long long configured = 0;
  LEMAS_CUDA_OK(ensure_dynamic_smem(attention6_kernel<kPolyMask>, A6_SMEM, configured));
  const int grid = p.n_items < sm_count() ? p.n_items : sm_count();
  LEMAS_CUDA_OK(launch_pdl(attention6_kernel<kPolyMask>, dim3(grid), dim3(A6_THREADS), A6_SMEM, (cudaStream_t)stream,
                           tmQK, tmVT, p));
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

// poly: 0 = all exp2 on the SFU, 1 = 1/4 of the key pairs on the FMA pipe, 2 = 3/8
int attention_v6_launch(const CUtensorMap& tmQK, const CUtensorMap& tmVT, const AttnParams& p, int poly, void* stream) {
  switch (poly) {
    case 0: return launch_v6<0u>(tmQK, tmVT, p, stream);
    case 2: return launch_v6<0x49249249u>(tmQK, tmVT, p, stream);
    default: return launch_v6<0x11111111u>(tmQK, tmVT, p, stream);
  }
}

}  // namespace lemas
