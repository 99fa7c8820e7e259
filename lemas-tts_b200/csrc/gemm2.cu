// CTA-pair tcgen05 GEMM for sm_100a: 256 x 256 output tiles computed by two CTAs on the two SMs of a TPC
// (tcgen05.mma.cta_group::2, M = 256, N = 256, K = 16 per instruction).
//
//   acc[m, n] = sum_k A[m, k] * W[n, k]            fp16 operands, fp32 accumulation in TMEM
//
// Why pairs: a single-CTA 128 x 256 tile needs 48 KB of operands per 512 tensor-core clocks (94 B/clk/SM); the L2
// can deliver about 43 B/clk/SM with every SM streaming, so the main loop ran at ~45 % of the MMA rate.  In a pair
// each CTA stages its own 128 rows of A and only HALF of the B tile (128 of the 256 weight rows); the MMA reads the
// other half from the peer's shared memory.  32 KB per 512 clocks = 64 B/clk/SM — the same operand economy as the
// 256 x 256 2-SM tiles cuBLAS uses on this part.
//
// Roles per CTA (192 threads), persistent over tiles:
//   warp 0    TMA producer : both CTAs load their A rows and their half of B; completion bytes are credited to the
//                            LEADER CTA's full barrier (cp.async.bulk.tensor ... cta_group::2)
//   warp 1    MMA issuer   : leader CTA only; tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs
//   warps 2-9 epilogue     : each CTA drains its own 128 TMEM lanes x 256 columns (two warps per 32-lane
//                            sub-partition, 128 columns each); double-buffered accumulators
//                            (2 x 256 TMEM columns) let the epilogue of tile i overlap the main loop of tile i+1
//
// Epilogue I/O is coalesced: tcgen05.ld gives each lane one ROW (32 consecutive columns); the warp transposes the
// 32 x 32 fp32 block through a padded shared-memory buffer so that, for global memory, 8 lanes cover 32 consecutive
// columns of one row (128 B of fp32 / 64 B of fp16) — the residual read, the gated residual write and the fp16
// activations all move as full sectors.  Only the transposed V copy is written lane-per-row (its fast axis is the
// sequence position).
#include "common.h"
#include "gemm_params.cuh"
#include "ptx.cuh"

namespace lemas {

constexpr int G2_BM = 128;   // rows per CTA (256 per pair)
constexpr int G2_BN = 256;   // columns per pair tile
constexpr int G2_BK = 64;
constexpr int G2_STAGES = 5;
constexpr int G2_EPI_WARPS = 8;                        // two warps per TMEM sub-partition, each takes half the columns
constexpr int G2_THREADS = 64 + 32 * G2_EPI_WARPS;
constexpr int G2_A_BYTES = G2_BM * G2_BK * 2;          // 16 KB
constexpr int G2_B_BYTES = (G2_BN / 2) * G2_BK * 2;    // 16 KB: this CTA's half of the B tile
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_PITCH = 36;                           // floats per staged row (32 + 4: conflict-free float4 access)
constexpr int G2_STG_BYTES = G2_EPI_WARPS * 32 * G2_PITCH * 4;
constexpr int G2_OFF_STG = G2_STAGES * G2_STAGE_BYTES;
constexpr int G2_OFF_LNS = G2_OFF_STG + G2_STG_BYTES;                 // folded LayerNorm: float2 [epi warps][32 rows]
constexpr int G2_OFF_BAR = G2_OFF_LNS + G2_EPI_WARPS * 32 * 8;
constexpr int G2_SMEM = G2_OFF_BAR + 256 + 1024;

DEVI uint2 pack4_half(float4 v) { return make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w)); }

// Row bookkeeping of one epilogue warp: 32 consecutive rows of ONE batch item (tiles never straddle batch items, so
// no per-row division is needed and sequence positions of a warp start at a multiple of 32).
struct RowCtx {
  int b;          // batch item
  int pos0;       // sequence position of the warp's first row
  int nrows;      // valid rows of the warp (<= 32; rows beyond the sequence are skipped)
  long grow0;     // global row of the warp's first row = b * rows + pos0
};

// Global loads a 32 x 32 block needs besides the accumulator (residual rows / RoPE cos,sin), fetched one block
// ahead of its use so that their L2 latency overlaps the previous block's math and stores.
template <int EPI>
DEVI void epi2_prefetch(const GemmParams& p, const RowCtx& rc, int col0, int lane, float4 (&pre)[8]) {
  const int c4 = lane & 7, rsub = lane >> 3;
  const int col = col0 + c4 * 4;
  if constexpr (EPI == LEMAS_EPI_GATE_RESID_F32) {
    if (p.red_add) return;   // in-place reduction: the residual is never read by the SM
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + rsub;
      if (rr < rc.nrows) pre[it] = *reinterpret_cast<const float4*>(p.resid + (rc.grow0 + rr) * p.ldr + col);
    }
  } else if constexpr (EPI == LEMAS_EPI_QKV_ROPE) {
    if (col0 < 2 * p.inner) {
      const int within = col % p.inner;
      const bool rot = within < p.rope_cols;
      const int pair0 = (within & 63) >> 1;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + rsub;
        pre[it] = make_float4(1.f, 0.f, 1.f, 0.f);
        if (rot && rr < rc.nrows)
          pre[it] = __ldg(reinterpret_cast<const float4*>(p.rope + (long)(rc.pos0 + rr) * 32 + pair0));
      }
    }
  }
}

// One 32-row x 32-column block of the accumulator.  r[i] = acc[row lane of the warp][col0 + i].
// bias4 / gate4: this thread's 4 columns (col0 + 4*(lane&7) ..) of the per-column vectors.
// Folded LayerNorm, per-thread state of one tile.  Consumer side: mean / rstd of the 8 rows this thread touches in the
// transposed layout (row it*4 + lane/8) and of row `lane` (the lane-per-row V path); producer side: running partial
// (sum, sum of squares) of those 8 rows over the warp's 128 columns.
struct LnState {
  float mu[8], rs[8];
  float mu_l, rs_l;
  float2* acc;   // producer: this warp's shared-memory accumulators, one (sum, sum of squares) per row
};

template <int EPI, bool FOLD>
DEVI void epi2_block(const GemmParams& p, float* stg, const uint32_t (&r)[32], const RowCtx& rc, int col0, int lane,
                     float4 bias4, float4 gate4, const float4 (&pre)[8], LnState& ln, float4 lnu4, float4 lnv4,
                     const float* ln_u, const float* ln_v) {
  if constexpr (EPI == LEMAS_EPI_QKV_ROPE) {
    if (col0 >= 2 * p.inner) {
      // V: transposed copy vt[b, head, d, pos].  The block is transposed through shared memory as fp16 so that each
      // lane stores 8 consecutive positions (16 B) of one head-dim row instead of 32 scattered 2-byte stores.
      const int vcol = col0 - 2 * p.inner;
      const int heads = p.inner >> 6;
      __half* sth = reinterpret_cast<__half*>(stg);  // [32 d][40 pos] halves (80 B rows: conflict-free 16 B reads)
      float4 bv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        bv[j] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      constexpr bool fold = FOLD;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 a = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                               __uint_as_float(r[4 * j + 3]));
        if (fold) {  // acc -> rstd (acc - mean u) + v for this lane's row
          const float4 u = __ldg(reinterpret_cast<const float4*>(ln_u + col0) + j);
          const float4 w = __ldg(reinterpret_cast<const float4*>(ln_v + col0) + j);
          a.x = ln.rs_l * (a.x - ln.mu_l * u.x) + w.x; a.y = ln.rs_l * (a.y - ln.mu_l * u.y) + w.y;
          a.z = ln.rs_l * (a.z - ln.mu_l * u.z) + w.z; a.w = ln.rs_l * (a.w - ln.mu_l * u.w) + w.w;
        }
        sth[(4 * j + 0) * 40 + lane] = __float2half_rn(a.x + bv[j].x);
        sth[(4 * j + 1) * 40 + lane] = __float2half_rn(a.y + bv[j].y);
        sth[(4 * j + 2) * 40 + lane] = __float2half_rn(a.z + bv[j].z);
        sth[(4 * j + 3) * 40 + lane] = __float2half_rn(a.w + bv[j].w);
      }
      __syncwarp();
      __half* dst0 = p.vt + ((long)(rc.b * heads + (vcol >> 6)) * 64 + (vcol & 63)) * p.vt_ld + rc.pos0;
      const int oct = lane & 3, dsub = lane >> 2;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int d = it * 8 + dsub;
        const uint4 v = *reinterpret_cast<const uint4*>(sth + d * 40 + oct * 8);
        __half* dst = dst0 + (long)d * p.vt_ld + oct * 8;
        if (oct * 8 + 8 <= rc.nrows) {
          *reinterpret_cast<uint4*>(dst) = v;  // pos0 % 32 == 0 and vt_ld % 8 == 0: 16-byte aligned
        } else {
          const __half* hv = reinterpret_cast<const __half*>(&v);
          for (int i = 0; i < 8; ++i)
            if (oct * 8 + i < rc.nrows) dst[i] = hv[i];
        }
      }
      __syncwarp();
      return;
    }
  }
  const int c4 = lane & 7, rsub = lane >> 3;
  const int col = col0 + c4 * 4;

  // ---- transpose through shared memory: lane-per-row -> 8 lanes per row
  float4* srow = reinterpret_cast<float4*>(stg + lane * G2_PITCH);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    srow[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
  __syncwarp();

#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + rsub;
    if (rr >= rc.nrows) continue;
    const long grow = rc.grow0 + rr;
    float4 v = *reinterpret_cast<const float4*>(stg + rr * G2_PITCH + c4 * 4);
    if constexpr (FOLD && (EPI == LEMAS_EPI_QKV_ROPE || EPI == LEMAS_EPI_GELU_TANH_F16)) {
      {  // folded LayerNorm: acc -> rstd (acc - mean u[col]) + v[col]
        const float m = ln.mu[it], rsd = ln.rs[it];
        v.x = rsd * (v.x - m * lnu4.x) + lnv4.x; v.y = rsd * (v.y - m * lnu4.y) + lnv4.y;
        v.z = rsd * (v.z - m * lnu4.z) + lnv4.z; v.w = rsd * (v.w - m * lnu4.w) + lnv4.w;
      }
    }
    v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
    if constexpr (EPI == LEMAS_EPI_GATE_RESID_F32) {
      float4 g = gate4;
      if (p.gate != nullptr && p.gate_bstride != 0)
        g = __ldg(reinterpret_cast<const float4*>(p.gate + (long)rc.b * p.gate_bstride + col));
      const bool dead = p.row_valid != nullptr && (rc.pos0 + rr) >= __ldg(p.row_valid + rc.b);
      if (!FOLD && p.red_add) {
        // In place (resid == out32, the DiT residual stream): x += gate (acc + bias) as ONE 16-byte reduction executed
        // by the L2 — the SM neither reads the residual (half of this epilogue's bytes and its only dependent global
        // load) nor waits for it.  One reduction per element per launch: deterministic.  Measured in the step: to_out
        // -12 % at C2 and -29 % at C4, C2 126.3 -> 123.5 ms (profiles/r02aa_red_add_epilogue.log).
        if (!dead)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.out32 + grow * p.ld32 + col),
                       "f"(g.x * v.x), "f"(g.y * v.y), "f"(g.z * v.z), "f"(g.w * v.w) : "memory");
        continue;
      }
      float4 o = pre[it];
      if (!dead) { o.x += g.x * v.x; o.y += g.y * v.y; o.z += g.z * v.z; o.w += g.w * v.w; }
      *reinterpret_cast<float4*>(p.out32 + grow * p.ld32 + col) = o;
      if constexpr (FOLD) {
        // folded LayerNorm, producer: the next GEMM's A operand is x_new (1 + scale) — UN-normalised, the row
        // statistics are applied in that GEMM's epilogue — plus this row's partial sums over the thread's 4 columns,
        // reduced over the 8 lanes that share the row (lnu4 carries the fp16-rounded 1 + scale here)
        *reinterpret_cast<uint2*>(p.ln_out16 + grow * p.ln_ld16 + col) =
            pack4_half(make_float4(o.x * lnu4.x, o.y * lnu4.y, o.z * lnu4.z, o.w * lnu4.w));
        float sm = (o.x + o.y) + (o.z + o.w);
        float sq = (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
        const unsigned grp = 0xFFu << (rsub * 8);
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
          sm += __shfl_xor_sync(grp, sm, off);
          sq += __shfl_xor_sync(grp, sq, off);
        }
        if (c4 == 0) {  // one lane per row owns the accumulator; the 4 column blocks of a tile come one after the other
          float2 a = ln.acc[rr];
          a.x += sm;
          a.y += sq;
          ln.acc[rr] = a;
        }
      }
    } else if constexpr (EPI == LEMAS_EPI_BIAS_F16) {
      *reinterpret_cast<uint2*>(p.out16 + grow * p.ld16 + col) = pack4_half(v);
    } else if constexpr (EPI == LEMAS_EPI_GELU_TANH_F16) {
      v.x = gelu_tanh_f(v.x); v.y = gelu_tanh_f(v.y); v.z = gelu_tanh_f(v.z); v.w = gelu_tanh_f(v.w);
      *reinterpret_cast<uint2*>(p.out16 + grow * p.ld16 + col) = pack4_half(v);
    } else if constexpr (EPI == LEMAS_EPI_GELU_ERF_F16) {
      v.x = gelu_erf_f(v.x); v.y = gelu_erf_f(v.y); v.z = gelu_erf_f(v.z); v.w = gelu_erf_f(v.w);
      *reinterpret_cast<uint2*>(p.out16 + grow * p.ld16 + col) = pack4_half(v);
    } else if constexpr (EPI == LEMAS_EPI_BIAS_F32) {
      *reinterpret_cast<float4*>(p.out32 + grow * p.ld32 + col) = v;
    } else if constexpr (EPI == LEMAS_EPI_QKV_ROPE) {  // q | k columns (V handled above); cs = (cos0, sin0, cos1, sin1)
      const float4 cs = pre[it];
      const float x0 = v.x, x1 = v.y, x2 = v.z, x3 = v.w;
      v.x = x0 * cs.x - x1 * cs.y; v.y = x1 * cs.x + x0 * cs.y;
      v.z = x2 * cs.z - x3 * cs.w; v.w = x3 * cs.z + x2 * cs.w;
      *reinterpret_cast<uint2*>(p.out16 + grow * p.ld16 + col) = pack4_half(v);
    }
  }
  __syncwarp();  // the staging buffer is rewritten by the next block
}

// Ragged batches: a 256-row pair tile that starts at or beyond its sequence's row limit is skipped by every role
// (producer, MMA issuer, epilogue evaluate the same predicate, so ring / accumulator phases stay in step).
DEVI bool tile_skipped(const GemmParams& p, int batch_item, int first_row) {
  return p.row_limit != nullptr && first_row >= __ldg(p.row_limit + batch_item);
}

// BN = 256: the production tile.  BN = 128 (256 x 128 pair tiles, same shared-memory layout, half the B bytes per
// stage): chosen by gemm2_launch when the 256-wide tiling would leave more than half of the CTA pairs idle (small M:
// one short utterance) — twice the tiles, half the serial MMA chain per tile.
template <int EPI, bool FOLD, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_OFF_BAR);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* acc_full = empty_bar + G2_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int m_tiles_b = (p.rows + 2 * G2_BM - 1) / (2 * G2_BM);  // per batch item: tiles never straddle items
  const int n_tiles = p.n / BN;
  const int num_tiles = p.batches * m_tiles_b * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(full_bar + s, 1);    // leader's own arrive.expect_tx; both CTAs' TMA bytes are credited here
      mbar_init(empty_bar + s, 1);   // one multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full + a, 1);    // one multicast tcgen05.commit
      mbar_init(acc_empty + a, 2 * G2_EPI_WARPS);   // every epilogue warp of both CTAs (used in the leader only)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();  // the next kernel in the stream may start its own prologue
  pdl_wait();     // everything above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int n_idx = tile % n_tiles;
        const int mt = tile / n_tiles;
        const int bi = mt / m_tiles_b;
        if (tile_skipped(p, bi, (mt - bi * m_tiles_b) * (2 * G2_BM))) continue;
        const int m0 = (mt - bi * m_tiles_b) * (2 * G2_BM) + (int)rank * G2_BM;
        const int n0 = n_idx * BN + (int)rank * (BN / 2);
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(full_bar + stage, 2 * (G2_A_BYTES + (BN / 2) * G2_BK * 2));
          const uint32_t full_leader = mapa_shared(smem_u32(full_bar + stage), 0);
          uint8_t* sa = smem + stage * G2_STAGE_BYTES;
          tma_load_3d_pair(sa, &tmA, full_leader, it * G2_BK, m0, bi);
          tma_load_2d_pair(sa + G2_A_BYTES, &tmW, full_leader, it * G2_BK, n0);
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * G2_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        {
          const int mt = tile / n_tiles, bi = mt / m_tiles_b;
          if (tile_skipped(p, bi, (mt - bi * m_tiles_b) * (2 * G2_BM))) continue;
        }
        mbar_wait_cluster(acc_empty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
            const uint64_t adesc = umma_desc_sw128(sa);
            const uint64_t bdesc = umma_desc_sw128(sa + G2_A_BYTES);
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k)
              umma_f16_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
            umma_commit_pair(empty_bar + stage, 3);
            if (it == p.k_iters - 1) umma_commit_pair(acc_full + acc, 3);
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int sub = warp & 3;           // TMEM sub-partition of this warp: lanes [32*sub, 32*sub+32)
    const int half = (warp - 2) >> 2;   // which 128 columns of the 256-wide tile this warp drains
    float* stg = reinterpret_cast<float*>(smem + G2_OFF_STG) + (warp - 2) * 32 * G2_PITCH;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int n_idx = tile % n_tiles;
      const int mt = tile / n_tiles;
      RowCtx rc;
      rc.b = mt / m_tiles_b;
      if (tile_skipped(p, rc.b, (mt - rc.b * m_tiles_b) * (2 * G2_BM))) continue;
      rc.pos0 = (mt - rc.b * m_tiles_b) * (2 * G2_BM) + (int)rank * G2_BM + sub * 32;
      rc.nrows = min(32, p.rows - rc.pos0);
      rc.grow0 = (long)rc.b * p.rows + rc.pos0;
      const int n0 = n_idx * BN;
      // Per-column vectors (bias, gate) for this thread's 4 columns of each 32-column block are fetched one block
      // ahead; block 0's are issued before the accumulator wait so their latency hides behind the main loop.
      // The block loop stays rolled: these kernels run a few microseconds, every instruction executes only a
      // handful of times, and a fully unrolled epilogue (~6 k instructions) was instruction-fetch bound.
      const int ccol = n0 + (lane & 7) * 4;
      auto load_bias = [&](int j) {
        return p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + ccol + j * 32)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      auto load_gate = [&](int j) {
        if constexpr (EPI == LEMAS_EPI_GATE_RESID_F32)
          if (p.gate != nullptr && p.gate_bstride == 0) return __ldg(reinterpret_cast<const float4*>(p.gate + ccol + j * 32));
        return make_float4(1.f, 1.f, 1.f, 1.f);
      };
      static_assert(!FOLD || BN == 256, "the folded LayerNorm is built for 256-column tiles");
      constexpr int NB = BN / 64;   // 32-column blocks per epilogue warp (the warp owns half of the tile's columns)
      const int j0 = half * NB;
      float4 bias_nxt = load_bias(j0), gate_nxt = load_gate(j0);
      const uint32_t t_addr = tmem_base + (uint32_t(sub * 32) << 16) + acc * BN;
      if (rc.nrows > 0) {
        float4 pre_nxt[8];
        epi2_prefetch<EPI>(p, rc, n0 + j0 * 32, lane, pre_nxt);  // does not depend on the accumulator either
        // ---- folded LayerNorm (see lemas_gemm_desc): per-tile set-up, all from global memory written by earlier kernels
        LnState ln;
        const float* ln_u = nullptr;
        const float* ln_v = nullptr;
        constexpr bool ln_cons = FOLD && (EPI == LEMAS_EPI_QKV_ROPE || EPI == LEMAS_EPI_GELU_TANH_F16);
        constexpr bool ln_prod = FOLD && EPI == LEMAS_EPI_GATE_RESID_F32;
        if (ln_cons) {
          const int step = __ldg(p.ln_step) - 1;
          ln_u = p.ln_uv + (long)(2 * step) * p.n;
          ln_v = ln_u + p.n;
          const int c4 = lane & 7, rsub = lane >> 3;
          const unsigned grp = 0xFFu << (rsub * 8);
#pragma unroll
          for (int it = 0; it < 8; ++it) {  // lane c4 fetches partial c4 of row it*4 + rsub; the 8 lanes add them up
            const int rr = it * 4 + rsub;
            float2 pt = make_float2(0.f, 0.f);
            if (rr < rc.nrows && c4 < p.ln_parts)
              pt = __ldg(reinterpret_cast<const float2*>(p.ln_stats_in) + (rc.grow0 + rr) * p.ln_parts + c4);
#pragma unroll
            for (int off = 1; off < 8; off <<= 1) {
              pt.x += __shfl_xor_sync(grp, pt.x, off);
              pt.y += __shfl_xor_sync(grp, pt.y, off);
            }
            const float mean = pt.x * p.ln_inv_k;
            ln.mu[it] = mean;
            ln.rs[it] = rsqrtf(fmaxf(pt.y * p.ln_inv_k - mean * mean, 0.f) + 1e-6f);
          }
          float sl = 0.f, ql = 0.f;  // row `lane` (the transposed V copy works lane-per-row)
          if (lane < rc.nrows)
            for (int k = 0; k < p.ln_parts; ++k) {
              const float2 pt = __ldg(reinterpret_cast<const float2*>(p.ln_stats_in) + (rc.grow0 + lane) * p.ln_parts + k);
              sl += pt.x;
              ql += pt.y;
            }
          ln.mu_l = sl * p.ln_inv_k;
          ln.rs_l = rsqrtf(fmaxf(ql * p.ln_inv_k - ln.mu_l * ln.mu_l, 0.f) + 1e-6f);
        }
        if (ln_prod) {
          ln.acc = reinterpret_cast<float2*>(smem + G2_OFF_LNS) + (warp - 2) * 32;
          ln.acc[lane] = make_float2(0.f, 0.f);
          __syncwarp();
        }
        auto load_lnu = [&](int j) {  // consumer: u[col]; producer: fp16-rounded 1 + scale[col]
          if (ln_cons) return __ldg(reinterpret_cast<const float4*>(ln_u + ccol + j * 32));
          if (ln_prod) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.ln_scale + ccol + j * 32));
            return make_float4(__half2float(__float2half_rn(1.f + sc.x)), __half2float(__float2half_rn(1.f + sc.y)),
                               __half2float(__float2half_rn(1.f + sc.z)), __half2float(__float2half_rn(1.f + sc.w)));
          }
          return make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto load_lnv = [&](int j) {
          return ln_cons ? __ldg(reinterpret_cast<const float4*>(ln_v + ccol + j * 32)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        float4 lnu_nxt = make_float4(0.f, 0.f, 0.f, 0.f), lnv_nxt = lnu_nxt;
        if constexpr (FOLD) { lnu_nxt = load_lnu(j0); lnv_nxt = load_lnv(j0); }
        mbar_wait(acc_full + acc, acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int j = j0; j < j0 + NB; ++j) {
          const float4 bias4 = bias_nxt, gate4 = gate_nxt, lnu4 = lnu_nxt, lnv4 = lnv_nxt;
          float4 pre[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) pre[i] = pre_nxt[i];
          if (j < j0 + NB - 1) {
            bias_nxt = load_bias(j + 1);
            gate_nxt = load_gate(j + 1);
            if constexpr (FOLD) { lnu_nxt = load_lnu(j + 1); lnv_nxt = load_lnv(j + 1); }
            epi2_prefetch<EPI>(p, rc, n0 + (j + 1) * 32, lane, pre_nxt);
          }
          uint32_t r[32];
          tmem_ld_32x32(t_addr + j * 32, r);
          tmem_ld_wait();
          epi2_block<EPI, FOLD>(p, stg, r, rc, n0 + j * 32, lane, bias4, gate4, pre, ln, lnu4, lnv4, ln_u, ln_v);
        }
        if (ln_prod) {  // this warp's 128-column partial of its 32 rows
          __syncwarp();
          const int parts = 2 * n_tiles, part = n_idx * 2 + half;
          if (lane < rc.nrows) reinterpret_cast<float2*>(p.ln_stats)[(rc.grow0 + lane) * parts + part] = ln.acc[lane];
        }
      } else {
        mbar_wait(acc_full + acc, acc_phase);
        tc_fence_after();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(acc_empty + acc), 0));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still touch this CTA's memory / barriers
  if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

template <int EPI, bool FOLD, int BN>
static int launch2_bn(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmParams& p, int max_ctas, cudaStream_t st) {
  static unsigned long long configured = 0;
  auto kern = gemm2_kernel<EPI, FOLD, BN>;
  LEMAS_CUDA_OK(ensure_dynamic_smem(kern, G2_SMEM, configured));
  const int tiles = p.batches * ((p.rows + 2 * G2_BM - 1) / (2 * G2_BM)) * (p.n / BN);
  int pairs = (max_ctas > 0 ? max_ctas : sm_count()) / 2;
  if (pairs < 1) pairs = 1;
  if (pairs > tiles) pairs = tiles;
  LEMAS_CUDA_OK(launch_pdl(kern, dim3(2 * pairs), dim3(G2_THREADS), G2_SMEM, st, tmA, tmW, p));
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

template <int EPI, bool FOLD = false>
static int launch2(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmParams& p, int max_ctas, cudaStream_t st,
                   int bn = 256) {
  if constexpr (!FOLD)
    if (bn == 128) return launch2_bn<EPI, false, 128>(tmA, tmW, p, max_ctas, st);
  return launch2_bn<EPI, FOLD, 256>(tmA, tmW, p, max_ctas, st);
}

// 256 x 128 tiles for small problems when they need fewer (weighted) waves than 256 x 256 ones — one short utterance:
// to_out 12.7 -> 9.2 us, FF2 17.7 -> 12.7 us, QKV 18.6 -> 16.9 us at M = 1880 (LEMAS_G2_BN = 128 | 256 forces one)
static int pick_bn2(const lemas_gemm_desc& d, int batches, int rows, const GemmParams& p) {
  static const int forced = [] { const char* e = getenv("LEMAS_G2_BN"); return e ? atoi(e) : 0; }();
  if (p.ln_out16 != nullptr || p.ln_stats_in != nullptr) return 256;
  if (forced == 128 || forced == 256) return forced;
  const long tiles256 = (long)batches * ((rows + 2 * G2_BM - 1) / (2 * G2_BM)) * (d.n / 256);
  const long pairs = (d.max_ctas > 0 ? d.max_ctas : sm_count()) / 2;
  if (tiles256 > 2 * pairs) return 256;   // large problems: the wider tile moves fewer operand bytes (measured at M = 4374)
  // small problems are a handful of waves: a 256 x 128 tile costs ~0.62 of a 256 x 256 one (measured, M = 1880)
  const double cost256 = (double)((tiles256 + pairs - 1) / pairs);
  const double cost128 = 0.62 * (double)((2 * tiles256 + pairs - 1) / pairs);
  return cost128 < cost256 ? 128 : 256;
}

bool gemm2_eligible(const lemas_gemm_desc& d) {
  const bool epi_ok = d.epilogue == LEMAS_EPI_BIAS_F16 || d.epilogue == LEMAS_EPI_QKV_ROPE ||
                      d.epilogue == LEMAS_EPI_GELU_TANH_F16 || d.epilogue == LEMAS_EPI_GELU_ERF_F16 ||
                      d.epilogue == LEMAS_EPI_GATE_RESID_F32 || d.epilogue == LEMAS_EPI_BIAS_F32;
  return epi_ok && (d.batches == 1 || d.seq_len == d.rows) && d.taps == 1 && d.group_cols == 0 && d.block_n == 256 && d.n % G2_BN == 0 &&
         d.a_cols == d.k_per_tap && (d.bias == nullptr || (reinterpret_cast<uintptr_t>(d.bias) & 15) == 0);
}

// Called by gemm_launch (gemm.cu) after validation, with the kernel parameters already filled in.
int gemm2_launch(const lemas_gemm_desc& d, const GemmParams& p_in, cudaStream_t st) {
  // A flat [rows, K] operand made of whole sequences (rows = batches * seq_len) is tiled per sequence, so that tiles
  // never straddle batch items: the epilogues index per-batch vectors without dividing, and sequence positions of a
  // warp's 32 rows start at a multiple of 32 (aligned 16-byte stores of the transposed V copy).
  int batches = d.batches, rows = d.rows;
  if (batches == 1 && d.seq_len > 0 && d.seq_len < rows && rows % d.seq_len == 0) {
    batches = rows / d.seq_len;
    rows = d.seq_len;
  }
  GemmParams p = p_in;
  p.batches = batches;
  p.rows = rows;
  const int bn = pick_bn2(d, batches, rows, p);
  CUtensorMap tmA, tmW;
  {
    uint64_t dims[3] = {(uint64_t)d.a_cols, (uint64_t)rows, (uint64_t)batches};
    uint64_t strides[2] = {(uint64_t)d.lda * 2, (uint64_t)rows * d.lda * 2};
    uint32_t box[3] = {G2_BK, G2_BM, 1};
    LEMAS_TRY(make_tensor_map_f16(&tmA, d.a, 3, dims, strides, box));
  }
  {
    uint64_t dims[2] = {(uint64_t)d.ldw, (uint64_t)d.w_rows};
    uint64_t strides[1] = {(uint64_t)d.ldw * 2};
    uint32_t box[2] = {G2_BK, (uint32_t)(bn / 2)};
    LEMAS_TRY(make_tensor_map_f16(&tmW, d.w, 2, dims, strides, box));
  }
  switch (d.epilogue) {
    case LEMAS_EPI_BIAS_F16: return launch2<LEMAS_EPI_BIAS_F16>(tmA, tmW, p, d.max_ctas, st, bn);
    case LEMAS_EPI_QKV_ROPE:
      return p.ln_stats_in ? launch2<LEMAS_EPI_QKV_ROPE, true>(tmA, tmW, p, d.max_ctas, st)
                           : launch2<LEMAS_EPI_QKV_ROPE>(tmA, tmW, p, d.max_ctas, st, bn);
    case LEMAS_EPI_GELU_TANH_F16:
      return p.ln_stats_in ? launch2<LEMAS_EPI_GELU_TANH_F16, true>(tmA, tmW, p, d.max_ctas, st)
                           : launch2<LEMAS_EPI_GELU_TANH_F16>(tmA, tmW, p, d.max_ctas, st, bn);
    case LEMAS_EPI_GELU_ERF_F16: return launch2<LEMAS_EPI_GELU_ERF_F16>(tmA, tmW, p, d.max_ctas, st, bn);
    case LEMAS_EPI_GATE_RESID_F32:
      return p.ln_out16 ? launch2<LEMAS_EPI_GATE_RESID_F32, true>(tmA, tmW, p, d.max_ctas, st)
                        : launch2<LEMAS_EPI_GATE_RESID_F32>(tmA, tmW, p, d.max_ctas, st, bn);
    case LEMAS_EPI_BIAS_F32: return launch2<LEMAS_EPI_BIAS_F32>(tmA, tmW, p, d.max_ctas, st, bn);
  }
  return fail(LEMAS_ERR_INVALID, "gemm2: unsupported epilogue");
}

}  // namespace lemas
