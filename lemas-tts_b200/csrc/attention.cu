// Non-causal self-attention for head_dim 64 on tcgen05 tensor cores (sm_100a), key-padding aware.
//
//   O = softmax(Q K^T / 8 + keymask) V        modules.py:483-491 (F.scaled_dot_product_attention call site)
//
// One CTA = 128 query rows of one (batch, head); two CTAs are co-resident per SM so that one CTA's softmax
// overlaps the other's MMAs.  Roles (192 threads):
//   warp 0    TMA producer : Q tile once, then K tile [128 keys x 64] and V^T tile [64 x 128 keys] per KV block
//                            into a 2-stage ring (128B swizzle, mbarrier tx-count)
//   warp 1    MMA issuer   : S = Q K^T (M128 N128 K16 x4) into TMEM cols [0,128); O_part = P V (M128 N64 K16 x8)
//                            into TMEM cols [128,192); S for block j+1 is issued while softmax j runs
//   warps 2-5 softmax      : thread == query row (tcgen05.ld 32x32b gives each lane one row): two passes over S in
//                            TMEM (row max, then exp2 / row sum), P written as fp16 into the swizzled K-major smem
//                            layout UMMA reads as the A operand; O accumulated in registers with online rescale.
// Fully masked KV blocks (keys >= kv_len) are skipped; the partial block is masked to -inf.
// K comes from the fused QKV buffer [b*seq, ld_qk]; V is read from the transposed copy [b, head, 64, vt_ld] that
// the QKV GEMM epilogue writes, so both MMAs use K-major operands.
#include "common.h"
#include "ptx.cuh"

namespace lemas {

constexpr int ATT_THREADS = 192;
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
constexpr int ATT_OFF_BAR = ATT_OFF_P + ATT_P_BYTES;
constexpr int ATT_SMEM = ATT_OFF_BAR + 128;

struct AttnParams {
  const int* kv_len;
  __half* out;
  int seq, heads, inner;
};

DEVI float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVT,
                 const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                 // [2]
  uint64_t* kv_empty = bars + 3;                // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
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
      mbar_init(kv_full + s, 1);
      mbar_init(kv_empty + s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_o = tmem_base + ATT_BN;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, ATT_Q_BYTES);
      tma_load_3d(smem, &tmQK, q_full, h * ATT_D, q0, b);
      for (int j = 0; j < n_blocks; ++j) {
        const int s = j & 1;
        mbar_wait(kv_empty + s, ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full + s, ATT_KV_STAGE);
        uint8_t* sk = smem + ATT_OFF_KV + s * ATT_KV_STAGE;
        tma_load_3d(sk, &tmQK, kv_full + s, p.inner + h * ATT_D, j * ATT_BN, b);
        tma_load_3d(sk + ATT_K_BYTES, &tmVT, kv_full + s, j * ATT_BN, 0, b * p.heads + h);
        tma_load_3d(sk + ATT_K_BYTES + ATT_V_BYTES / 2, &tmVT, kv_full + s, j * ATT_BN + 64, 0, b * p.heads + h);
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
    };
    mbar_wait(q_full, 0);
    mbar_wait(kv_full + 0, 0);
    tc_fence_after();
    if (elect_one()) issue_s(0);
    __syncwarp();
    for (int j = 0; j < n_blocks; ++j) {
      if (j + 1 < n_blocks) {
        mbar_wait(kv_full + ((j + 1) & 1), ((j + 1) >> 1) & 1);
        mbar_wait(s_empty, j & 1);  // softmax has pulled S_j out of TMEM
        tc_fence_after();
        if (elect_one()) issue_s(j + 1);
        __syncwarp();
      }
      mbar_wait(p_full, j & 1);  // P_j is in smem (and O_part_{j-1} has been consumed)
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sv = smem_u32(smem + ATT_OFF_KV + (j & 1) * ATT_KV_STAGE + ATT_K_BYTES);
#pragma unroll
        for (int ks = 0; ks < ATT_BN / 16; ++ks) {
          const uint64_t adesc = umma_desc_sw128(sp + (ks >> 2) * (ATT_P_BYTES / 2)) + 2 * (ks & 3);
          const uint64_t bdesc = umma_desc_sw128(sv + (ks >> 2) * (ATT_V_BYTES / 2)) + 2 * (ks & 3);
          umma_f16_ss(tmem_o, adesc, bdesc, idesc_o, ks != 0);
        }
        umma_commit(o_full);
        umma_commit(kv_empty + (j & 1));
      }
      __syncwarp();
    }
  } else {
    const int sub = warp & 3;
    const int r = sub * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(sub * 32) << 16;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float o[ATT_D];
#pragma unroll
    for (int i = 0; i < ATT_D; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_pending = 1.f;
    uint8_t* prow = smem + ATT_OFF_P + (r >> 3) * 1024 + (r & 7) * 128;

    for (int j = 0; j < n_blocks; ++j) {
      const int valid = kvl - j * ATT_BN;  // keys of this block that exist (>= 1)
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < ATT_BN; cc += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_s + lane_addr + cc, v);
        tmem_ld_wait();
        if (cc + 32 <= valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2f((m_run - m_new) * c);
      const float mc = m_new * c;
      m_run = m_new;
      // fold in the previous block's P V (also guarantees the P buffer is free again)
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_o + lane_addr + hh * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[hh * 32 + i] = o[hh * 32 + i] * alpha_pending + __uint_as_float(v[i]);
        }
      }
      alpha_pending = alpha;
      // pass 2: p = exp2(s*c - m*c), row sum, fp16 P into swizzled smem
      float rs = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < ATT_BN; cc += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_s + lane_addr + cc, v);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e = ex2f(fmaf(__uint_as_float(v[i]), c, -mc));
          if (cc + 32 > valid && cc + i >= valid) e = 0.f;
          pv[i] = e;
          rs += e;
        }
        uint8_t* dst = prow + (cc >> 6) * (ATT_P_BYTES / 2);
        const int u0 = (cc & 32) >> 3;  // first 16-byte unit of this chunk inside the 128-byte row
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 w;
          w.x = pack_half2(pv[8 * u + 0], pv[8 * u + 1]);
          w.y = pack_half2(pv[8 * u + 2], pv[8 * u + 3]);
          w.z = pack_half2(pv[8 * u + 4], pv[8 * u + 5]);
          w.w = pack_half2(pv[8 * u + 6], pv[8 * u + 7]);
          *reinterpret_cast<uint4*>(dst + (((u0 + u) ^ (r & 7)) << 4)) = w;
        }
      }
      l_run = l_run * alpha + rs;
      tc_fence_before();
      mbar_arrive(s_empty);        // S_j fully read: the MMA warp may overwrite it with S_{j+1}
      fence_proxy_async_smem();    // make the generic-proxy P stores visible to the tensor core (async proxy)
      mbar_arrive(p_full);
    }
    // last block's P V
    mbar_wait(o_full, (n_blocks - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_o + lane_addr + hh * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[hh * 32 + i] = o[hh * 32 + i] * alpha_pending + __uint_as_float(v[i]);
    }
    const int row = q0 + r;
    if (row < p.seq) {
      const float inv = 1.0f / l_run;
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((long)b * p.seq + row) * p.inner + h * ATT_D);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint4 w;
        w.x = pack_half2(o[8 * u + 0] * inv, o[8 * u + 1] * inv);
        w.y = pack_half2(o[8 * u + 2] * inv, o[8 * u + 3] * inv);
        w.z = pack_half2(o[8 * u + 4] * inv, o[8 * u + 5] * inv);
        w.w = pack_half2(o[8 * u + 6] * inv, o[8 * u + 7] * inv);
        dst[u] = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

}  // namespace lemas

using namespace lemas;

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
  static bool configured = false;
  if (!configured) {
    LEMAS_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    configured = true;
  }
  AttnParams p;
  p.kv_len = kv_len;
  p.out = static_cast<__half*>(out16);
  p.seq = seq;
  p.heads = heads;
  p.inner = inner;
  dim3 grid((seq + ATT_BM - 1) / ATT_BM, heads, batch);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, (cudaStream_t)stream>>>(tmQK, tmVT, p);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}
