// Helpers shared by the attention kernels (attention.cu: v3, one 128-query tile per CTA, two CTAs per SM;
// attention7.cu: v7, persistent CTA with double-buffered scores): packed fp32 arithmetic, MUFU, the TS-form MMA and the
// hang-hunting waits.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace lemas {

struct AttnParams {
  long long* trace;   // debug: clock64 stamps (tools/trace_att.py); nullptr in production
  const int* kv_len;
  __half* out;
  int seq, heads, inner;
  int debug;          // timing experiments (LEMAS_A7_DEBUG): skip parts of the MMA issuers' work
  int dephase_half, dephase_tile;   // v7 only: start-up stagger of the key parts in clocks
  int n_pairs, n_items;   // v7 only: 128-query tiles per (batch, head); work items = n_pairs * heads * batch
};

int attention_v9_launch(const void* qk, int32_t ld_qk, const void* vt, int32_t vt_ld, const AttnParams& p, int batch,
                        void* stream);
int attention_v8_launch(const CUtensorMap& tmQK, const CUtensorMap& tmVT, const AttnParams& p, void* stream);
int attention_v7_launch(const CUtensorMap& tmQK, const CUtensorMap& tmVT, const AttnParams& p, int poly, void* stream);
int attention_v3_launch(const void* qk, int32_t ld_qk, const void* vt, int32_t vt_ld, const int32_t* kv_len,
                        void* out16, int32_t batch, int32_t seq, int32_t heads, long long* trace, void* stream);

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
// Packed fp32 pairs (sm_100 FFMA2 / FADD2): one issue slot for two lanes of arithmetic.
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
// D[tmem] (+)= A[tmem] * B[smem]^T : A = 128 lanes x (K/2) columns of packed fp16 pairs (row-major along K).
DEVI void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

#ifdef LEMAS_ATT_DEBUG  // hang hunting: every wait has a deadline; a waiter that misses it records who it is in
                        // (pinned host) memory at p.trace and traps.  tools/att_hang_probe.py --debug
DEVI void dbg_wait(long long* dbg, uint32_t bar_addr, uint32_t parity, int tag, int j) {
  bool done = false;
  for (int outer = 0; outer < 200000 && !done; ++outer) {   // same tight polling as the production waits
#pragma unroll 1
    for (int inner = 0; inner < 64; ++inner) {
      uint32_t ok;
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
      if (ok) { done = true; break; }
    }
  }
  if (done) return;
  if ((threadIdx.x & 31) == 0 || tag < 8) {
    const int slot = atomicAdd(reinterpret_cast<int*>(dbg), 1);
    if (slot < 500) {
      long long* r = dbg + 1 + slot * 4;
      unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      r[0] = (long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      r[1] = threadIdx.x >> 5;
      r[2] = tag * 1000 + j;
      r[3] = smid;
    }
    __threadfence_system();
  }
  const long long t2 = clock64() + 4000000ll;
  while (clock64() < t2) { }
  __trap();
}
#define ATT_WAIT_P(barptr, parity, tag, j) dbg_wait(p.trace, smem_u32(barptr), parity, tag, j)
#define ATT_WAIT_A(addr, parity, tag, j) dbg_wait(p.trace, addr, parity, tag, j)
#else
#define ATT_WAIT_P(barptr, parity, tag, j) mbar_wait(barptr, parity)
#define ATT_WAIT_A(addr, parity, tag, j) mbar_wait_lean(addr, parity)
#endif


}  // namespace lemas
