// Host-side driver of the hot path: CFM.sample's ODE loop (cfm.py:382-456) over the co-batched cond/uncond
// DiT forward (dit.py:194-254), and Vocos.decode.  Pure launch sequencing — every FLOP is in the kernels of
// gemm.cu / attention.cu / elementwise.cu.  No allocation: the caller's workspace is carved up here.
#include <cstdlib>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; ranges are emitted only with LEMAS_NVTX=1 (nsys / ncu --nvtx)

#include "common.h"
#include "ptx.cuh"

namespace lemas {

int gemm_launch(const lemas_gemm_desc& d, cudaStream_t stream);
int step_begin_launch(const float* mod_table, long mod_w, float* mod_cur, const float* t_grid, int* step_ctr,
                      float* state, float cfg_strength, int* row_limit, const int* kv_len2, int n_seq, int seq, int steps,
                      cudaStream_t st);
int cfg_euler_dev_launch(const float* pred, int ld_pred, float* y, void* x16, int ld_x16, int copies, float* traj,
                         long traj_stride, int rows, int mel, const float* state, int use_cfg, cudaStream_t st);
int cfg_split_exchange_launch(const float* local, float* peer, long slot_floats, int variant, const float* state,
                              const int* flags_local, int* flags_peer, cudaStream_t st);
int ln_fold_pack_launch(const float* mod, long mod_w, int steps, int depth, int dim, void* out16, cudaStream_t st);

struct Carver {
  uint8_t* base;
  int64_t off = 0;
  explicit Carver(void* p) : base(static_cast<uint8_t*>(p)) {}
  template <class T>
  T* take(int64_t count) {
    off = align_up(off, 1024);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * (int64_t)sizeof(T);
    return p;
  }
};

struct DitBuffers {
  __half *x16, *ct16, *h0_16, *c1_16, *a16, *o16, *qk16, *vt16, *ff16;
  float *inv_embed, *h0, *x, *pred, *t_dev, *sinus, *t1, *temb, *mod, *mod_cur, *state, *y_state;
  int *kv_len2, *step_ctr, *row_limit;
  __half* ln_a16;     // folded LayerNorm: [depth][2 norms][2 steps][dim] fp16 (1 + scale | shift rows)
  float *ln_uv_qkv, *ln_uv_ff1, *ln_stats;   // [depth][2 steps][3 inner] / [depth][2 steps][F] / [M2][8][2]
  int npad;
  int64_t bytes;
};

static DitBuffers carve(const lemas_dit_config& c, int batch, int seq, int steps, int ct_ld, void* ws) {
  DitBuffers b;
  Carver cv(ws);
  const int64_t M2 = 2LL * batch * seq;
  const int D = c.dim, inner = c.heads * 64, F = c.dim * c.ff_mult;
  b.npad = (int)align_up(seq, 64);
  b.x16 = cv.take<__half>(M2 * 128);
  b.ct16 = cv.take<__half>(M2 * ct_ld);
  b.inv_embed = cv.take<float>(M2 * D);
  b.h0 = cv.take<float>(M2 * D);
  b.h0_16 = cv.take<__half>(M2 * D);
  b.c1_16 = cv.take<__half>(M2 * D);
  b.x = cv.take<float>(M2 * D);
  b.a16 = cv.take<__half>(M2 * D);
  b.o16 = cv.take<__half>(M2 * inner);
  b.qk16 = cv.take<__half>(M2 * 2 * inner);
  b.vt16 = cv.take<__half>(2LL * batch * inner * b.npad);
  b.ff16 = cv.take<__half>(M2 * F);
  b.pred = cv.take<float>(M2 * 128);
  b.t_dev = cv.take<float>(steps + 1);
  b.sinus = cv.take<float>((int64_t)steps * 256);
  b.t1 = cv.take<float>((int64_t)steps * D);
  b.temb = cv.take<float>((int64_t)steps * D);
  b.mod = cv.take<float>((int64_t)steps * ((int64_t)c.depth * 6 * D + 2 * D));
  b.kv_len2 = cv.take<int>(2 * batch);
  b.row_limit = cv.take<int>(2 * batch);
  b.mod_cur = cv.take<float>((int64_t)c.depth * 6 * D + 2 * D);
  b.state = cv.take<float>(4);
  b.step_ctr = cv.take<int>(1);
  b.y_state = cv.take<float>((int64_t)batch * seq * c.mel_dim);
  b.ln_a16 = cv.take<__half>((int64_t)c.depth * 4 * steps * D);
  b.ln_uv_qkv = cv.take<float>((int64_t)c.depth * 2 * steps * 3 * inner);
  b.ln_uv_ff1 = cv.take<float>((int64_t)c.depth * 2 * steps * F);
  b.ln_stats = cv.take<float>(M2 * 16);
  b.bytes = align_up(cv.off, 1024);
  return b;
}

}  // namespace lemas

using namespace lemas;

struct ProfRecord { int kind; cudaEvent_t e0, e1; };

struct lemas_engine {
  lemas_dit_config cfg;
  lemas_dit_weights w;
  std::vector<lemas_dit_layer> layers;
  int profile = 0;                     // 0 off, 1 events around eager launches, 2 events INSIDE the replayed step graph
  bool nvtx = false;                   // LEMAS_NVTX=1: an NVTX range per launch, named after its SURVEY §2.3 K-row
  double acc_ms[LEMAS_PROF_KINDS] = {0};   // graph-profile mode: accumulated per kind
  int64_t acc_n[LEMAS_PROF_KINDS] = {0};
  std::vector<ProfRecord> records;     // event pairs in flight since the last profile_read
  std::vector<cudaEvent_t> free_events;
  // One captured ODE step per (shape, workspace) — replayed `steps` times; everything step-dependent is read from
  // device memory (step_begin_kernel), so the same executable graph serves every step and every later call.
  struct StepGraph {
    int batch, seq, steps, variants, has_kv, flags;   // steps: the workspace carve-up (hence every captured pointer) depends on it
    float cfg;
    const void *ws, *rope, *traj, *split;
    cudaGraphExec_t exec;
    int nodes;
  };
  std::vector<StepGraph> graphs;
  int split_epoch = 0;                 // two-GPU CFG split: calls so far (both processes count alike)
  cudaStream_t cap_stream = nullptr;   // capture happens on a private stream: the caller's may be the legacy default
                                       // stream, which cannot be captured; replays go to the caller's stream
};

// Brackets one launch with events when profiling is on (lemas_engine_profile); otherwise free.
static const char* const kProfNames[LEMAS_PROF_KINDS] = {
    "K1/K12 preloop (time MLP, AdaLN table, invariant input projection)", "K9 in_proj", "K10 conv_pos + Mish",
    "K2 LN + modulate", "K3/K5 QKV GEMM + RoPE", "K6 attention", "K7 to_out GEMM + gate + residual",
    "K8 FF1 GEMM + GELU", "K8 FF2 GEMM + gate + residual", "K13 proj_out", "K13-K15 CFG + clamp + Euler",
    "empty event pair (measurement tare)"};

struct ProfScope {
  lemas_engine* e; cudaStream_t st; ProfRecord r; bool on; bool range; unsigned ext = 0;
  ProfScope(const lemas_engine* ce, int kind, cudaStream_t s)
      : e(const_cast<lemas_engine*>(ce)), st(s), on(ce->profile != 0), range(ce->nvtx) {
    if (range) nvtxRangePushA(kProfNames[kind]);
    if (!on) return;
    auto get = [&]() { cudaEvent_t ev; if (!e->free_events.empty()) { ev = e->free_events.back(); e->free_events.pop_back(); }
                       else cudaEventCreate(&ev); return ev; };
    r.kind = kind; r.e0 = get(); r.e1 = get();
    // inside a stream capture a plain record is only an internal dependency; an EXTERNAL record becomes an event-record
    // node whose timestamp can be read after the graph has run (graph-profile mode)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    ext = cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault;
    cudaEventRecordWithFlags(r.e0, st, ext);
  }
  ~ProfScope() {
    if (on) { cudaEventRecordWithFlags(r.e1, st, ext); e->records.push_back(r); }
    if (range) nvtxRangePop();
  }
};
#define PROF(kind) ProfScope _prof_scope_##__LINE__(e, kind, st)

static int ct_ld_for(const lemas_dit_config* c) { return (int)align_up(c->mel_dim + c->text_dim, 64); }

extern "C" {

int64_t lemas_engine_workspace_bytes(const lemas_dit_config* cfg, int32_t batch, int32_t seq, int32_t steps) {
  if (!cfg) return -1;
  return carve(*cfg, batch, seq, steps, ct_ld_for(cfg), nullptr).bytes;
}

int lemas_engine_create(const lemas_dit_config* cfg, const lemas_dit_weights* w, lemas_engine** out) {
  LEMAS_REQUIRE(cfg && w && out, "lemas_engine_create: null argument");
  LEMAS_REQUIRE(cfg->dim % 128 == 0 && cfg->dim <= 1024, "lemas_engine_create: dim must be a multiple of 128, <= 1024");
  LEMAS_REQUIRE(cfg->heads >= 1 && cfg->mel_dim <= 128 && cfg->mel_dim >= 1, "lemas_engine_create: bad heads/mel_dim");
  LEMAS_REQUIRE((cfg->heads * 64) % 64 == 0 && cfg->text_dim >= 1, "lemas_engine_create: bad text_dim");
  LEMAS_REQUIRE(w->ct_ld == ct_ld_for(cfg), "lemas_engine_create: ct_ld must be round_up(mel_dim + text_dim, 64)");
  LEMAS_REQUIRE(w->conv_dense == 1 || cfg->dim / 16 == 64, "lemas_engine_create: grouped conv path needs dim == 1024");
  LEMAS_REQUIRE(w->layers != nullptr, "lemas_engine_create: layers missing");
  if (!lemas_device_supported())
    return fail(LEMAS_ERR_UNSUPPORTED,
                "CUDA error: no kernel image is available for execution on the device (liblemas_b200 is sm_100a only)");
  auto* e = new lemas_engine;
  e->cfg = *cfg;
  e->w = *w;
  e->layers.assign(w->layers, w->layers + cfg->depth);
  e->w.layers = e->layers.data();
  const char* nv = getenv("LEMAS_NVTX");
  e->nvtx = nv && nv[0] == '1';
  *out = e;
  return LEMAS_OK;
}

void lemas_engine_destroy(lemas_engine* e) {
  if (!e) return;
  for (auto& r : e->records) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto ev : e->free_events) cudaEventDestroy(ev);
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  delete e;
}

int lemas_engine_profile(lemas_engine* e, int32_t enable) {
  LEMAS_REQUIRE(e, "lemas_engine_profile: null engine");
  e->profile = enable;
  return LEMAS_OK;
}

int lemas_engine_profile_read(lemas_engine* e, double* ms, int64_t* launches, void* stream) {
  LEMAS_REQUIRE(e && ms && launches, "lemas_engine_profile_read: null argument");
  LEMAS_CUDA_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  for (auto& r : e->records) {
    float t = 0.f;
    LEMAS_CUDA_OK(cudaEventElapsedTime(&t, r.e0, r.e1));
    if (r.kind >= 0 && r.kind < LEMAS_PROF_KINDS) { ms[r.kind] += t; launches[r.kind] += 1; }
    e->free_events.push_back(r.e0);
    e->free_events.push_back(r.e1);
  }
  e->records.clear();
  for (int k = 0; k < LEMAS_PROF_KINDS; ++k) {   // graph-profile mode accumulates here (lemas_sampler_run)
    ms[k] += e->acc_ms[k]; launches[k] += e->acc_n[k];
    e->acc_ms[k] = 0; e->acc_n[k] = 0;
  }
  return LEMAS_OK;
}
}

namespace lemas {

static lemas_gemm_desc base_desc(const void* a, int batches, int rows, int lda, const void* w, int w_rows, int ldw,
                                 int n, int seq_len, int epilogue, int block_n) {
  lemas_gemm_desc d = {};
  d.a = a; d.batches = batches; d.rows = rows; d.lda = lda; d.a_cols = lda;
  d.w = w; d.w_rows = w_rows; d.ldw = ldw; d.n = n;
  d.k_per_tap = ldw; d.taps = 1; d.tap_pad = 0; d.w_tap_stride = 0; d.group_cols = 0;
  d.block_n = block_n; d.epilogue = epilogue; d.seq_len = seq_len;
  return d;
}

// One DiT forward on `variants*batch` co-batched sequences; mod = this step's modulation row.
// fold_steps > 0: LayerNorm of every block except layer 0's attn_norm and the final norm is folded into the GEMMs
// around it (lemas_gemm_desc.ln_*); the u / v tables for `fold_steps` ODE steps were built by prepare().
static int dit_forward(const lemas_engine* e, const DitBuffers& b, int batch, int seq, int variants, const float* mod,
                       const int* kv_len2, const float* rope, float* pred, const int* row_limit, int fold_steps,
                       cudaStream_t st) {
  const lemas_dit_config& c = e->cfg;
  const lemas_dit_weights& w = e->w;
  const int D = c.dim, inner = c.heads * 64, F = c.dim * c.ff_mult;
  const int B2 = variants * batch;
  const int M = B2 * seq;
  const int bn_d = D % 256 == 0 ? 256 : 128;

  {  // dit.py:97  x-columns of the input projection + the precomputed (cond|text) part
    lemas_gemm_desc d = base_desc(b.x16, 1, M, 128, w.w_in_x, D, 128, D, seq, LEMAS_EPI_ADD_F32_F16, bn_d);
    d.resid = b.inv_embed; d.ldr = D; d.out32 = b.h0; d.ld32 = D; d.out16 = b.h0_16; d.ld16 = D;
    PROF(LEMAS_PROF_IN_PROJ);
    LEMAS_TRY(gemm_launch(d, st));
  }
  for (int j = 0; j < 2; ++j) {  // modules.py:171-176 grouped conv k=31 + Mish, twice; dit.py:98 residual
    const __half* in = j == 0 ? b.h0_16 : b.c1_16;
    lemas_gemm_desc d = {};
    d.a = in; d.batches = B2; d.rows = seq; d.lda = D; d.a_cols = D;
    d.w = w.conv_w[j]; d.w_rows = 31 * D; d.n = D; d.taps = 31; d.tap_pad = 15; d.w_tap_stride = D;
    if (w.conv_dense) { d.ldw = D; d.k_per_tap = D; d.group_cols = 0; d.block_n = bn_d; }
    else { d.ldw = 64; d.k_per_tap = 64; d.group_cols = 64; d.block_n = 64; }
    d.bias = w.conv_b[j]; d.seq_len = seq;
    if (j == 0) { d.epilogue = LEMAS_EPI_MISH_F16; d.out16 = b.c1_16; d.ld16 = D; }
    else { d.epilogue = LEMAS_EPI_MISH_RESID_F32; d.resid = b.h0; d.ldr = D; d.out32 = b.x; d.ld32 = D; }
    PROF(LEMAS_PROF_CONV_POS);
    LEMAS_TRY(gemm_launch(d, st));
  }
  for (int l = 0; l < c.depth; ++l) {
    const lemas_dit_layer& L = w.layers[l];
    const float* m = mod + (int64_t)l * 6 * D;  // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
    const bool fold = fold_steps > 0;
    const int parts = 2 * (D / 256);   // 128-column partials per row written by the gated-residual epilogues
    if (!fold || l == 0) {
      PROF(LEMAS_PROF_LN_MOD);
      LEMAS_TRY(lemas_ln_modulate_rows(b.x, m + D, m, 0, b.a16, M, D, seq, row_limit, st));
    }
    {
      lemas_gemm_desc d = base_desc(b.a16, 1, M, D, L.w_qkv, 3 * inner, D, 3 * inner, seq, LEMAS_EPI_QKV_ROPE, 256);
      if (fold && l > 0) {  // A = x (1 + scale_msa) written by the previous layer's FF2 epilogue
        d.ln_stats_in = b.ln_stats; d.ln_parts = parts; d.ln_k = D; d.ln_step = b.step_ctr;
        d.ln_uv = b.ln_uv_qkv + (int64_t)l * 2 * fold_steps * 3 * inner;
      }
      d.bias = L.b_qkv; d.out16 = b.qk16; d.ld16 = 2 * inner; d.rope = rope;
      d.rope_cols = c.rope_heads * 64; d.inner = inner; d.vt = b.vt16; d.vt_ld = b.npad;
      d.row_limit = row_limit;
      PROF(LEMAS_PROF_GEMM_QKV);
      LEMAS_TRY(gemm_launch(d, st));
    }
    {
      PROF(LEMAS_PROF_ATTENTION);
      LEMAS_TRY(lemas_attention_f16(b.qk16, 2 * inner, b.vt16, b.npad, kv_len2, b.o16, B2, seq, c.heads, st));
    }
    {
      lemas_gemm_desc d = base_desc(b.o16, 1, M, inner, L.w_out, D, inner, D, seq, LEMAS_EPI_GATE_RESID_F32, bn_d);
      d.bias = L.b_out; d.resid = b.x; d.ldr = D; d.out32 = b.x; d.ld32 = D; d.gate = m + 2 * D; d.gate_bstride = 0;
      d.row_valid = kv_len2;
      d.row_limit = row_limit;
      if (fold) { d.ln_scale = m + 4 * D; d.ln_out16 = b.a16; d.ln_ld16 = D; d.ln_stats = b.ln_stats; }  // ff_norm
      PROF(LEMAS_PROF_GEMM_OUT);
      LEMAS_TRY(gemm_launch(d, st));
    }
    if (!fold) {
      PROF(LEMAS_PROF_LN_MOD);
      LEMAS_TRY(lemas_ln_modulate_rows(b.x, m + 4 * D, m + 3 * D, 0, b.a16, M, D, seq, row_limit, st));
    }
    {
      lemas_gemm_desc d = base_desc(b.a16, 1, M, D, L.w_ff1, F, D, F, seq, LEMAS_EPI_GELU_TANH_F16, 256);
      d.bias = L.b_ff1; d.out16 = b.ff16; d.ld16 = F;
      d.row_limit = row_limit;
      if (fold) {
        d.ln_stats_in = b.ln_stats; d.ln_parts = parts; d.ln_k = D; d.ln_step = b.step_ctr;
        d.ln_uv = b.ln_uv_ff1 + (int64_t)l * 2 * fold_steps * F;
      }
      PROF(LEMAS_PROF_GEMM_FF1);
      LEMAS_TRY(gemm_launch(d, st));
    }
    {
      lemas_gemm_desc d = base_desc(b.ff16, 1, M, F, L.w_ff2, D, F, D, seq, LEMAS_EPI_GATE_RESID_F32, bn_d);
      d.bias = L.b_ff2; d.resid = b.x; d.ldr = D; d.out32 = b.x; d.ld32 = D; d.gate = m + 5 * D; d.gate_bstride = 0;
      d.row_limit = row_limit;
      if (fold && l + 1 < c.depth) {  // attn_norm of the next layer: its scale_msa sits 6 D further in the row
        d.ln_scale = m + 6 * D + D; d.ln_out16 = b.a16; d.ln_ld16 = D; d.ln_stats = b.ln_stats;
      }
      PROF(LEMAS_PROF_GEMM_FF2);
      LEMAS_TRY(gemm_launch(d, st));
    }
  }
  const float* mf = mod + (int64_t)c.depth * 6 * D;  // modules.py:333: (scale, shift)
  { PROF(LEMAS_PROF_LN_MOD); LEMAS_TRY(lemas_ln_modulate_rows(b.x, mf, mf + D, 0, b.a16, M, D, seq, row_limit, st)); }
  {
    lemas_gemm_desc d = base_desc(b.a16, 1, M, D, w.w_proj, 128, D, c.mel_dim, seq, LEMAS_EPI_BIAS_F32, 128);
    d.bias = w.b_proj; d.out32 = pred; d.ld32 = 128;
    PROF(LEMAS_PROF_PROJ_OUT);
    LEMAS_TRY(gemm_launch(d, st));
  }
  return LEMAS_OK;
}

// Everything that does not depend on the ODE state: time embeddings + all AdaLN modulations for every step
// (modules.py:311,332,725 hoisted), the (cond|text) half of the input projection, fp16 copy of y0.
// LEMAS_SAMPLE_FOLD_LAYERNORM: fold the LayerNorms into the GEMMs around them (off by default: measured 1.2 % slower
// at C2 than the stand-alone kernel).  Needs the CTA-pair GEMM on every projection it touches.
static bool fold_ok(const lemas_engine* e, const lemas_sample_args* a) {
  const lemas_dit_config& c = e->cfg;
  return (a->flags & LEMAS_SAMPLE_FOLD_LAYERNORM) != 0 && c.dim % 256 == 0 && (c.dim * c.ff_mult) % 256 == 0 &&
         (3 * c.heads * 64) % 256 == 0;
}

static int prepare(const lemas_engine* e, const DitBuffers& b, const lemas_sample_args* a, int variants, int n_times,
                   const float* times_host, int fold_steps, cudaStream_t st) {
  const lemas_dit_config& c = e->cfg;
  const lemas_dit_weights& w = e->w;
  const int D = c.dim;
  const int rows = a->batch * a->seq;
  const int64_t mod_w = (int64_t)c.depth * 6 * D + 2 * D;
  PROF(LEMAS_PROF_PRELOOP);
  LEMAS_CUDA_OK(cudaMemcpyAsync(b.t_dev, times_host, sizeof(float) * n_times, cudaMemcpyHostToDevice, st));
  LEMAS_TRY(lemas_time_sinusoid(b.t_dev, b.sinus, n_times, st));
  LEMAS_TRY(lemas_skinny_linear_f32(b.sinus, w.time_w0, w.time_b0, b.t1, n_times, 256, D, 0, 1, st));
  LEMAS_TRY(lemas_skinny_linear_f32(b.t1, w.time_w2, w.time_b2, b.temb, n_times, D, D, 0, 0, st));
  LEMAS_TRY(lemas_skinny_linear_f32(b.temb, w.adaln_w, w.adaln_b, b.mod, n_times, D, (int)mod_w, 1, 0, st));
  if (fold_steps > 0) {
    // u[n] = sum_k W[n,k] (1 + scale_k), v[n] = sum_k W[n,k] shift_k for every step, layer and folded norm: one small
    // GEMM per (layer, norm) over the fp16 weights the main GEMMs use (so the mean term cancels exactly)
    const int inner = c.heads * 64, F = c.dim * c.ff_mult, R = 2 * fold_steps;
    LEMAS_TRY(ln_fold_pack_launch(b.mod, mod_w, fold_steps, c.depth, D, b.ln_a16, st));
    for (int l = 0; l < c.depth; ++l)
      for (int wsel = 0; wsel < 2; ++wsel) {
        const int N = wsel == 0 ? 3 * inner : F;
        lemas_gemm_desc d = base_desc(b.ln_a16 + ((int64_t)(l * 2 + wsel) * R) * D, 1, R, D,
                                      wsel == 0 ? w.layers[l].w_qkv : w.layers[l].w_ff1, N, D, N, R, LEMAS_EPI_BIAS_F32, 256);
        d.out32 = (wsel == 0 ? b.ln_uv_qkv + (int64_t)l * R * N : b.ln_uv_ff1 + (int64_t)l * R * N);
        d.ld32 = N;
        LEMAS_TRY(gemm_launch(d, st));
      }
  }
  LEMAS_TRY(lemas_cast_pad_f16(a->y, b.x16, rows, c.mel_dim, 128, variants, st));
  LEMAS_TRY(lemas_pack_cond_text(a->step_cond, a->text_cond, a->text_uncond, b.ct16, rows, c.mel_dim, c.text_dim,
                                 w.ct_ld, variants, st));
  {
    const int bn_d = D % 256 == 0 ? 256 : 128;
    lemas_gemm_desc d = base_desc(b.ct16, 1, variants * rows, w.ct_ld, w.w_in_ct, D, w.ct_ld, D, a->seq,
                                  LEMAS_EPI_BIAS_F32, bn_d);
    d.bias = w.b_in; d.out32 = b.inv_embed; d.ld32 = D;
    LEMAS_TRY(gemm_launch(d, st));
  }
  if (a->kv_len) {
    for (int v = 0; v < variants; ++v)
      LEMAS_CUDA_OK(cudaMemcpyAsync(b.kv_len2 + v * a->batch, a->kv_len, sizeof(int) * a->batch,
                                    cudaMemcpyDeviceToDevice, st));
  }
  return LEMAS_OK;
}

static int check_args(const lemas_engine* e, const lemas_sample_args* a, int steps_for_ws, DitBuffers* out) {
  LEMAS_REQUIRE(e && a, "lemas sampler: null argument");
  LEMAS_REQUIRE(a->batch >= 1 && a->seq >= 1 && a->seq <= 4096, "lemas sampler: need 1 <= seq <= 4096, batch >= 1");
  LEMAS_REQUIRE(a->y && a->step_cond && a->text_cond && a->rope && a->workspace, "lemas sampler: null tensor");
  *out = carve(e->cfg, a->batch, a->seq, steps_for_ws, e->w.ct_ld, a->workspace);
  LEMAS_REQUIRE(a->workspace_bytes >= out->bytes, "lemas sampler: workspace too small");
  LEMAS_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 1023) == 0, "lemas sampler: workspace must be 1 KiB aligned");
  return LEMAS_OK;
}

}  // namespace lemas

extern "C" {

// One ODE step with every step-dependent quantity read from device memory: modulation row -> mod_cur, (cfg_t, dt,
// step) -> state.  Identical launch arguments for every step, so it can be captured once and replayed.
static int ode_step(lemas_engine* e, const DitBuffers& b, const lemas_sample_args* a, int variants, const int* kv2,
                    float* y, cudaStream_t st) {
  const lemas_dit_config& c = e->cfg;
  int* row_limit = (kv2 != nullptr && (a->flags & LEMAS_SAMPLE_SKIP_PADDED_ROWS)) ? b.row_limit : nullptr;
  const int rows = a->batch * a->seq;
  const int64_t mod_w = (int64_t)c.depth * 6 * c.dim + 2 * c.dim;
  { PROF(LEMAS_PROF_TARE); }   // nothing between the two records: what an event pair itself adds to every bracket
  {
    PROF(LEMAS_PROF_CFG_EULER);
    LEMAS_TRY(step_begin_launch(b.mod, mod_w, b.mod_cur, b.t_dev, b.step_ctr, b.state,
                                (variants == 2 || a->split_xchg_local) ? a->cfg_strength : 0.f, row_limit, kv2,
                                variants * a->batch, a->seq,
                                a->steps, st));
  }
  const bool split = a->split_xchg_local != nullptr;
  float* pred = split ? a->split_xchg_local + (int64_t)a->split_variant * rows * 128 : b.pred;
  LEMAS_TRY(dit_forward(e, b, a->batch, a->seq, variants, b.mod_cur, kv2, a->rope, pred, row_limit,
                        fold_ok(e, a) ? a->steps : 0, st));
  PROF(LEMAS_PROF_CFG_EULER);
  if (split)  // the other variant's pred arrives from the peer GPU; afterwards both slots are complete on both sides
    LEMAS_TRY(cfg_split_exchange_launch(a->split_xchg_local, a->split_xchg_peer, (long)rows * 128, a->split_variant,
                                        b.state, a->split_flags_local, a->split_flags_peer, st));
  LEMAS_TRY(cfg_euler_dev_launch(split ? a->split_xchg_local : b.pred, 128, y, b.x16, 128, variants, a->trajectory,
                                 (long)rows * c.mel_dim, rows, c.mel_dim, b.state, (variants == 2 || split) ? 1 : 0, st));
  return LEMAS_OK;
}

int lemas_sampler_run(lemas_engine* e, const lemas_sample_args* a, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DitBuffers b;
  LEMAS_TRY(check_args(e, a, a ? a->steps : 0, &b));
  LEMAS_REQUIRE(a->steps >= 1 && a->t_grid_host, "lemas_sampler_run: steps >= 1 and a t grid are required");
  const bool split = a->split_xchg_local != nullptr;
  if (split)
    LEMAS_REQUIRE(a->cfg_strength >= 1e-5f && a->split_xchg_peer && a->split_flags_local && a->split_flags_peer &&
                      (a->split_variant == 0 || a->split_variant == 1),
                  "lemas_sampler_run: the two-GPU CFG split needs cfg_strength > 0, both exchange buffers and both flag "
                  "arrays, and split_variant 0 or 1");
  const int variants = (a->cfg_strength >= 1e-5f && !split) ? 2 : 1;   // split: ONE variant here, the other on the peer
  LEMAS_REQUIRE(variants == 1 || a->text_uncond, "lemas_sampler_run: text_uncond required when cfg_strength > 0");
  const lemas_dit_config& c = e->cfg;
  const int rows = a->batch * a->seq;
  LEMAS_TRY(prepare(e, b, a, variants, a->steps, a->t_grid_host, fold_ok(e, a) ? a->steps : 0, st));
  // prepare() uploaded t[0..steps-1]; the step kernels also need t[steps] for the last dt
  LEMAS_CUDA_OK(cudaMemcpyAsync(b.t_dev + a->steps, a->t_grid_host + a->steps, sizeof(float), cudaMemcpyHostToDevice, st));
  LEMAS_CUDA_OK(cudaMemsetAsync(b.step_ctr, 0, sizeof(int), st));
  if (split) {  // state[3] = call epoch (the exchange kernel's flag targets grow monotonically across calls)
    e->split_epoch += 1;
    LEMAS_CUDA_OK(cudaMemcpyAsync(b.state + 3, &e->split_epoch, sizeof(int), cudaMemcpyHostToDevice, st));
  }
  const int64_t state = (int64_t)rows * c.mel_dim;
  if (a->trajectory)
    LEMAS_CUDA_OK(cudaMemcpyAsync(a->trajectory, a->y, sizeof(float) * state, cudaMemcpyDeviceToDevice, st));
  const int* kv2 = a->kv_len ? b.kv_len2 : nullptr;

  // the trajectory slot is indexed by the device-side step counter, so a requested trajectory replays too; its
  // pointer is captured, hence part of the cache key (the Python side passes a persistent staging buffer)
  const bool want_graph = a->use_graph && e->profile != 1 && a->steps >= 3;
  if (want_graph && e->profile == 2) {
    // Graph-profile mode: the event pairs of the PROF scopes are recorded INSIDE a captured step (event-record
    // nodes), the instrumented graph is replayed step by step and read back after every replay — per-kernel times of
    // the graph-replayed step (no host launch gaps), which is what the timed bench region runs.
    LEMAS_CUDA_OK(cudaMemcpyAsync(b.y_state, a->y, sizeof(float) * state, cudaMemcpyDeviceToDevice, st));
    std::vector<ProfRecord> eager;   // the pre-loop record made by prepare() stays an ordinary (eager) record
    eager.swap(e->records);
    e->profile = 0;
    int rc = ode_step(e, b, a, variants, kv2, b.y_state, st);   // step 0 eagerly (one-time kernel attribute set-up)
    e->profile = 2;
    LEMAS_TRY(rc);
    cudaGraph_t graph = nullptr;
    if (!e->cap_stream) LEMAS_CUDA_OK(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    const int64_t before = launches_so_far();
    LEMAS_CUDA_OK(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    rc = ode_step(e, b, a, variants, kv2, b.y_state, e->cap_stream);
    const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
    const int nodes = (int)(launches_so_far() - before);
    count_launches(-nodes);
    std::vector<ProfRecord> grec;
    grec.swap(e->records);
    e->records.swap(eager);
    if (rc != LEMAS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    LEMAS_CUDA_OK(ce);
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    LEMAS_CUDA_OK(ie);
    for (int i = 1; i < a->steps; ++i) {
      if (e->nvtx) nvtxRangePushA("ode_step (graph replay)");
      cudaError_t le = cudaGraphLaunch(exec, st);
      if (le == cudaSuccess) le = cudaStreamSynchronize(st);
      if (e->nvtx) nvtxRangePop();
      if (le != cudaSuccess) { cudaGraphExecDestroy(exec); LEMAS_CUDA_OK(le); }
      count_launches(nodes);
      for (auto& r : grec) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess && r.kind >= 0 && r.kind < LEMAS_PROF_KINDS) {
          e->acc_ms[r.kind] += t;
          e->acc_n[r.kind] += 1;
        } else {
          (void)cudaGetLastError();   // never leave a sticky error behind a measurement aid
        }
      }
    }
    cudaGraphExecDestroy(exec);
    for (auto& r : grec) { e->free_events.push_back(r.e0); e->free_events.push_back(r.e1); }
    LEMAS_CUDA_OK(cudaMemcpyAsync(a->y, b.y_state, sizeof(float) * state, cudaMemcpyDeviceToDevice, st));
    return LEMAS_OK;
  }
  if (!want_graph) {
    for (int i = 0; i < a->steps; ++i) LEMAS_TRY(ode_step(e, b, a, variants, kv2, a->y, st));
    return LEMAS_OK;
  }

  // graph path: the ODE state lives in the workspace so that the captured pointers stay valid across calls
  LEMAS_CUDA_OK(cudaMemcpyAsync(b.y_state, a->y, sizeof(float) * state, cudaMemcpyDeviceToDevice, st));
  lemas_engine::StepGraph* g = nullptr;
  for (auto& cand : e->graphs)
    if (cand.batch == a->batch && cand.seq == a->seq && cand.steps == a->steps && cand.variants == variants &&
        cand.has_kv == (kv2 != nullptr) && cand.flags == a->flags &&
        cand.cfg == a->cfg_strength && cand.ws == a->workspace && cand.rope == a->rope &&
        cand.traj == a->trajectory && cand.split == a->split_xchg_peer)
      g = &cand;
  int first_replayed = 0;
  if (!g) {
    // step 0 runs eagerly (it also performs every one-time kernel attribute set-up), then the step is captured
    LEMAS_TRY(ode_step(e, b, a, variants, kv2, b.y_state, st));
    first_replayed = 1;
    const int64_t before = launches_so_far();
    cudaGraph_t graph = nullptr;
    if (!e->cap_stream) LEMAS_CUDA_OK(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    LEMAS_CUDA_OK(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = ode_step(e, b, a, variants, kv2, b.y_state, e->cap_stream);
    const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
    const int nodes = (int)(launches_so_far() - before);
    count_launches(-nodes);  // capturing launched nothing
    if (rc != LEMAS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    LEMAS_CUDA_OK(ce);
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    LEMAS_CUDA_OK(ie);
    if (e->graphs.size() >= 8) {  // bounded cache: drop the oldest entry
      cudaGraphExecDestroy(e->graphs.front().exec);
      e->graphs.erase(e->graphs.begin());
    }
    e->graphs.push_back({a->batch, a->seq, a->steps, variants, kv2 != nullptr, a->flags, a->cfg_strength, a->workspace, a->rope,
                         a->trajectory, a->split_xchg_peer, exec, nodes});
    g = &e->graphs.back();
  }
  for (int i = first_replayed; i < a->steps; ++i) {
    LEMAS_CUDA_OK(cudaGraphLaunch(g->exec, st));
    count_launches(g->nodes);
  }
  LEMAS_CUDA_OK(cudaMemcpyAsync(a->y, b.y_state, sizeof(float) * state, cudaMemcpyDeviceToDevice, st));
  return LEMAS_OK;
}

int lemas_dit_forward(lemas_engine* e, const lemas_sample_args* a, float t, float* pred, float* hidden_out,
                      void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DitBuffers b;
  LEMAS_TRY(check_args(e, a, 1, &b));
  LEMAS_REQUIRE(pred && a->text_uncond, "lemas_dit_forward: pred and text_uncond are required");
  LEMAS_TRY(prepare(e, b, a, 2, 1, &t, 0, st));
  LEMAS_TRY(dit_forward(e, b, a->batch, a->seq, 2, b.mod, a->kv_len ? b.kv_len2 : nullptr, a->rope, pred, nullptr, 0, st));
  if (hidden_out)
    LEMAS_CUDA_OK(cudaMemcpyAsync(hidden_out, b.x, sizeof(float) * 2LL * a->batch * a->seq * e->cfg.dim,
                                  cudaMemcpyDeviceToDevice, st));
  return LEMAS_OK;
}
}

// ------------------------------------------------------------------------------------------------ Vocos
namespace lemas {

__global__ void mel_to_rows_kernel(const float* __restrict__ mel, __half* __restrict__ out, int batch, int ch, int t) {
  // [b, ch, t] fp32 -> [b, t, 128] fp16 (zero padded channels)
  const long total = (long)batch * t * 128;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int cidx = (int)(i & 127);
    const long bt = i >> 7;
    const int b = (int)(bt / t), tt = (int)(bt - (long)b * t);
    out[i] = __float2half_rn(cidx < ch ? mel[((long)b * ch + cidx) * t + tt] : 0.f);
  }
}

struct VocosBuffers {
  __half *mel16, *a16, *h16;
  float *e32, *x32, *head32, *frames;
  int head_ld;
  int64_t bytes;
};

static VocosBuffers carve_vocos(const lemas_vocos_weights& w, int batch, int t, void* ws) {
  VocosBuffers b;
  Carver cv(ws);
  const int64_t R = (int64_t)batch * t;
  b.head_ld = 1152;
  b.mel16 = cv.take<__half>(R * 128);
  b.e32 = cv.take<float>(R * w.dim);
  b.x32 = cv.take<float>(R * w.dim);
  b.a16 = cv.take<__half>(R * w.dim);
  b.h16 = cv.take<__half>(R * w.inter);
  b.head32 = cv.take<float>(R * b.head_ld);
  b.frames = cv.take<float>(R * 1024);
  b.bytes = align_up(cv.off, 1024);
  return b;
}

}  // namespace lemas

extern "C" {

int64_t lemas_vocos_workspace_bytes(const lemas_vocos_weights* w, int32_t batch, int32_t t) {
  if (!w) return -1;
  return carve_vocos(*w, batch, t, nullptr).bytes;
}

int lemas_vocos_decode(const lemas_vocos_weights* w, const float* mel, float* wav, int32_t batch, int32_t t,
                       void* workspace, int64_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LEMAS_REQUIRE(w && mel && wav && workspace, "lemas_vocos_decode: null argument");
  LEMAS_REQUIRE(w->in_ch <= 128 && w->dim % 128 == 0 && w->dim <= 1024 && w->inter % 64 == 0,
                "lemas_vocos_decode: unsupported dims");
  LEMAS_REQUIRE(t >= 2 && batch >= 1, "lemas_vocos_decode: need at least 2 frames");
  if (!lemas_device_supported())
    return fail(LEMAS_ERR_UNSUPPORTED,
                "CUDA error: no kernel image is available for execution on the device (liblemas_b200 is sm_100a only)");
  VocosBuffers b = carve_vocos(*w, batch, t, workspace);
  LEMAS_REQUIRE(workspace_bytes >= b.bytes, "lemas_vocos_decode: workspace too small");
  const int R = batch * t;
  const int dim = w->dim, inter = w->inter;
  const int bn_dim = dim % 256 == 0 ? 256 : 128;
  const int bn_inter = inter % 256 == 0 ? 256 : 128;
  {
    const long total = (long)R * 128;
    int grid = (int)((total + 255) / 256);
    if (grid > sm_count() * 16) grid = sm_count() * 16;
    mel_to_rows_kernel<<<grid, 256, 0, st>>>(mel, b.mel16, batch, w->in_ch, t);
    LEMAS_LAUNCHED(1);
  }
  {  // backbone.embed: Conv1d(in_ch -> dim, k=7, pad=3) as a 7-tap GEMM
    lemas_gemm_desc d = {};
    d.a = b.mel16; d.batches = batch; d.rows = t; d.lda = 128; d.a_cols = 128;
    d.w = w->embed_w; d.w_rows = 7 * dim; d.ldw = 128; d.n = dim; d.k_per_tap = 128; d.taps = 7; d.tap_pad = 3;
    d.w_tap_stride = dim; d.block_n = bn_dim; d.epilogue = LEMAS_EPI_BIAS_F32; d.bias = w->embed_b;
    d.out32 = b.e32; d.ld32 = dim; d.seq_len = t;
    LEMAS_TRY(gemm_launch(d, st));
  }
  LEMAS_TRY(lemas_ln_affine(b.e32, w->norm_w, w->norm_b, nullptr, b.x32, R, dim, 1e-6f, st));
  for (int l = 0; l < w->layers; ++l) {
    const lemas_vocos_layer& L = w->blocks[l];
    LEMAS_TRY(lemas_dwconv7_ln(b.x32, L.dw_w, L.dw_b, L.ln_w, L.ln_b, b.a16, batch, t, dim, st));
    {
      lemas_gemm_desc d = base_desc(b.a16, 1, R, dim, L.w1, inter, dim, inter, t, LEMAS_EPI_GELU_ERF_F16, bn_inter);
      d.bias = L.b1; d.out16 = b.h16; d.ld16 = inter;
      LEMAS_TRY(gemm_launch(d, st));
    }
    {
      lemas_gemm_desc d = base_desc(b.h16, 1, R, inter, L.w2, dim, inter, dim, t, LEMAS_EPI_GATE_RESID_F32, bn_dim);
      d.bias = L.b2; d.resid = b.x32; d.ldr = dim; d.out32 = b.x32; d.ld32 = dim; d.gate = L.gamma; d.gate_bstride = 0;
      LEMAS_TRY(gemm_launch(d, st));
    }
  }
  LEMAS_TRY(lemas_ln_affine(b.x32, w->final_w, w->final_b, b.a16, nullptr, R, dim, 1e-6f, st));
  {
    lemas_gemm_desc d = base_desc(b.a16, 1, R, dim, w->head_w, 1152, dim, 1026, t, LEMAS_EPI_BIAS_F32, 128);
    d.bias = w->head_b; d.out32 = b.head32; d.ld32 = b.head_ld;
    LEMAS_TRY(gemm_launch(d, st));
  }
  LEMAS_TRY(lemas_istft_1024(b.head32, b.head_ld, b.frames, wav, batch, t, st));
  return LEMAS_OK;
}
}
