/*
 * lemas_b200.h — C ABI of liblemas_b200.so: the B200 (sm_100a) implementation of the LEMAS-TTS acoustic hot path.
 *
 * The reference (LEMAS-Project/LEMAS-TTS) has NO native layer or FFI: its hot path is Python calling stock
 * PyTorch ops.  Each entry point below therefore replaces a *Python call site* of the reference (cited as
 * file:line relative to the reference root) rather than an existing native binding.  INTEGRATION.md shows the
 * ctypes stub a maintainer adds at those call sites; lemas-tts_b200/lemas_tts/_native.py is that stub.
 *
 * Conventions
 *  - plain pointers and sizes only; all data pointers are DEVICE pointers unless the name ends in _host
 *  - no hidden allocation: callers pass workspaces (sizes from the *_workspace_bytes functions)
 *  - every launch goes to the given cudaStream_t (passed as void*); nothing synchronises the device
 *  - return value 0 = ok, otherwise an error code; lemas_last_error() gives a thread-local message
 *  - activations that feed tensor cores are fp16, accumulators / residual stream / ODE state are fp32
 */
#ifndef LEMAS_B200_H
#define LEMAS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LEMAS_OK 0
#define LEMAS_ERR_INVALID 1
#define LEMAS_ERR_CUDA 2
#define LEMAS_ERR_UNSUPPORTED 3

const char* lemas_last_error(void);
int lemas_version(void);
/* 1 when the current device is compute capability 10.x (tcgen05/TMA kernels can run), else 0. */
int lemas_device_supported(void);
/* Two-GPU latency mode (lemas_sample_args.split_*): memory one process allocates and its peer process maps.
 * lemas_peer_alloc: zero-filled memory on the current device + its 64-byte CUDA IPC handle (send it to the peer by any
 * host channel); lemas_peer_open: map the peer's allocation for kernels of the current device (NVLink peer access). */
int lemas_peer_alloc(int64_t bytes, void** ptr, void* handle64);
int lemas_peer_open(const void* handle64, void** ptr);
int lemas_peer_close(void* ptr);
int lemas_peer_free(void* ptr);
/* sizeof() of the ABI structs, for bindings to self-check their mirrors: 0 lemas_gemm_desc, 1 lemas_dit_config,
 * 2 lemas_dit_layer, 3 lemas_dit_weights, 4 lemas_sample_args, 5 lemas_vocos_layer, 6 lemas_vocos_weights,
 * 7 lemas_text_block, 8 lemas_text_weights, 9 lemas_prosody_tdnn, 10 lemas_prosody_block, 11 lemas_prosody_weights,
 * 12 lemas_bigvgan_block, 13 lemas_bigvgan_stage, 14 lemas_bigvgan_weights; else -1. */
int lemas_abi_sizeof(int which);
/* Number of kernels this library has launched in the calling process since it was loaded (all entry points). */
int64_t lemas_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Op-level entry points (one kernel launch each).  Used by the engine below and by the op-level parity tests.
 * ---------------------------------------------------------------------------------------------------------- */

/* GEMM epilogues.  acc = A·Wᵀ (fp16 operands, fp32 accumulate in TMEM). */
enum lemas_epilogue {
  LEMAS_EPI_BIAS_F16 = 0,       /* out16 = acc + bias                                                          */
  LEMAS_EPI_QKV_ROPE = 1,       /* modules.py:452-480: +bias; RoPE on q,k heads; q,k -> out16, v -> transposed */
  LEMAS_EPI_GELU_TANH_F16 = 2,  /* modules.py:348-349: out16 = gelu_tanh(acc + bias)                            */
  LEMAS_EPI_GELU_ERF_F16 = 3,   /* vocos ConvNeXtBlock: out16 = gelu(acc + bias)                                */
  LEMAS_EPI_GATE_RESID_F32 = 4, /* modules.py:635,639: out32 = resid + gate[b,col]*mask_rows(acc + bias)        */
  LEMAS_EPI_BIAS_F32 = 5,       /* out32 = acc + bias                                                           */
  LEMAS_EPI_ADD_F32_F16 = 6,    /* dit.py:97 split: out32 = acc + addend[row,col]; out16 = same in fp16         */
  LEMAS_EPI_MISH_F16 = 7,       /* modules.py:172-173: out16 = mish(acc + bias)                                 */
  LEMAS_EPI_MISH_RESID_F32 = 8  /* modules.py:174-175 + dit.py:98: out32 = mish(acc + bias) + resid             */
};

typedef struct lemas_gemm_desc {
  /* A: fp16 activations viewed as [batches, rows, lda] (row-major, lda in elements, multiple of 8).
   * "taps" > 1 turns the GEMM into an im2col-free 1-D convolution along rows: k-iteration (tap, kc) reads rows
   * shifted by (tap - tap_pad) with zero fill outside [0, rows) of the same batch item (TMA out-of-bounds fill). */
  const void* a;
  int32_t batches, rows, lda;
  int32_t a_cols;          /* valid A columns (tensor-map inner extent; reads beyond are zero-filled)           */
  /* W: fp16 [w_rows, ldw] row-major ("out x in", K contiguous — nn.Linear layout). For taps>1 the weight is
   * tap-major: row (tap*w_tap_stride + out_channel). */
  const void* w;
  int32_t w_rows, ldw;
  int32_t n;               /* output columns                                                                   */
  int32_t k_per_tap;       /* reduction length per tap, multiple of 64                                         */
  int32_t taps, tap_pad, w_tap_stride;
  int32_t group_cols;      /* grouped conv: A column offset = (n0 / block_n) * group_cols (0 = dense)           */
  int32_t block_n;         /* 64, 128 or 256                                                                    */
  int32_t epilogue;        /* enum lemas_epilogue                                                               */
  const float* bias;       /* [n] or NULL                                                                       */
  void* out16; int32_t ld16;
  float* out32; int32_t ld32;
  const float* resid; int32_t ldr;      /* residual / addend, fp32                                             */
  const float* gate; int32_t gate_bstride; /* per-column gate, + b*gate_bstride                                 */
  const int32_t* row_valid;  /* [batches_seq] rows >= row_valid[b] contribute zero (modules.py:499-501) or NULL */
  int32_t seq_len;           /* rows per sequence for (b, pos) = divmod(global_row, seq_len)                    */
  const float* rope;         /* [seq_len, 32, 2] cos,sin (LEMAS_EPI_QKV_ROPE)                                   */
  int32_t rope_cols;         /* q/k columns that get RoPE per projection (pe_attn_head*64 or inner)             */
  int32_t inner;             /* heads*64: columns [0,inner)=q, [inner,2*inner)=k, rest=v                        */
  void* vt; int32_t vt_ld;   /* V transposed out: fp16 [b, head, 64, vt_ld]                                     */
  int32_t max_ctas;          /* 0 = one persistent CTA per SM                                                   */
  const int32_t* row_limit;  /* optional, device int32 [sequences]: 256-row tiles that start at or beyond row_limit[b]
                                of their sequence are not computed at all (ragged batches, see lemas_sample_args.flags);
                                honoured by the CTA-pair kernel (block_n 256), NULL = every row                       */
  /* LayerNorm folded into the GEMMs around it (CTA-pair kernel only; all NULL = off).
   * Producer (LEMAS_EPI_GATE_RESID_F32): besides out32 = x_new, write ln_out16 = fp16(x_new * fp16(1 + ln_scale[col]))
   * and the per-row partial (sum, sum of squares) of x_new over each 128-column slice: ln_stats fp32
   * [rows][n / 128][2].  Consumer (QKV_ROPE / GELU epilogues, A = that ln_out16): acc -> rstd (acc - mean u[col]) +
   * v[col] with mean / rstd (eps 1e-6) from the ln_parts partials of ln_stats_in, and u = ln_uv[2 step][col],
   * v = ln_uv[2 step + 1][col] (fp32 [2 steps, n]: u = W (1 + scale), v = W shift for every ODE step, hoisted), step =
   * *ln_step - 1.  Algebraically LayerNorm(x) (1 + scale) + shift fed to the same GEMM (modules.py:314, 637). */
  const float* ln_scale; void* ln_out16; int32_t ln_ld16; float* ln_stats;
  const float* ln_stats_in; int32_t ln_parts; const float* ln_uv; const int32_t* ln_step; int32_t ln_k;
  int32_t tap_dilation;      /* taps > 1: rows are shifted by (tap - tap_pad) * tap_dilation (dilated convolution);
                                0 means 1                                                                          */
} lemas_gemm_desc;

int lemas_gemm_f16(const lemas_gemm_desc* d, void* stream);

/* LayerNorm(eps 1e-6, no affine) * (1 + scale) + shift -> fp16   (modules.py:314, :637, :335).
 * x: fp32 [rows, dim]; scale/shift: fp32 [dim] (+ b*mod_bstride, b = row / seq_len). */
int lemas_ln_modulate(const float* x, const float* scale, const float* shift, int32_t mod_bstride, void* out16,
                      int32_t rows, int32_t dim, int32_t seq_len, void* stream);
/* Same, skipping rows whose position inside their sequence is >= row_limit[row / seq_len] (device int32, may be NULL). */
int lemas_ln_modulate_rows(const float* x, const float* scale, const float* shift, int32_t mod_bstride, void* out16,
                           int32_t rows, int32_t dim, int32_t seq_len, const int32_t* row_limit, void* stream);

/* LayerNorm with affine weight/bias, fp32 in -> fp16 and/or fp32 out (vocos backbone norms). */
int lemas_ln_affine(const float* x, const float* weight, const float* bias, void* out16, float* out32,
                    int32_t rows, int32_t dim, float eps, void* stream);

/* Self-attention over [b, seq, heads*64] (modules.py:483-491): softmax(q·kᵀ/8 + keymask)·v.
 * qk: fp16 [b*seq, ld_qk] with q at column 0 and k at column `inner`; vt: fp16 [b, heads, 64, vt_ld];
 * kv_len: int32 [b] valid keys per batch item (NULL = seq).  out: fp16 [b*seq, inner]. */
int lemas_attention_f16(const void* qk, int32_t ld_qk, const void* vt, int32_t vt_ld, const int32_t* kv_len,
                        void* out16, int32_t batch, int32_t seq, int32_t heads, void* stream);

/* y[m, n] = act_in(x)[m, :] · W[n, :] + b[n], all fp32, m <= 64 (time MLP, AdaLN linears: modules.py:311,332,725).
 * act_in: 0 = identity, 1 = SiLU.  act_out: 0 = identity, 1 = SiLU. */
int lemas_skinny_linear_f32(const float* x, const float* w, const float* b, float* y, int32_t m, int32_t k,
                            int32_t n, int32_t act_in, int32_t act_out, void* stream);

/* Sinusoidal time features (modules.py:149-161): out[m, 0:128] = sin(1000 t_m w_k), out[m,128:256] = cos(...). */
int lemas_time_sinusoid(const float* t, float* out, int32_t m, void* stream);

/* cat(cond | 0, text) -> fp16 A operand of the step-invariant half of the input projection (dit.py:93-97).
 * rows [0, B*N): (cond, text_c); rows [B*N, 2B*N): (0, text_u).  Output row stride ld (zero padded). */
int lemas_pack_cond_text(const float* cond, const float* text_c, const float* text_u, void* out16, int32_t rows,
                         int32_t mel, int32_t text_dim, int32_t ld, int32_t n_variants, void* stream);

/* fp32 [rows, cols] -> fp16 [copies][rows, ld] zero padded. */
int lemas_cast_pad_f16(const float* x, void* out16, int32_t rows, int32_t cols, int32_t ld, int32_t copies,
                       void* stream);

/* CFG + clamp + Euler (cfm.py:420-424 and the torchdiffeq Euler step):
 *   f = cfg>=1e-5 ? clamp(pc + (pc - pu) * cfg * (1-t)^2, -20, 20) : pc ;  y += dt * f
 * pred: fp32 [2 or 1][rows, ld_pred]; y: fp32 [rows, mel]; also refreshes the fp16 copy of y (x16, `copies` times)
 * and optionally stores the new state to traj_out. */
int lemas_cfg_euler(const float* pred, int32_t ld_pred, float* y, void* x16, int32_t ld_x16, int32_t copies,
                    float* traj_out, int32_t rows, int32_t mel, float t, float dt, float cfg_strength,
                    void* stream);

/* Depthwise conv k=7 (pad 3) along time + LayerNorm(affine, eps 1e-6) -> fp16   (vocos ConvNeXtBlock).
 * x: fp32 [b, t, dim]; dw_w: fp32 [7, dim] (tap-major), others fp32 [dim]. */
int lemas_dwconv7_ln(const float* x, const float* dw_w, const float* dw_b, const float* ln_w, const float* ln_b,
                     void* out16, int32_t batch, int32_t t, int32_t dim, void* stream);

/* ISTFT head tail (vocos ISTFTHead + torch.istft(center=True), n_fft 1024, hop 256, hann):
 * head: fp32 [b*t, ld_head] rows = (log-magnitude[513] | phase[513]); wav: fp32 [b, (t-1)*256]; frames: workspace
 * fp32 [b, t, 1024]. */
int lemas_istft_1024(const float* head, int32_t ld_head, float* frames_ws, float* wav, int32_t batch, int32_t t,
                     void* stream);

/* Mel front-end of the reference audio (MelSpec.forward / get_vocos_mel_spectrogram, modules.py:75-101,130-143):
 * |STFT| (n_fft 1024, hop 256, periodic hann, center + reflect padding, power 1) -> mel filterbank -> log(clamp 1e-5).
 * wav: fp32 [batch, wav_ld] (nw valid samples per row, nw > 512); fb: fp32 [513, n_mels] filterbank (HTK, norm=None:
 * torchaudio.functional.melscale_fbanks layout); fb_range: int32 [n_mels, 2] = [first, last+1) non-zero bin of every
 * filter; mel: fp32 [batch, n_mels, nw/256 + 1]. */
int lemas_mel_spectrogram_1024(const float* wav, int32_t batch, int32_t nw, int32_t wav_ld, const float* fb,
                               const int32_t* fb_range, int32_t n_mels, float* mel, void* stream);
/* Same kernel for the `mel_spec_type: bigvgan` front-end (get_bigvgan_mel_spectrogram, modules.py:30-72): reflect padding of
 * (1024 - 256) / 2 samples, no centering, sqrt(re^2 + im^2 + 1e-9); fb = Slaney mel filterbank with Slaney norm
 * (librosa.filters.mel) in the [513, n_mels] layout; nw > 384; mel: fp32 [batch, n_mels, (nw - 256) / 256 + 1]. */
int lemas_mel_spectrogram_bigvgan_1024(const float* wav, int32_t batch, int32_t nw, int32_t wav_ld, const float* fb,
                                       const int32_t* fb_range, int32_t n_mels, float* mel, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Engine-level entry points: the whole sampler / vocoder as a sequence of the launches above, driven from C++.
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct lemas_dit_config {
  int32_t dim, depth, heads, ff_mult, text_dim, mel_dim, rope_heads;
} lemas_dit_config;

typedef struct lemas_dit_layer {
  const void* w_qkv; const float* b_qkv;  /* fp16 [3*inner, dim]; fp32 [3*inner]   (to_q|to_k|to_v stacked)    */
  const void* w_out; const float* b_out;  /* fp16 [dim, inner]                                                  */
  const void* w_ff1; const float* b_ff1;  /* fp16 [dim*ff_mult, dim]                                            */
  const void* w_ff2; const float* b_ff2;  /* fp16 [dim, dim*ff_mult]                                            */
} lemas_dit_layer;

typedef struct lemas_dit_weights {
  const float* time_w0; const float* time_b0;   /* fp32 [dim,256],[dim]     transformer.time_embed.time_mlp.0   */
  const float* time_w2; const float* time_b2;   /* fp32 [dim,dim],[dim]     transformer.time_embed.time_mlp.2   */
  const float* adaln_w; const float* adaln_b;   /* fp32 [depth*6*dim + 2*dim, dim]: blocks' attn_norm.linear
                                                   stacked in order, then norm_out.linear                      */
  const void* w_in_x;                           /* fp16 [dim, 128]   input_embed.proj columns of x (zero pad)   */
  const void* w_in_ct; const float* b_in;       /* fp16 [dim, ct_ld] columns of (cond | text), zero padded      */
  int32_t ct_ld;
  const void* conv_w[2]; const float* conv_b[2];/* fp16 tap-major conv_pos_embed.conv1d.{0,2}: [31*dim, dim/16] when
                                                   dim/16 == 64 (grouped path), else block-diagonal [31*dim, dim]  */
  int32_t conv_dense;                           /* 0: grouped path (dim == 1024); 1: block-diagonal dense weights   */
  const void* w_proj; const float* b_proj;      /* fp16 [128, dim] (rows >= mel_dim zero), fp32 [mel_dim]       */
  const lemas_dit_layer* layers;                /* host array [depth]                                           */
} lemas_dit_weights;

typedef struct lemas_engine lemas_engine;

int64_t lemas_engine_workspace_bytes(const lemas_dit_config* cfg, int32_t batch, int32_t seq, int32_t steps);
int lemas_engine_create(const lemas_dit_config* cfg, const lemas_dit_weights* w, lemas_engine** out);
void lemas_engine_destroy(lemas_engine* e);

typedef struct lemas_sample_args {
  int32_t batch, seq, steps;
  const float* t_grid_host;   /* [steps+1] host floats (cfm.py:445-453)                                        */
  float cfg_strength;
  float* y;                   /* in: y0 noise, out: final state; fp32 [batch, seq, mel]                         */
  const float* step_cond;     /* fp32 [batch, seq, mel]   (cfm.py:387-390, already masked)                      */
  const float* text_cond;     /* fp32 [batch, seq, text_dim] (dit.py:212-220 cached embeds)                     */
  const float* text_uncond;
  const int32_t* kv_len;      /* int32 [batch] = duration per row (mask of cfm.py:336-337), NULL when batch==1  */
  const float* rope;          /* fp32 [seq, 32, 2] cos/sin table                                                */
  float* trajectory;          /* optional fp32 [steps+1, batch, seq, mel] (cfm.py:456), NULL to skip            */
  void* workspace; int64_t workspace_bytes;
  int32_t use_graph;          /* 1: capture ONE ODE step into a CUDA graph (cached per shape/workspace in the engine)
                                 and replay it for every step; step-dependent values are read from device memory.
                                 Ignored (eager launches) when profiling is on.                                   */
  int32_t flags;              /* LEMAS_SAMPLE_SKIP_PADDED_ROWS: ragged batches (kv_len != NULL) — the transformer blocks
                                 of Euler step s only compute rows r < kv_len[b] + 30 (steps-1-s), rounded up to 128:
                                 the two k=31 convolutions of the position embedding (dit.py:97-98) are the only path
                                 from a padded row to a valid one, 30 rows per step, so rows beyond that cone can never
                                 reach a valid row of the result.  Valid rows (r < kv_len[b]) are unchanged; padded rows
                                 of the returned state then differ from the reference's (callers slice them off).     */
  /* Two-GPU latency mode: the conditional and the unconditional forward of every Euler step
   * (cfm.py:393-417) run on two GPUs, one process each, and swap their `pred` ([batch * seq, mel] fp32, 0.9 MB at C2)
   * over NVLink with peer stores inside one kernel per step — no NCCL call, no host involvement, graph-replayable.
   * All NULL = off.  This process computes variant `split_variant` (0 = conditional, 1 = unconditional: the caller then
   * passes the dropped inputs as step_cond / text_cond) and both processes end every step with the same state y.
   * split_xchg_local: this GPU, fp32 [2 variants][batch * seq][128]; split_flags_local: this GPU, int32 [2], zeroed once
   * at allocation, written only by the peer; *_peer: the other process' buffers mapped through CUDA IPC.  Both
   * processes must issue the same sequence of lemas_sampler_run calls (flags carry a call epoch). */
  float* split_xchg_local; float* split_xchg_peer;
  int32_t* split_flags_local; int32_t* split_flags_peer;
  int32_t split_variant;
} lemas_sample_args;
#define LEMAS_SAMPLE_SKIP_PADDED_ROWS 1
/* Fold every LayerNorm except layer 0's attn_norm and the final norm into the GEMMs around it (lemas_gemm_desc.ln_*):
 * 64 instead of 1 440 LayerNorm launches per 32-step utterance, same parity — but 1.2 % SLOWER end to end on C2 (the
 * extra epilogue work costs more than the 5 us kernel it replaces), so it is off unless asked for. */
#define LEMAS_SAMPLE_FOLD_LAYERNORM 2

/* CFM.sample's ODE loop (cfm.py:382-456): `steps` Euler steps of the CFG-combined DiT flow. */
int lemas_sampler_run(lemas_engine* e, const lemas_sample_args* a, void* stream);

/* Per-kernel-kind device timing of the sampler (measurement aid, off by default).  enable = 1: every launch issued
 * by lemas_sampler_run / lemas_dit_forward is bracketed by CUDA events on the launching stream (eager launches);
 * enable = 2: lemas_sampler_run records the event pairs INSIDE the captured ODE-step graph and reads them back after
 * every replay, i.e. per-kernel times of the graph-replayed step the production path runs (the pre-loop kernels and
 * step 0 are not included); enable = 0: off.
 * lemas_engine_profile_read synchronises the stream, adds the elapsed times since the last read to ms[kind] and
 * launches[kind] (arrays of LEMAS_PROF_KINDS entries) and clears the record. */
enum lemas_prof_kind {
  LEMAS_PROF_PRELOOP = 0,   /* time MLP, AdaLN table, packing, step-invariant input projection */
  LEMAS_PROF_IN_PROJ = 1,   /* dit.py:97 x-columns GEMM                                         */
  LEMAS_PROF_CONV_POS = 2,  /* modules.py:171-176 grouped conv k31 + Mish (2 launches / forward)*/
  LEMAS_PROF_LN_MOD = 3,    /* LayerNorm + AdaLN modulate                                       */
  LEMAS_PROF_GEMM_QKV = 4,
  LEMAS_PROF_ATTENTION = 5,
  LEMAS_PROF_GEMM_OUT = 6,
  LEMAS_PROF_GEMM_FF1 = 7,
  LEMAS_PROF_GEMM_FF2 = 8,
  LEMAS_PROF_PROJ_OUT = 9,
  LEMAS_PROF_CFG_EULER = 10,
  LEMAS_PROF_TARE = 11,     /* an EMPTY event pair per ODE step: the cost of the measurement itself      */
  LEMAS_PROF_KINDS = 12
};
int lemas_engine_profile(lemas_engine* e, int32_t enable);
int lemas_engine_profile_read(lemas_engine* e, double* ms, int64_t* launches, void* stream);

/* One DiT.forward (dit.py:194-254) on the co-batched cond/uncond rows; pred: fp32 [2*batch*seq, 128]. Testing aid. */
int lemas_dit_forward(lemas_engine* e, const lemas_sample_args* a, float t, float* pred, float* hidden_out,
                      void* stream);

typedef struct lemas_vocos_layer {
  const float* dw_w; const float* dw_b; const float* ln_w; const float* ln_b;  /* fp32 [dim,7],[dim],[dim],[dim] */
  const void* w1; const float* b1;     /* fp16 [inter, dim] */
  const void* w2; const float* b2;     /* fp16 [dim, inter] */
  const float* gamma;                  /* fp32 [dim]        */
} lemas_vocos_layer;

typedef struct lemas_vocos_weights {
  int32_t dim, inter, layers, in_ch;
  const void* embed_w; const float* embed_b;        /* fp16 [7*dim, 128] tap-major, fp32 [dim]                  */
  const float* norm_w; const float* norm_b;
  const lemas_vocos_layer* blocks;                  /* host array [layers]                                      */
  const float* final_w; const float* final_b;
  const void* head_w; const float* head_b;          /* fp16 [1152, dim] (rows >= 1026 zero), fp32 [1026]        */
} lemas_vocos_weights;

int64_t lemas_vocos_workspace_bytes(const lemas_vocos_weights* w, int32_t batch, int32_t t);
/* Vocos.decode (utils_infer.py:549): mel fp32 [batch, in_ch, t] -> wav fp32 [batch, (t-1)*256]. */
int lemas_vocos_decode(const lemas_vocos_weights* w, const float* mel, float* wav, int32_t batch, int32_t t,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * BigVGAN-v2 generator — the reference's `mel_spec_type: bigvgan` vocoder branch (utils_infer.py:144-158 builds
 * `bigvgan.BigVGAN.from_pretrained("nvidia/bigvgan_v2_24khz_100band_256x", use_cuda_kernel=False)` from the un-vendored
 * third_party/BigVGAN submodule; utils_infer.py:550-551 calls `vocoder(mel)`).  Replaces BigVGAN.forward.
 * All activations are time-major [batch, T, Cpad], Cpad = channels rounded up to a multiple of 64, padded channels zero.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct lemas_bigvgan_block {      /* one AMPBlock1: x += conv2_d(act(conv1_d(act(x)))) for 3 dilations          */
  int32_t kernel;                         /* odd kernel size (3, 7, 11)                                                 */
  int32_t dilation[3];                    /* of convs1 (convs2 have dilation 1)                                         */
  const void* w1[3]; const float* b1[3];  /* fp16 tap-major [kernel * Cpad, Cpad] (row = tap * Cpad + out), fp32 [Cpad]  */
  const void* w2[3]; const float* b2[3];
  const float* act[6];                    /* fp32 [2, Cpad]: e^alpha | 1 / (e^beta + 1e-9)   (activations.{0..5})        */
} lemas_bigvgan_block;

typedef struct lemas_bigvgan_stage {
  int32_t rate, ch_in, ch_out;            /* up-sampling rate r; padded channel counts                                   */
  const void* up_w; const float* up_b;    /* ConvTranspose1d(k 2r, stride r, pad r/2) as a 3-tap GEMM: fp16 [3 * r * ch_out,
                                             ch_in], row = tap * r * ch_out + phase * ch_out + out (tap 0, 1, 2 = input row
                                             m-1, m, m+1; unused (tap, phase) blocks zero); bias fp32 [r * ch_out]          */
  lemas_bigvgan_block block[3];
} lemas_bigvgan_stage;

typedef struct lemas_bigvgan_weights {
  int32_t num_mels, ch0, stages, use_tanh;
  const void* pre_w; const float* pre_b;  /* conv_pre: fp16 tap-major [7 * ch0, 128], fp32 [ch0]                          */
  const lemas_bigvgan_stage* stage;       /* host array [stages]                                                         */
  const float* post_act;                  /* activation_post: fp32 [2, Cpad_last]                                         */
  const float* post_w; float post_bias;   /* conv_post: fp32 [7, Cpad_last] tap-major; bias (0 when use_bias_at_final = 0) */
  float aa_filter[12];                    /* Kaiser-sinc low-pass of the anti-aliased activations (cutoff 0.25, hw 0.3)   */
} lemas_bigvgan_weights;

int64_t lemas_bigvgan_workspace_bytes(const lemas_bigvgan_weights* w, int32_t batch, int32_t t);
/* mel fp32 [batch, num_mels, t] -> wav fp32 [batch, t * prod(rate)]. */
int lemas_bigvgan_decode(const lemas_bigvgan_weights* w, const float* mel, float* wav, int32_t batch, int32_t t,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Text embedding (dit.py:51-81 TextEmbedding.forward + ConvNeXtV2Block modules.py:241-269 + GRN modules.py:225-234):
 * runs once per CFM.sample for the conditional and unconditional copies of the text (dit.py:212-220).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct lemas_text_block {
  const float* dw_w; const float* dw_b;       /* fp32 [7, dim] tap-major, [dim]      text_blocks.i.dwconv            */
  const float* ln_w; const float* ln_b;       /* fp32 [dim]                          text_blocks.i.norm              */
  const void* w1; const float* b1;            /* fp16 [inter, dim], fp32 [inter]     text_blocks.i.pwconv1           */
  const float* grn_gamma; const float* grn_beta; /* fp32 [inter]                     text_blocks.i.grn               */
  const void* w2; const float* b2;            /* fp16 [dim, inter], fp32 [dim]       text_blocks.i.pwconv2           */
} lemas_text_block;

typedef struct lemas_text_weights {
  int32_t dim, inter, layers, mask_padding;
  const float* table;                         /* fp32 [text_num_embeds + 1, dim]     text_embed.text_embed.weight    */
  const float* pos;                           /* fp32 [4096, dim] cat(cos, sin) table (modules.py:196-207)           */
  const lemas_text_block* blocks;             /* host array [layers]                                                 */
} lemas_text_weights;

int64_t lemas_text_workspace_bytes(const lemas_text_weights* w, int32_t batch, int32_t seq);
/* ids: int32 [batch, seq], ALREADY shifted by +1, truncated / zero-padded to seq (0 = filler, dit.py:52-55);
 * drop: uint8 [batch], 1 = unconditional copy (token ids replaced by 0 AFTER the filler mask is taken, dit.py:56-60);
 * out: fp32 [batch, seq, dim]. */
int lemas_text_embedding(const lemas_text_weights* w, const int32_t* ids, const uint8_t* drop, float* out, int32_t batch,
                         int32_t seq, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Waveform-side pre / post-processing of infer_batch_process on the device.
 * ---------------------------------------------------------------------------------------------------------- */
int64_t lemas_audio_prep_workspace_bytes(int64_t samples);
/* utils_infer.py:487-493: wav fp32 [channels, samples] (channel stride ch_stride elements) -> mono fp32 [samples]
 * (mean over channels), stats[0] = rms = sqrt(mean(mono^2)), stats[1] = target_rms, and mono = mono * target_rms / rms
 * when rms < target_rms.  The 24 kHz resampling that follows is lemas_resample_sinc.  No host synchronisation. */
int lemas_audio_prep(const float* wav, int32_t channels, int64_t samples, int64_t ch_stride, float target_rms,
                     float* mono, float* stats, void* workspace, int64_t workspace_bytes, void* stream);
/* utils_infer.py:552-553: wav = wav * rms / target_rms in place when rms < target_rms (stats from lemas_audio_prep). */
int lemas_audio_unscale(float* wav, int64_t samples, const float* stats, void* stream);
/* utils_infer.py:581-622, one step of the chunk loop: out (fp64 [na + nb - fade]) = a[: na - fade] ++ (a[na - fade :] *
 * linspace(1, 0, fade) + b[: fade] * linspace(0, 1, fade)) ++ b[fade :], evaluated like numpy's float64 arithmetic,
 * then clipped to [-clip, clip] when clip > 0.  na == 0: out = (double) b (the first chunk).  a: fp64, b: fp32. */
int lemas_audio_crossfade(const double* a, int64_t na, const float* b, int64_t nb, int64_t fade, double* out, double clip,
                          void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Prosody path of the prosody-conditioned model, once per utterance (cfm.py:248-262): 24 kHz -> 16 kHz resampling,
 * kaldi fbank, Pretssel ECAPA-TDNN (prosody_encoder.py:30-334).  fp32 throughout; activations channels-last.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct lemas_prosody_tdnn {           /* TDNNBlock: Conv1d(same padding) -> ReLU -> LayerNorm(eps 1e-12)      */
  const float* w;                             /* fp32 [k][cin/groups][cout]: nn.Conv1d weight permuted (2, 1, 0)      */
  const float* b;                             /* fp32 [cout]                                                          */
  const float* ln_w; const float* ln_b;       /* fp32 [cout]                                                          */
  int32_t cin, cout, k, dil, groups;
} lemas_prosody_tdnn;

typedef struct lemas_prosody_block {          /* SERes2NetBlock (prosody_encoder.py:282-334), in == out channels      */
  lemas_prosody_tdnn tdnn1;
  lemas_prosody_tdnn res2[7];                 /* res2net_scale - 1 TDNNs on channels/scale channels each              */
  lemas_prosody_tdnn tdnn2;
  const float* se_w1; const float* se_b1;     /* fp32 [channels][se_channels] (transposed), [se_channels]             */
  const float* se_w2; const float* se_b2;     /* fp32 [se_channels][channels] (transposed), [channels]                */
} lemas_prosody_block;

typedef struct lemas_prosody_weights {
  int32_t input_dim, channels, n_blocks, scale, se_channels, att_channels, mfa_channels, embed_dim;
  lemas_prosody_tdnn block0;                  /* encoder.blocks.0                                                      */
  const lemas_prosody_block* blocks;          /* host array [n_blocks]: encoder.blocks.1 ..                            */
  lemas_prosody_tdnn mfa;                     /* encoder.mfa                                                           */
  lemas_prosody_tdnn asp_tdnn;                /* encoder.asp.tdnn (3*mfa_channels -> att_channels), tanh after         */
  const float* asp_conv_w; const float* asp_conv_b;   /* fp32 [att_channels][mfa_channels] (transposed), [mfa_channels] */
  const float* asp_norm_w; const float* asp_norm_b;   /* fp32 [2*mfa_channels]                                          */
  const float* fc_w; const float* fc_b;               /* fp32 [2*mfa_channels][embed_dim] (transposed), [embed_dim]     */
} lemas_prosody_weights;

int64_t lemas_prosody_workspace_bytes(const lemas_prosody_weights* w, int32_t batch, int32_t t);
/* ProsodyEncoder.forward(fbank, padding_mask=None) (prosody_encoder.py:421-432): fbank fp32 [batch, t, input_dim] ->
 * L2-normalised embedding fp32 [batch, embed_dim]. */
int lemas_prosody_encode(const lemas_prosody_weights* w, const float* fbank, int32_t batch, int32_t t, float* out,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* torchaudio.functional.resample (cfm.py:254) as a polyphase FIR: out[q*up + j] = sum_k taps[j][k] * xpad[q*down + k],
 * xpad[p] = x[p - width].  taps: fp32 [up][n_taps] (host: sinc_interp_hann, lowpass width 6, rolloff 0.99). */
int lemas_resample_sinc(const float* x, int32_t batch, int32_t n, int32_t x_ld, const float* taps, int32_t n_taps,
                        int32_t width, int32_t up, int32_t down, float* y, int32_t n_out, int32_t y_ld, void* stream);

/* extract_fbank_16k = torchaudio.compliance.kaldi.fbank(num_mel_bins, 16 kHz) (prosody_encoder.py:337-361):
 * wav fp32 [batch, wav_ld] (n >= 400 valid samples) -> fp32 [batch, 1 + (n-400)/160, n_mels].
 * window: fp32 [400] povey window; banks: fp32 [n_mels][257] kaldi mel banks; bank_range: int32 [n_mels][2]. */
int lemas_kaldi_fbank_16k(const float* wav, int32_t batch, int32_t n, int32_t wav_ld, const float* window,
                          const float* banks, const int32_t* bank_range, int32_t n_mels, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LEMAS_B200_H */
