// HBM-bound kernels of the hot path: fused LayerNorm+modulate, skinny fp32 linears for the time/AdaLN
// path, operand packing, CFG+Euler update, Vocos depthwise-conv+LN and the inverse STFT.
// All loads/stores are 16-byte vectorised and coalesced (one warp per row for the row-wise reductions).
#include "common.h"
#include "ptx.cuh"

namespace lemas {

// ------------------------------------------------------------------------------------------------
// LayerNorm (no affine, eps) * (1 + scale) + shift -> fp16      modules.py:314 / :637 / :335
// One warp per row, dim <= 1024 and dim % 128 == 0: lane owns float4 #(lane + 32 j).
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAX_VEC = 8;

template <bool AFFINE>
__global__ void __launch_bounds__(256)
ln_kernel(const float* __restrict__ x, const float* __restrict__ p0, const float* __restrict__ p1, int mod_bstride,
          __half* __restrict__ out16, float* __restrict__ out32, int rows, int dim, int seq_len, float eps,
          const int* __restrict__ row_limit) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (row_limit != nullptr) {  // ragged batches: rows beyond the sequence's limit are not needed (lemas_sample_args.flags)
    const int b = row / seq_len;
    if (row - b * seq_len >= __ldg(row_limit + b)) return;
  }
  const int lane = threadIdx.x & 31;
  const int nvec = dim >> 7;
  const float4* xr = reinterpret_cast<const float4*>(x + (long)row * dim);
  float4 v[LN_MAX_VEC];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_VEC; ++j)
    if (j < nvec) {
      v[j] = xr[lane + 32 * j];
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  const float mean = warp_sum(s) / dim;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_VEC; ++j)
    if (j < nvec) {
      float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  const float rstd = rsqrtf(warp_sum(q) / dim + eps);
  const long moff = AFFINE ? 0 : (long)(row / seq_len) * mod_bstride;
  const float4* a4 = reinterpret_cast<const float4*>(p0 + moff);  // scale (modulate) or weight (affine)
  const float4* b4 = reinterpret_cast<const float4*>(p1 + moff);  // shift / bias
#pragma unroll
  for (int j = 0; j < LN_MAX_VEC; ++j)
    if (j < nvec) {
      const int idx = lane + 32 * j;
      float4 a = __ldg(a4 + idx), b = __ldg(b4 + idx);
      float4 o;
      if (AFFINE) {
        o.x = (v[j].x - mean) * rstd * a.x + b.x;
        o.y = (v[j].y - mean) * rstd * a.y + b.y;
        o.z = (v[j].z - mean) * rstd * a.z + b.z;
        o.w = (v[j].w - mean) * rstd * a.w + b.w;
      } else {
        o.x = (v[j].x - mean) * rstd * (1.f + a.x) + b.x;
        o.y = (v[j].y - mean) * rstd * (1.f + a.y) + b.y;
        o.z = (v[j].z - mean) * rstd * (1.f + a.z) + b.z;
        o.w = (v[j].w - mean) * rstd * (1.f + a.w) + b.w;
      }
      if (out16) {
        uint2 u = make_uint2(pack_half2(o.x, o.y), pack_half2(o.z, o.w));
        reinterpret_cast<uint2*>(out16 + (long)row * dim)[idx] = u;
      }
      if (out32) reinterpret_cast<float4*>(out32 + (long)row * dim)[idx] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// Skinny fp32 linear: y[m, n] = act_out(act_in(x)[m, :] . W[n, :] + b[n]), m <= 32 per launch.
// Persistent blocks stage act_in(x) in shared memory once; each warp then streams pairs of weight
// rows (coalesced float4) and keeps 2 x m accumulators per lane.  Weight-read bound by design:
// these are the 25 MB-per-block AdaLN linears (modules.py:311) hoisted out of the ODE loop.
// ------------------------------------------------------------------------------------------------
DEVI float silu_f(float x) { return x / (1.f + __expf(-x)); }

__global__ void __launch_bounds__(256)
skinny_linear_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     float* __restrict__ y, int m, int k, int n, int act_in, int act_out) {
  extern __shared__ float xs[];  // [32][k], rows >= m zero
  for (int i = threadIdx.x; i < 32 * k; i += blockDim.x) {
    int r = i / k;
    float val = r < m ? x[i] : 0.f;
    xs[i] = act_in ? silu_f(val) : val;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int kvec = k >> 2;
  for (int n0 = wid * 2; n0 < n; n0 += warps_total * 2) {
    const bool has2 = n0 + 1 < n;
    const float4* w0 = reinterpret_cast<const float4*>(w + (long)n0 * k);
    const float4* w1 = reinterpret_cast<const float4*>(w + (long)(has2 ? n0 + 1 : n0) * k);
    float acc0[32], acc1[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) acc0[r] = acc1[r] = 0.f;
    for (int c = lane; c < kvec; c += 32) {
      const float4 a = __ldg(w0 + c), b = __ldg(w1 + c);
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const float4 xv = reinterpret_cast<const float4*>(xs + r * k)[c];
        acc0[r] += (a.x * xv.x + a.y * xv.y) + (a.z * xv.z + a.w * xv.w);
        acc1[r] += (b.x * xv.x + b.y * xv.y) + (b.z * xv.z + b.w * xv.w);
      }
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      acc0[r] = warp_sum(acc0[r]);
      acc1[r] = warp_sum(acc1[r]);
    }
    // lane r writes row r (static register indexing via unrolled select)
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r)
      if (lane == r) { o0 = acc0[r]; o1 = acc1[r]; }
    if (lane < m) {
      float v0 = o0 + (bias ? bias[n0] : 0.f);
      y[(long)lane * n + n0] = act_out ? silu_f(v0) : v0;
      if (has2) {
        float v1 = o1 + (bias ? bias[n0 + 1] : 0.f);
        y[(long)lane * n + n0 + 1] = act_out ? silu_f(v1) : v1;
      }
    }
  }
}

// modules.py:149-161 (dim 256): w_k = exp(-k ln(1e4)/127); out = [sin(1000 t w_k) | cos(1000 t w_k)]
__global__ void time_sinusoid_kernel(const float* __restrict__ t, float* __restrict__ out, int m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * 128) return;
  const int r = i >> 7, k = i & 127;
  const float wk = expf((float)k * -(9.210340371976184f / 127.0f));
  const float arg = 1000.0f * t[r] * wk;
  out[r * 256 + k] = sinf(arg);
  out[r * 256 + 128 + k] = cosf(arg);
}

// dit.py:93-97, step-invariant columns: rows [0,rows) = (cond | text_c), rows [rows, 2 rows) = (0 | text_u)
__global__ void pack_cond_text_kernel(const float* __restrict__ cond, const float* __restrict__ text_c,
                                      const float* __restrict__ text_u, __half* __restrict__ out, int rows, int mel,
                                      int text_dim, int ld, int n_variants) {
  const long total = (long)n_variants * rows * ld;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int col = (int)(i % ld);
    const long r = i / ld;
    const int variant = (int)(r / rows);
    const long row = r - (long)variant * rows;
    float v = 0.f;
    if (col < mel) {
      v = variant == 0 ? cond[row * mel + col] : 0.f;
    } else if (col < mel + text_dim) {
      v = (variant == 0 ? text_c : text_u)[row * text_dim + (col - mel)];
    }
    out[i] = __float2half_rn(v);
  }
}

__global__ void cast_pad_kernel(const float* __restrict__ x, __half* __restrict__ out, int rows, int cols, int ld,
                                int copies) {
  const long per = (long)rows * ld;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < per; i += (long)gridDim.x * blockDim.x) {
    const int col = (int)(i % ld);
    const long row = i / ld;
    const __half h = __float2half_rn(col < cols ? x[row * cols + col] : 0.f);
    for (int c = 0; c < copies; ++c) out[c * per + i] = h;
  }
}

// cfm.py:420-424 + Euler step of torchdiffeq: one thread per (row, mel) element
__global__ void cfg_euler_kernel(const float* __restrict__ pred, int ld_pred, float* __restrict__ y,
                                 __half* __restrict__ x16, int ld_x16, int copies, float* __restrict__ traj, int rows,
                                 int mel, float cfg_t, int use_cfg, float dt) {
  const long total = (long)rows * mel;
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long row = i / mel;
  const int c = (int)(i - row * mel);
  float f = pred[row * ld_pred + c];
  if (use_cfg) {
    const float pu = pred[(rows + row) * ld_pred + c];
    f = f + (f - pu) * cfg_t;
    f = fminf(fmaxf(f, -20.f), 20.f);
  }
  const float yn = y[i] + dt * f;
  y[i] = yn;
  if (traj) traj[i] = yn;
  const __half h = __float2half_rn(yn);
  for (int k = 0; k < copies; ++k) x16[((long)k * rows + row) * ld_x16 + c] = h;
}

// Same update with the step's scalars read from device memory (graph-replayed ODE steps: the launch arguments of a
// captured step are frozen, so everything that changes from step to step lives in `state`).
//   state[0] = cfg * (1 - t)^2, state[1] = dt, state[2] = bit pattern of the step index
__global__ void cfg_euler_dev_kernel(const float* __restrict__ pred, int ld_pred, float* __restrict__ y,
                                     __half* __restrict__ x16, int ld_x16, int copies, float* __restrict__ traj,
                                     long traj_stride, int rows, int mel, const float* __restrict__ state, int use_cfg) {
  const long total = (long)rows * mel;
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float cfg_t = state[0], dt = state[1];
  const int step = __float_as_int(state[2]);
  const long row = i / mel;
  const int c = (int)(i - row * mel);
  float f = pred[row * ld_pred + c];
  if (use_cfg) {
    const float pu = pred[(rows + row) * ld_pred + c];
    f = f + (f - pu) * cfg_t;
    f = fminf(fmaxf(f, -20.f), 20.f);
  }
  const float yn = y[i] + dt * f;
  y[i] = yn;
  if (traj) traj[(long)(step + 1) * traj_stride + i] = yn;
  const __half h = __float2half_rn(yn);
  for (int k = 0; k < copies; ++k) x16[((long)k * rows + row) * ld_x16 + c] = h;
}

// Start of an ODE step (single block): copy this step's row of the pre-computed AdaLN modulation table into the
// fixed buffer every kernel of the step reads, publish the step scalars, advance the step counter.
__global__ void __launch_bounds__(1024)
step_begin_kernel(const float* __restrict__ mod_table, long mod_w, float* __restrict__ mod_cur,
                  const float* __restrict__ t_grid, int* __restrict__ step_ctr, float* __restrict__ state,
                  float cfg_strength, int* __restrict__ row_limit, const int* __restrict__ kv_len2, int n_seq, int seq,
                  int steps) {
  const int step = *step_ctr;
  __syncthreads();  // everyone has read the counter before thread 0 advances it
  // Ragged batches (LEMAS_SAMPLE_SKIP_PADDED_ROWS): rows of sequence b that can still reach one of its valid rows
  // before the last step — 30 rows per remaining step through the two k=31 convolutions of the position embedding
  // (dit.py:97-98) — rounded up to the 128-key attention block so that every key row a valid query can read is fresh.
  if (row_limit != nullptr)
    for (int b = threadIdx.x; b < n_seq; b += blockDim.x) {
      const int need = kv_len2[b] + 30 * (steps - 1 - step);
      row_limit[b] = min(seq, (need + 127) / 128 * 128);
    }
  const float4* src = reinterpret_cast<const float4*>(mod_table + (long)step * mod_w);
  float4* dst = reinterpret_cast<float4*>(mod_cur);
  for (long i = threadIdx.x; i < mod_w / 4; i += blockDim.x) dst[i] = src[i];
  if (threadIdx.x == 0) {
    const float t = t_grid[step];
    const float one_minus_t = 1.0f - t;
    state[0] = cfg_strength * (one_minus_t * one_minus_t);  // cfm.py:420
    state[1] = t_grid[step + 1] - t;
    state[2] = __int_as_float(step);
    *step_ctr = step + 1;
  }
}

int step_begin_launch(const float* mod_table, long mod_w, float* mod_cur, const float* t_grid, int* step_ctr,
                      float* state, float cfg_strength, int* row_limit, const int* kv_len2, int n_seq, int seq, int steps,
                      cudaStream_t st) {
  step_begin_kernel<<<1, 1024, 0, st>>>(mod_table, mod_w, mod_cur, t_grid, step_ctr, state, cfg_strength, row_limit,
                                        kv_len2, n_seq, seq, steps);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

// Folded LayerNorm, pre-loop: A operands of the u / v GEMMs.  For layer l and norm w (0 = attn_norm, 1 = ff_norm) the
// matrix [2 steps, dim] holds row 2 s = fp16(1 + scale_{s,l,w}) and row 2 s + 1 = fp16(shift_{s,l,w}), read from the
// all-steps modulation table (chunk order shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp; modules.py:312).
__global__ void ln_fold_pack_kernel(const float* __restrict__ mod, long mod_w, int steps, int depth, int dim,
                                    __half* __restrict__ out) {
  const long total = (long)depth * 2 * 2 * steps * dim;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dim);
    long t = i / dim;
    const int row = (int)(t % (2 * steps));
    t /= 2 * steps;
    const int w = (int)(t & 1), l = (int)(t >> 1);
    const int s = row >> 1, is_shift = row & 1;
    const float* m = mod + (long)s * mod_w + (long)l * 6 * dim + (w ? 3 * dim : 0);   // shift, scale of this norm
    const float v = is_shift ? m[c] : 1.0f + m[dim + c];
    out[i] = __float2half_rn(v);
  }
}

int ln_fold_pack_launch(const float* mod, long mod_w, int steps, int depth, int dim, void* out16, cudaStream_t st) {
  const long total = (long)depth * 4 * steps * dim;
  long grid = (total + 255) / 256;
  if (grid > (long)sm_count() * 16) grid = (long)sm_count() * 16;
  ln_fold_pack_kernel<<<(int)grid, 256, 0, st>>>(mod, mod_w, steps, depth, dim, (__half*)out16);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

// Two-GPU CFG split: swap this step's `pred` with the peer GPU over NVLink (peer stores), one CTA, one kernel per step.
//   state[2] = step, state[3] = call epoch (bit patterns);  flags[0] = "ready", flags[1] = "data", written by the PEER.
// Protocol per step (targets grow monotonically: epoch * 4096 + step + 1, so nothing is ever reset):
//   1. tell the peer that everything that read our exchange buffer in the previous step has finished (this kernel runs
//      after it in stream order), i.e. it may overwrite the slot it owns here;  2. wait for the same from the peer;
//   3. copy our variant's pred into the peer's buffer, fence, raise its data flag;  4. wait for our data flag.
// Every wait has a deadline (a peer that died must end this process with a launch failure, not hang the GPU).
DEVI void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
DEVI int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
DEVI void split_wait(const int* flag, int target, const char* what) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < target) {
    if (clock64() - t0 > 6000000000ll) {   // ~3 s
      printf("lemas: two-GPU CFG split timed out waiting for the peer (%s, target %d, flag %d)\n", what, target,
             ld_acquire_sys(flag));
      __trap();
    }
    __nanosleep(64);
  }
}

__global__ void __launch_bounds__(1024)
cfg_split_exchange_kernel(const float* __restrict__ local, float* __restrict__ peer, long slot_floats, int variant,
                          const float* __restrict__ state, const int* flags_local, int* flags_peer) {
  const int step = __float_as_int(state[2]);
  const int target = __float_as_int(state[3]) * 4096 + step + 1;
  if (threadIdx.x == 0) {
    st_release_sys(flags_peer + 0, target);
    split_wait(flags_local + 0, target, "ready");
  }
  __syncthreads();
  const float4* src = reinterpret_cast<const float4*>(local + (long)variant * slot_floats);
  float4* dst = reinterpret_cast<float4*>(peer + (long)variant * slot_floats);
  for (long i = threadIdx.x; i < slot_floats / 4; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    st_release_sys(flags_peer + 1, target);
    split_wait(flags_local + 1, target, "data");
  }
}

int cfg_split_exchange_launch(const float* local, float* peer, long slot_floats, int variant, const float* state,
                              const int* flags_local, int* flags_peer, cudaStream_t st) {
  cfg_split_exchange_kernel<<<1, 1024, 0, st>>>(local, peer, slot_floats, variant, state, flags_local, flags_peer);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int cfg_euler_dev_launch(const float* pred, int ld_pred, float* y, void* x16, int ld_x16, int copies, float* traj,
                         long traj_stride, int rows, int mel, const float* state, int use_cfg, cudaStream_t st) {
  const long total = (long)rows * mel;
  cfg_euler_dev_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(pred, ld_pred, y, (__half*)x16, ld_x16, copies, traj,
                                                                  traj_stride, rows, mel, state, use_cfg);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

// ------------------------------------------------------------------------------------------------
// Vocos ConvNeXtBlock front: depthwise conv k=7 pad 3 over time, then LayerNorm(affine) -> fp16.
// One warp per (b, t) row; weights tap-major [7, dim].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dwconv7_ln_kernel(const float* __restrict__ x, const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                  const float* __restrict__ ln_w, const float* __restrict__ ln_b, __half* __restrict__ out, int batch,
                  int t_len, int dim) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= batch * t_len) return;
  const int lane = threadIdx.x & 31;
  const int b = row / t_len, t = row - b * t_len;
  const int nvec = dim >> 7;
  float4 acc[LN_MAX_VEC];
#pragma unroll
  for (int j = 0; j < LN_MAX_VEC; ++j)
    if (j < nvec) acc[j] = __ldg(reinterpret_cast<const float4*>(dw_b) + lane + 32 * j);
#pragma unroll
  for (int tap = 0; tap < 7; ++tap) {
    const int tt = t + tap - 3;
    if (tt < 0 || tt >= t_len) continue;
    const float4* xr = reinterpret_cast<const float4*>(x + ((long)b * t_len + tt) * dim);
    const float4* wr = reinterpret_cast<const float4*>(dw_w + (long)tap * dim);
#pragma unroll
    for (int j = 0; j < LN_MAX_VEC; ++j)
      if (j < nvec) {
        const float4 xv = xr[lane + 32 * j];
        const float4 wv = __ldg(wr + lane + 32 * j);
        acc[j].x += xv.x * wv.x; acc[j].y += xv.y * wv.y; acc[j].z += xv.z * wv.z; acc[j].w += xv.w * wv.w;
      }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_VEC; ++j)
    if (j < nvec) s += (acc[j].x + acc[j].y) + (acc[j].z + acc[j].w);
  const float mean = warp_sum(s) / dim;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_VEC; ++j)
    if (j < nvec) {
      float a = acc[j].x - mean, bb = acc[j].y - mean, c = acc[j].z - mean, d = acc[j].w - mean;
      q += (a * a + bb * bb) + (c * c + d * d);
    }
  const float rstd = rsqrtf(warp_sum(q) / dim + 1e-6f);
#pragma unroll
  for (int j = 0; j < LN_MAX_VEC; ++j)
    if (j < nvec) {
      const int idx = lane + 32 * j;
      const float4 g = __ldg(reinterpret_cast<const float4*>(ln_w) + idx);
      const float4 be = __ldg(reinterpret_cast<const float4*>(ln_b) + idx);
      uint2 u = make_uint2(pack_half2((acc[j].x - mean) * rstd * g.x + be.x, (acc[j].y - mean) * rstd * g.y + be.y),
                           pack_half2((acc[j].z - mean) * rstd * g.z + be.z, (acc[j].w - mean) * rstd * g.w + be.w));
      reinterpret_cast<uint2*>(out + (long)row * dim)[idx] = u;
    }
}

// ------------------------------------------------------------------------------------------------
// ISTFT head (vocos ISTFTHead + torch.istft(center=True)), n_fft = 1024, hop = 256, periodic hann.
// Kernel 1: one block per frame: polar -> Hermitian spectrum -> 1024-point inverse FFT in shared memory
//           (radix-2, decimation in time, bit-reversed load) -> * window -> frames[b, t, 1024].
// Kernel 2: overlap-add of the <= 4 frames covering each output sample / window-square envelope.
// ------------------------------------------------------------------------------------------------
constexpr int NFFT = 1024;
constexpr int HOP = 256;
constexpr int NBINS = NFFT / 2 + 1;

__global__ void __launch_bounds__(256)
istft_frames_kernel(const float* __restrict__ head, int ld_head, float* __restrict__ frames) {
  __shared__ float2 buf[NFFT];
  __shared__ float2 tw[NFFT / 2];
  const long frame = blockIdx.x;
  const float* hrow = head + frame * ld_head;
  for (int k = threadIdx.x; k < NFFT / 2; k += blockDim.x) {
    float s, c;
    sincospif((float)k * (2.0f / NFFT), &s, &c);  // exp(+2 pi i k / N): inverse transform
    tw[k] = make_float2(c, s);
  }
  for (int k = threadIdx.x; k < NBINS; k += blockDim.x) {
    const float mag = fminf(expf(hrow[k]), 100.0f);
    float s, c;
    sincosf(hrow[NBINS + k], &s, &c);
    float re = mag * c, im = mag * s;
    if (k == 0 || k == NFFT / 2) im = 0.f;  // C2R ignores the imaginary part of DC and Nyquist
    const int r = __brev((unsigned)k) >> 22;
    buf[r] = make_float2(re, im);
    if (k > 0 && k < NFFT / 2) {
      const int r2 = __brev((unsigned)(NFFT - k)) >> 22;
      buf[r2] = make_float2(re, -im);
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int len = 2; len <= NFFT; len <<= 1) {
    const int half = len >> 1;
    const int tstep = NFFT / len;
    for (int i = threadIdx.x; i < NFFT / 2; i += blockDim.x) {
      const int grp = i / half, j = i - grp * half;
      const int i0 = grp * len + j, i1 = i0 + half;
      const float2 w = tw[j * tstep];
      const float2 a = buf[i0], b = buf[i1];
      const float2 wb = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
      buf[i0] = make_float2(a.x + wb.x, a.y + wb.y);
      buf[i1] = make_float2(a.x - wb.x, a.y - wb.y);
    }
    __syncthreads();
  }
  for (int n = threadIdx.x; n < NFFT; n += blockDim.x) {
    const float win = 0.5f - 0.5f * cospif((float)n * (2.0f / NFFT));
    frames[frame * NFFT + n] = buf[n].x * (1.0f / NFFT) * win;
  }
}

__global__ void istft_ola_kernel(const float* __restrict__ frames, float* __restrict__ wav, int batch, int t_len) {
  const int out_len = (t_len - 1) * HOP;
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)batch * out_len) return;
  const int b = (int)(i / out_len);
  const int q = (int)(i - (long)b * out_len) + NFFT / 2;  // position in the un-trimmed signal
  const int t_hi = min(q / HOP, t_len - 1);
  const int t_lo = max((q - (NFFT - 1) + HOP - 1) / HOP, 0);
  float acc = 0.f, env = 0.f;
  for (int t = t_lo; t <= t_hi; ++t) {
    const int n = q - t * HOP;
    const float win = 0.5f - 0.5f * cospif((float)n * (2.0f / NFFT));
    acc += frames[((long)b * t_len + t) * NFFT + n];
    env += win * win;
  }
  wav[i] = acc / env;
}

}  // namespace lemas

using namespace lemas;

extern "C" {

int lemas_ln_modulate_rows(const float* x, const float* scale, const float* shift, int32_t mod_bstride, void* out16,
                           int32_t rows, int32_t dim, int32_t seq_len, const int32_t* row_limit, void* stream) {
  LEMAS_REQUIRE(dim % 128 == 0 && dim <= 128 * LN_MAX_VEC, "lemas_ln_modulate: dim must be a multiple of 128, <= 1024");
  LEMAS_REQUIRE(rows > 0 && seq_len > 0, "lemas_ln_modulate: bad shape");
  LEMAS_CUDA_OK(launch_pdl(ln_kernel<false>, dim3((rows + 7) / 8), dim3(256), 0, (cudaStream_t)stream, x, scale, shift,
                           mod_bstride, (__half*)out16, (float*)nullptr, rows, dim, seq_len, 1e-6f,
                           (const int*)row_limit));
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int lemas_ln_modulate(const float* x, const float* scale, const float* shift, int32_t mod_bstride, void* out16,
                      int32_t rows, int32_t dim, int32_t seq_len, void* stream) {
  return lemas_ln_modulate_rows(x, scale, shift, mod_bstride, out16, rows, dim, seq_len, nullptr, stream);
}

int lemas_ln_affine(const float* x, const float* weight, const float* bias, void* out16, float* out32, int32_t rows,
                    int32_t dim, float eps, void* stream) {
  LEMAS_REQUIRE(dim % 128 == 0 && dim <= 128 * LN_MAX_VEC, "lemas_ln_affine: dim must be a multiple of 128, <= 1024");
  ln_kernel<true><<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, weight, bias, 0, (__half*)out16, out32, rows,
                                                                    dim, 1 << 30, eps, nullptr);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int lemas_skinny_linear_f32(const float* x, const float* w, const float* b, float* y, int32_t m, int32_t k, int32_t n,
                            int32_t act_in, int32_t act_out, void* stream) {
  LEMAS_REQUIRE(k % 128 == 0 && k <= 1536, "lemas_skinny_linear_f32: k must be a multiple of 128, <= 1536");
  LEMAS_REQUIRE(m >= 1, "lemas_skinny_linear_f32: m >= 1");
  const size_t smem = (size_t)32 * k * sizeof(float);
  static unsigned long long configured = 0;
  LEMAS_CUDA_OK(ensure_dynamic_smem(skinny_linear_kernel, 32 * 1536 * 4, configured));
  int grid = sm_count();
  const int warps_needed = (n + 1) / 2;
  if (grid * 8 > warps_needed) grid = (warps_needed + 7) / 8;
  for (int m0 = 0; m0 < m; m0 += 32) {
    const int mm = m - m0 < 32 ? m - m0 : 32;
    skinny_linear_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x + (long)m0 * k, w, b, y + (long)m0 * n, mm, k, n,
                                                                    act_in, act_out);
  }
  LEMAS_LAUNCHED((m + 31) / 32);
  return LEMAS_OK;
}

int lemas_time_sinusoid(const float* t, float* out, int32_t m, void* stream) {
  time_sinusoid_kernel<<<(m * 128 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t, out, m);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int lemas_pack_cond_text(const float* cond, const float* text_c, const float* text_u, void* out16, int32_t rows,
                         int32_t mel, int32_t text_dim, int32_t ld, int32_t n_variants, void* stream) {
  LEMAS_REQUIRE(ld >= mel + text_dim && (n_variants == 1 || n_variants == 2), "lemas_pack_cond_text: bad shape");
  const long total = (long)n_variants * rows * ld;
  int grid = (int)((total + 255) / 256);
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  pack_cond_text_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cond, text_c, text_u, (__half*)out16, rows, mel,
                                                               text_dim, ld, n_variants);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int lemas_cast_pad_f16(const float* x, void* out16, int32_t rows, int32_t cols, int32_t ld, int32_t copies,
                       void* stream) {
  LEMAS_REQUIRE(ld >= cols && copies >= 1, "lemas_cast_pad_f16: bad shape");
  const long per = (long)rows * ld;
  int grid = (int)((per + 255) / 256);
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  cast_pad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__half*)out16, rows, cols, ld, copies);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int lemas_cfg_euler(const float* pred, int32_t ld_pred, float* y, void* x16, int32_t ld_x16, int32_t copies,
                    float* traj_out, int32_t rows, int32_t mel, float t, float dt, float cfg_strength, void* stream) {
  const int use_cfg = cfg_strength >= 1e-5f ? 1 : 0;
  const float one_minus_t = 1.0f - t;
  const float cfg_t = cfg_strength * (one_minus_t * one_minus_t);  // cfm.py:420, fp32 like the reference tensor math
  const long total = (long)rows * mel;
  cfg_euler_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pred, ld_pred, y, (__half*)x16, ld_x16,
                                                                               copies, traj_out, rows, mel, cfg_t,
                                                                               use_cfg, dt);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int lemas_dwconv7_ln(const float* x, const float* dw_w, const float* dw_b, const float* ln_w, const float* ln_b,
                     void* out16, int32_t batch, int32_t t, int32_t dim, void* stream) {
  LEMAS_REQUIRE(dim % 128 == 0 && dim <= 128 * LN_MAX_VEC, "lemas_dwconv7_ln: dim must be a multiple of 128, <= 1024");
  const int rows = batch * t;
  dwconv7_ln_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, dw_w, dw_b, ln_w, ln_b, (__half*)out16, batch,
                                                                      t, dim);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int lemas_istft_1024(const float* head, int32_t ld_head, float* frames_ws, float* wav, int32_t batch, int32_t t,
                     void* stream) {
  LEMAS_REQUIRE(t >= 2 && ld_head >= 2 * NBINS, "lemas_istft_1024: need t >= 2 and ld_head >= 1026");
  istft_frames_kernel<<<batch * t, 256, 0, (cudaStream_t)stream>>>(head, ld_head, frames_ws);
  const long total = (long)batch * (t - 1) * HOP;
  istft_ola_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(frames_ws, wav, batch, t);
  LEMAS_LAUNCHED(2);
  return LEMAS_OK;
}
}
