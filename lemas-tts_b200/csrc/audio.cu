// Waveform-side pre / post-processing of infer_batch_process on the device:
//   * reference-audio preparation, utils_infer.py:487-493: mono mix, RMS = sqrt(mean(x^2)), scale up to the target RMS
//     when quieter (the 24 kHz resampling that follows is lemas_resample_sinc, csrc/prosody.cu);
//   * per-chunk output: undo the RMS scaling, utils_infer.py:552-553;
//   * linear cross-fade between consecutive chunks and the final clip, utils_infer.py:581-622.
// The reductions are two-stage with a fixed summation order (bit-reproducible from call to call); the cross-fade is
// evaluated in fp64 with numpy's linspace formula (start + i * step, last weight forced to `stop`), so the faded
// samples carry the same bits as the reference's float64 numpy arithmetic.
#include "common.h"

namespace lemas {

constexpr int AUD_THREADS = 256;
constexpr int AUD_CHUNK = 4096;   // samples per block of the first reduction stage

// mono[i] = mean over channels; part[block] = sum of mono^2 over the block's samples (fixed order inside the block)
__global__ void __launch_bounds__(AUD_THREADS)
audio_mono_sumsq_kernel(const float* __restrict__ wav, int channels, long n, long ch_stride, float* __restrict__ mono,
                        double* __restrict__ part) {
  __shared__ double red[AUD_THREADS];
  const long i0 = (long)blockIdx.x * AUD_CHUNK;
  double acc = 0.0;
  for (long i = i0 + threadIdx.x; i < min(n, i0 + AUD_CHUNK); i += AUD_THREADS) {
    float m = 0.f;
    for (int c = 0; c < channels; ++c) m += wav[c * ch_stride + i];
    if (channels > 1) m /= (float)channels;          // torch.mean over the channel dim
    mono[i] = m;
    acc += (double)(m * m);                          // torch.square in fp32, accumulated wider than fp32
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = AUD_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// stats[0] = rms, stats[1] = target rms
__global__ void audio_rms_kernel(const double* __restrict__ part, int n_part, long n, float target_rms,
                                 float* __restrict__ stats) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n_part; ++i) s += part[i];
    stats[0] = sqrtf((float)(s / (double)n));
    stats[1] = target_rms;
  }
}

// invert = 0: `audio * target_rms / rms` (utils_infer.py:491-492); invert = 1: `wave * rms / target_rms` (:552-553);
// both only when rms < target_rms, with the reference's operation order (multiply, then divide, fp32).
__global__ void audio_scale_kernel(float* __restrict__ x, long n, const float* __restrict__ stats, int invert) {
  const float rms = stats[0], target = stats[1];
  if (!(rms < target)) return;
  const float mul = invert ? rms : target, div = invert ? target : rms;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    x[i] = __fdiv_rn(__fmul_rn(x[i], mul), div);
}

// out = concat(a[: na - nf], a[na - nf :] * linspace(1, 0, nf) + b[: nf] * linspace(0, 1, nf), b[nf :]) in fp64,
// clipped to [-clip, clip] when clip > 0 (utils_infer.py:600-622).
__global__ void crossfade_kernel(const double* __restrict__ a, long na, const float* __restrict__ b, long nb, long nf,
                                 double* __restrict__ out, double clip) {
  const long total = na + nb - nf;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    double v;
    if (i < na - nf) {
      v = a[i];
    } else if (i < na) {
      const long k = i - (na - nf);
      // numpy.linspace(start, stop, nf): step = (stop - start) / (nf - 1); y = k * step + start; y[-1] = stop
      const double div = (double)(nf - 1);
      // separate roundings everywhere (no fused multiply-add): numpy evaluates y = k * step, y += start, and the
      // fade as prev * fade_out + next * fade_in with one rounding per operation
      const double down = (nf > 1 && k == nf - 1) ? 0.0 : (nf > 1 ? __dadd_rn(__dmul_rn((double)k, -1.0 / div), 1.0) : 1.0);
      const double up = (nf > 1 && k == nf - 1) ? 1.0 : (nf > 1 ? __dadd_rn(__dmul_rn((double)k, 1.0 / div), 0.0) : 0.0);
      v = __dadd_rn(__dmul_rn(a[i], down), __dmul_rn((double)b[k], up));
    } else {
      v = (double)b[i - (na - nf)];
    }
    if (clip > 0.0) v = fmin(fmax(v, -clip), clip);
    out[i] = v;
  }
}

__global__ void f32_to_f64_kernel(const float* __restrict__ x, double* __restrict__ y, long n, double clip) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double v = (double)x[i];
    if (clip > 0.0) v = fmin(fmax(v, -clip), clip);
    y[i] = v;
  }
}

static int grid_for(long n) {
  long g = (n + 255) / 256;
  const long cap = (long)sm_count() * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace lemas

using namespace lemas;

extern "C" {

int64_t lemas_audio_prep_workspace_bytes(int64_t samples) { return ((samples + AUD_CHUNK - 1) / AUD_CHUNK) * 8 + 64; }

int lemas_audio_prep(const float* wav, int32_t channels, int64_t samples, int64_t ch_stride, float target_rms,
                     float* mono, float* stats, void* workspace, int64_t workspace_bytes, void* stream) {
  LEMAS_REQUIRE(wav && mono && stats && workspace, "lemas_audio_prep: null pointer");
  LEMAS_REQUIRE(channels >= 1 && samples >= 1, "lemas_audio_prep: bad shape");
  LEMAS_REQUIRE(workspace_bytes >= lemas_audio_prep_workspace_bytes(samples), "lemas_audio_prep: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_part = (int)((samples + AUD_CHUNK - 1) / AUD_CHUNK);
  double* part = static_cast<double*>(workspace);
  audio_mono_sumsq_kernel<<<n_part, AUD_THREADS, 0, st>>>(wav, channels, samples, ch_stride, mono, part);
  LEMAS_LAUNCHED(1);
  audio_rms_kernel<<<1, 32, 0, st>>>(part, n_part, samples, target_rms, stats);
  LEMAS_LAUNCHED(1);
  audio_scale_kernel<<<grid_for(samples), 256, 0, st>>>(mono, samples, stats, 0);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int lemas_audio_unscale(float* wav, int64_t samples, const float* stats, void* stream) {
  LEMAS_REQUIRE(wav && stats && samples >= 1, "lemas_audio_unscale: bad argument");
  audio_scale_kernel<<<grid_for(samples), 256, 0, static_cast<cudaStream_t>(stream)>>>(wav, samples, stats, 1);
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

int lemas_audio_crossfade(const double* a, int64_t na, const float* b, int64_t nb, int64_t fade, double* out, double clip,
                          void* stream) {
  LEMAS_REQUIRE(b && out && nb >= 0 && na >= 0 && fade >= 0 && fade <= na && fade <= nb, "lemas_audio_crossfade: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (na == 0) {
    f32_to_f64_kernel<<<grid_for(nb), 256, 0, st>>>(b, out, nb, clip);
  } else {
    LEMAS_REQUIRE(a != nullptr, "lemas_audio_crossfade: null accumulator");
    crossfade_kernel<<<grid_for(na + nb - fade), 256, 0, st>>>(a, na, b, nb, fade, out, clip);
  }
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}
}
