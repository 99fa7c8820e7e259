// Mel front-end of the reference audio:
//
//   MelSpec.forward / get_vocos_mel_spectrogram       lemas_tts/model/modules.py:75-101,130-143
//     |STFT| (n_fft 1024, hop 256, periodic hann, center=True with reflect padding, power 1)
//       -> HTK mel filterbank (n_mels x 513, norm=None) -> log(clamp(., 1e-5))
//
// One CTA per (batch item, frame): the reflect-padded, windowed frame goes into shared memory in bit-reversed order,
// a radix-2 decimation-in-time FFT runs in place (same butterfly schedule as the inverse transform of the vocoder
// head, csrc/elementwise.cu), the 513 magnitudes stay in shared memory and each of the first n_mels threads takes the
// dot product of one triangular filter over ITS non-zero bin range only (the filters are sparse: ~2-60 bins each).
// HBM traffic per frame: 1 KB of new samples in (frames overlap 4x, served by L2), 4 * n_mels bytes out.
#include "common.h"
#include "ptx.cuh"

namespace lemas {

constexpr int MEL_NFFT = 1024;
constexpr int MEL_HOP = 256;
constexpr int MEL_BINS = MEL_NFFT / 2 + 1;

__global__ void __launch_bounds__(256)
mel_frames_kernel(const float* __restrict__ wav, int nw, int wav_ld, const float* __restrict__ fb,
                  const int* __restrict__ fb_range, int n_mels, int t_len, float* __restrict__ mel, int pad, float mag_eps) {
  __shared__ float2 buf[MEL_NFFT];
  __shared__ float2 tw[MEL_NFFT / 2];
  __shared__ float mag[MEL_BINS];
  const int frame = blockIdx.x % t_len;
  const int b = blockIdx.x / t_len;
  const float* w = wav + (long)b * wav_ld;
  for (int k = threadIdx.x; k < MEL_NFFT / 2; k += blockDim.x) {
    float s, c;
    sincospif((float)k * (-2.0f / MEL_NFFT), &s, &c);  // exp(-2 pi i k / N): forward transform
    tw[k] = make_float2(c, s);
  }
  for (int n = threadIdx.x; n < MEL_NFFT; n += blockDim.x) {
    int i = frame * MEL_HOP + n - pad;                   // pad 512: center=True, frame t covers [t*hop - 512, t*hop + 512);
                                                         // pad 384: bigvgan's explicit (n_fft - hop) / 2 padding, center=False
    if (i < 0) i = -i;                                   // pad_mode="reflect" (no edge repeat)
    if (i >= nw) i = 2 * (nw - 1) - i;
    const float win = 0.5f - 0.5f * cospif((float)n * (2.0f / MEL_NFFT));   // periodic hann
    buf[__brev((unsigned)n) >> 22] = make_float2(w[i] * win, 0.f);
  }
  __syncthreads();
#pragma unroll 1
  for (int len = 2; len <= MEL_NFFT; len <<= 1) {
    const int half = len >> 1;
    const int tstep = MEL_NFFT / len;
    for (int i = threadIdx.x; i < MEL_NFFT / 2; i += blockDim.x) {
      const int grp = i / half, j = i - grp * half;
      const int i0 = grp * len + j, i1 = i0 + half;
      const float2 t = tw[j * tstep];
      const float2 a = buf[i0], c = buf[i1];
      const float2 tc = make_float2(c.x * t.x - c.y * t.y, c.x * t.y + c.y * t.x);
      buf[i0] = make_float2(a.x + tc.x, a.y + tc.y);
      buf[i1] = make_float2(a.x - tc.x, a.y - tc.y);
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < MEL_BINS; k += blockDim.x) mag[k] = sqrtf(buf[k].x * buf[k].x + buf[k].y * buf[k].y + mag_eps);
  __syncthreads();
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    const int lo = fb_range[2 * m], hi = fb_range[2 * m + 1];
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc = fmaf(mag[k], __ldg(fb + (long)k * n_mels + m), acc);
    mel[((long)b * n_mels + m) * t_len + frame] = logf(fmaxf(acc, 1e-5f));
  }
}

}  // namespace lemas

using namespace lemas;

extern "C" int lemas_mel_spectrogram_1024(const float* wav, int32_t batch, int32_t nw, int32_t wav_ld, const float* fb,
                                          const int32_t* fb_range, int32_t n_mels, float* mel, void* stream) {
  LEMAS_REQUIRE(wav && fb && fb_range && mel, "lemas_mel_spectrogram_1024: null pointer");
  LEMAS_REQUIRE(batch >= 1 && n_mels >= 1 && wav_ld >= nw, "lemas_mel_spectrogram_1024: bad shape");
  LEMAS_REQUIRE(nw > MEL_NFFT / 2, "lemas_mel_spectrogram_1024: reflect padding needs more than 512 samples");
  const int t_len = nw / MEL_HOP + 1;
  mel_frames_kernel<<<batch * t_len, 256, 0, (cudaStream_t)stream>>>(wav, nw, wav_ld, fb, fb_range, n_mels, t_len, mel,
                                                                     MEL_NFFT / 2, 0.f);
  LEMAS_CUDA_OK(cudaGetLastError());
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}

// The `mel_spec_type: bigvgan` front-end (get_bigvgan_mel_spectrogram, modules.py:30-72): reflect padding of
// (n_fft - hop) / 2 = 384 samples, STFT without centering, sqrt(re^2 + im^2 + 1e-9), filterbank (Slaney mel, Slaney norm:
// what librosa.filters.mel returns, in the same [513, n_mels] layout as above), log(clamp 1e-5).
// mel: fp32 [batch, n_mels, (nw - 256) / 256 + 1].
extern "C" int lemas_mel_spectrogram_bigvgan_1024(const float* wav, int32_t batch, int32_t nw, int32_t wav_ld,
                                                  const float* fb, const int32_t* fb_range, int32_t n_mels, float* mel,
                                                  void* stream) {
  LEMAS_REQUIRE(wav && fb && fb_range && mel, "lemas_mel_spectrogram_bigvgan_1024: null pointer");
  LEMAS_REQUIRE(batch >= 1 && n_mels >= 1 && wav_ld >= nw, "lemas_mel_spectrogram_bigvgan_1024: bad shape");
  constexpr int pad = (MEL_NFFT - MEL_HOP) / 2;
  LEMAS_REQUIRE(nw > pad, "lemas_mel_spectrogram_bigvgan_1024: reflect padding needs more than 384 samples");
  const int t_len = (nw + 2 * pad - MEL_NFFT) / MEL_HOP + 1;
  mel_frames_kernel<<<batch * t_len, 256, 0, (cudaStream_t)stream>>>(wav, nw, wav_ld, fb, fb_range, n_mels, t_len, mel, pad,
                                                                     1e-9f);
  LEMAS_CUDA_OK(cudaGetLastError());
  LEMAS_LAUNCHED(1);
  return LEMAS_OK;
}
