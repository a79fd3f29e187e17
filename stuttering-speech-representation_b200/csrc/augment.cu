// Waveform augmentation on the device, batched: the second caller of the hot path.
//
// Replaces, for a whole batch of clips at once,
//   augment_audio            /root/reference/model_training_1.py:166-213   (speed / noise / volume / none)
//   augment_audio            /root/reference/model_training_01.py:140-192  (same kinds, wider ranges; pitch excluded)
// whose arithmetic lives in torchaudio.transforms.Resample (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99:
// torchaudio/functional/functional.py `_get_sinc_resample_kernel` / `_apply_sinc_resample_kernel`) and plain torch
// elementwise ops.  The reference materialises, per clip and per direction, a [new/gcd, 2*width + orig/gcd] filter
// bank (up to 16 000 x 16 014 fp32 = 1 GB for co-prime rates) and runs a strided conv1d over it; almost all of it is
// zero.  Here each output sample evaluates only its ~13 non-zero taps:
//
//   out[n] = sum_m  x[q*o + m] * h(m, p),   n = q*nw + p,   o = orig/gcd, nw = new/gcd
//   h(m, p) = float32( scale * sinc(pi*t) * cos^2(pi*t/12) ),  t = base*(m/o + float32(-p/nw)),  |t| < 6
//   base = 0.99*min(o, nw),  scale = base/o
//
// t is formed in fp64 with the reference's own operations in its own order (including its float32 phase); sin / cos
// advance from tap to tap by a fixed angle, so they are produced by one fp64 sincos per output sample and a plane
// rotation per tap. Tap products and the running sum are fp32, like the reference's conv1d.
//
// Memory-bound by design (HBM): per clip 2 x (read n + write n) floats for the speed round trip, 1 x for the others.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/ssr_b200.h"

namespace ssr {

struct AugClip {
  int kind;       // SSR_AUG_*
  int n_in;       // samples in
  int n_mid;      // speed: length after the first resample
  int n_out;      // samples out
  int o1, n1;     // speed: reduced orig / new rate of the first pass. pitch: reduced int(sr / rate) and sr
  float factor;   // noise std / gain
  unsigned long long seed;
  // pitch shift only
  double rate;             // 2 ** (-n_steps / 12)
  int frames, frames2;     // STFT frames before / after the phase vocoder
  int n_res;               // length after the final resample (before crop / pad to n_in)
  long long spec_off, spec2_off, fr_off, y_off;  // float offsets into the pitch scratch area
};

// ------------------------------------------------------------------------------------------------ sinc resampler
// One output sample of torchaudio's sinc_interp_hann resampler (orig rate o, new rate nw, both already divided by
// their gcd) from `x[0, n_in)`, zero outside.
__device__ __forceinline__ float resample_one(const float* __restrict__ x, int n_in, int o, int nw, long long n) {
  const long long q = n / nw;
  const int p = (int)(n - q * nw);
  const double od = (double)o;
  const double base = (double)(o < nw ? o : nw) * 0.99;  // base_freq = min(orig, new) * rolloff
  const double scale = base / od;
  // torchaudio forms the phase as `arange(0, -new, -1) / new` on an int64 tensor: a FLOAT32 true division, promoted
  // to float64 only by the following addition. That rounding is observable (~5e-5 for co-prime rates), so it is kept.
  const double phase = (double)__fdiv_rn(-(float)p, (float)nw);
  const double half = 6.0 * od / base;
  const double centre = (double)((long long)p * o) / (double)nw;
  long long m_lo = (long long)floor(centre - half) - 1;
  long long m_hi = (long long)ceil(centre + half) + 1;
  const long long base_idx = q * o;
  if (base_idx + m_lo < 0) m_lo = -base_idx;
  if (base_idx + m_hi > (long long)n_in - 1) m_hi = (long long)n_in - 1 - base_idx;
  if (m_lo > m_hi) return 0.f;

  const double a0 = ((phase + (double)m_lo / od) * base) * M_PI;  // pi * t at the first tap
  const double da = (base / od) * M_PI;                           // per-tap advance
  double s, c, ds, dc, sw, cw, dsw, dcw;
  sincos(a0, &s, &c);
  sincos(da, &ds, &dc);
  sincos(a0 / 6.0, &sw, &cw);  // window: cos^2(pi t / 12) = (1 + cos(pi t / 6)) / 2
  sincos(da / 6.0, &dsw, &dcw);
  float acc = 0.f;
  for (long long m = m_lo; m <= m_hi; ++m) {
    const double t = (phase + (double)m / od) * base;  // the reference's own expression, operation for operation
    if (fabs(t) < 6.0) {
      const double tp = t * M_PI;
      // sin(tp) comes from the rotated phasor; next to zero the series is used (the phasor's absolute error would
      // otherwise be divided by a tiny number)
      const double tp2 = tp * tp;
      const double sinc = fabs(tp) < 1e-2 ? 1.0 - tp2 * (1.0 / 6.0) + tp2 * tp2 * (1.0 / 120.0) : s / tp;
      const double win = 0.5 + 0.5 * cw;
      const float h = (float)(sinc * (win * scale));
      acc = fmaf(x[base_idx + m], h, acc);
    }
    // rotate both phasors by one tap
    const double s2 = s * dc + c * ds, c2 = c * dc - s * ds;
    s = s2;
    c = c2;
    const double sw2 = sw * dcw + cw * dsw, cw2 = cw * dcw - sw * dsw;
    sw = sw2;
    cw = cw2;
  }
  return acc;
}

// The same resampler with the filter bank evaluated in FLOAT32 throughout, operation for operation as torchaudio
// does when `_get_sinc_resample_kernel` is given dtype=float32 — which is what transforms.PitchShift does
// (initialize_parameters passes dtype=input.dtype). No fused multiply-adds in the tap weight.
__device__ __forceinline__ float resample_one_f32(const float* __restrict__ x, int n_in, int o, int nw, long long n) {
  const long long q = n / nw;
  const int p = (int)(n - q * nw);
  const double based = (double)(o < nw ? o : nw) * 0.99;
  const float base = (float)based;
  const float scale = (float)(based / (double)o);
  const float pi = 3.14159274101257324f;
  const float phase = __fdiv_rn(-(float)p, (float)nw);
  const double half = 6.0 * (double)o / based;
  const double centre = (double)((long long)p * o) / (double)nw;
  long long m_lo = (long long)floor(centre - half) - 2;
  long long m_hi = (long long)ceil(centre + half) + 2;
  const long long base_idx = q * o;
  if (base_idx + m_lo < 0) m_lo = -base_idx;
  if (base_idx + m_hi > (long long)n_in - 1) m_hi = (long long)n_in - 1 - base_idx;
  float acc = 0.f;
  for (long long m = m_lo; m <= m_hi; ++m) {
    const float t = __fmul_rn(__fadd_rn(phase, __fdiv_rn((float)m, (float)o)), base);
    if (fabsf(t) < 6.0f) {
      const float cw = cosf(__fdiv_rn(__fdiv_rn(__fmul_rn(t, pi), 6.0f), 2.0f));
      const float window = __fmul_rn(cw, cw);
      const float tp = __fmul_rn(t, pi);
      const float sinc = tp == 0.f ? 1.0f : __fdiv_rn(sinf(tp), tp);
      const float h = __fmul_rn(sinc, __fmul_rn(window, scale));
      acc = fmaf(x[base_idx + m], h, acc);
    }
  }
  return acc;
}

// ------------------------------------------------------------------------------------------------ pitch shift
// torchaudio.transforms.PitchShift(sr, n_steps) (REF/model_training_01.py:174-178): STFT (512, hop 128, periodic
// hann, centred, reflect) -> phase vocoder at rate 2^(-n_steps/12) -> inverse STFT to round(len / rate) samples ->
// float32-kernel resample int(sr / rate) -> sr -> crop / zero-pad to the input length.
// The phase vocoder accumulates each bin's phase over ALL frames, including frames where the bin holds nothing but
// rounding noise; for tonal input the reference's own output therefore depends on FFT rounding at the percent level
// (measured: two correct float32 STFTs differing by 2e-7 relative change its output by 7e-2 on a unit chirp). Parity
// is stated accordingly: tight on broadband input, statistical on tonal input (see tests/test_augment_gpu.py).
constexpr int PS_NFFT = 512, PS_HOP = 128, PS_BINS = 257;

// In-place radix-2 decimation-in-time FFT of 512 complex points in shared memory by 256 threads (one butterfly each
// per stage). `s` must hold the input in bit-reversed order; tw[k] = exp(-2 pi i k / 512), k < 256; INVERSE conjugates
// the twiddles (no 1/N). Ends with a barrier.
template <bool INVERSE>
__device__ __forceinline__ void fft512(float2* s, const float2* tw) {
#pragma unroll 1
  for (int half = 1; half < PS_NFFT; half <<= 1) {
    const int pos = threadIdx.x & (half - 1);
    const int i = ((threadIdx.x - pos) << 1) + pos, j = i + half;
    float2 w = tw[pos * (PS_NFFT / 2 / half)];
    if (INVERSE) w.y = -w.y;
    const float2 a = s[i], b = s[j];
    const float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
    s[i] = make_float2(a.x + bw.x, a.y + bw.y);
    s[j] = make_float2(a.x - bw.x, a.y - bw.y);
    __syncthreads();
  }
}
__device__ __forceinline__ int brev9(int n) { return (int)(__brev((unsigned)n) >> 23); }

// grid (max frames, B), 256 threads: one windowed, reflect-padded frame -> 257 bins (frame-major complex output)
__global__ void __launch_bounds__(256) pitch_stft_kernel(const float* __restrict__ in, long long in_stride,
                                                         const AugClip* __restrict__ clips,
                                                         float* __restrict__ scratch) {
  const AugClip cl = clips[blockIdx.y];
  const int f = blockIdx.x;
  if (cl.kind != SSR_AUG_PITCH || f >= cl.frames) return;
  __shared__ float2 s[PS_NFFT];
  __shared__ float2 tw[PS_NFFT / 2];
  const float* x = in + (long long)blockIdx.y * in_stride;
  {
    float sn, cs;
    sincospif((float)threadIdx.x * (1.0f / 256.0f), &sn, &cs);  // angle 2 pi k / 512
    tw[threadIdx.x] = make_float2(cs, -sn);
  }
  for (int n = threadIdx.x; n < PS_NFFT; n += 256) {
    long long i = (long long)f * PS_HOP + n - PS_NFFT / 2;
    if (i < 0) i = -i;
    if (i >= cl.n_in) i = 2LL * (cl.n_in - 1) - i;
    const float w = 0.5f - 0.5f * cospif((float)n * (1.0f / 256.0f));  // periodic hann
    s[brev9(n)] = make_float2(x[i] * w, 0.f);
  }
  __syncthreads();
  fft512<false>(s, tw);
  float2* spec = reinterpret_cast<float2*>(scratch + cl.spec_off) + (long long)f * PS_BINS;
  for (int k = threadIdx.x; k < PS_BINS; k += 256) spec[k] = s[k];
}

// grid (ceil(257 / 64), B), 64 threads: one thread walks one frequency bin through time (the phase accumulates)
__global__ void __launch_bounds__(64) pitch_vocoder_kernel(const AugClip* __restrict__ clips,
                                                           float* __restrict__ scratch) {
  const AugClip cl = clips[blockIdx.y];
  const int k = blockIdx.x * 64 + threadIdx.x;
  if (cl.kind != SSR_AUG_PITCH || k >= PS_BINS) return;
  const float2* spec = reinterpret_cast<const float2*>(scratch + cl.spec_off);
  float2* out = reinterpret_cast<float2*>(scratch + cl.spec2_off);
  // torch.linspace(0, pi * hop, 257) in float32: evaluated from both ends towards the middle
  const float end = (float)(3.14159265358979323846 * PS_HOP);
  const float step = __fdiv_rn(end, (float)(PS_BINS - 1));
  const float pa = k < PS_BINS / 2 ? __fmul_rn(step, (float)k) : __fsub_rn(end, __fmul_rn(step, (float)(PS_BINS - 1 - k)));
  const float two_pi = 6.28318548202514648f;
  const float2 first = spec[k];
  double acc = 0.0;
  float prev = atan2f(first.y, first.x);  // phase_0
  for (int i = 0; i < cl.frames2; ++i) {
    const float ts = (float)(cl.rate * (double)i);
    const long long i0 = (long long)ts, i1 = (long long)__fadd_rn(ts, 1.0f);
    const float alpha = fmodf(ts, 1.0f);
    const float2 s0 = i0 < cl.frames ? spec[i0 * PS_BINS + k] : make_float2(0.f, 0.f);
    const float2 s1 = i1 < cl.frames ? spec[i1 * PS_BINS + k] : make_float2(0.f, 0.f);
    const float a0 = atan2f(s0.y, s0.x), a1 = atan2f(s1.y, s1.x);
    const float n0 = hypotf(s0.x, s0.y), n1 = hypotf(s1.x, s1.y);
    float ph = __fsub_rn(__fsub_rn(a1, a0), pa);
    ph = __fsub_rn(ph, __fmul_rn(two_pi, rintf(__fdiv_rn(ph, two_pi))));
    ph = __fadd_rn(ph, pa);
    acc += (double)prev;  // cumsum of cat([phase_0, phase[:-1]]), accumulated in double like torch's CPU cumsum
    prev = ph;
    const float pacc = (float)acc;
    const float mag = __fadd_rn(__fmul_rn(alpha, n1), __fmul_rn(__fsub_rn(1.0f, alpha), n0));
    float s, c;
    sincosf(pacc, &s, &c);
    out[(long long)i * PS_BINS + k] = make_float2(mag * c, mag * s);
  }
}

// grid (ceil(max n_mid / 128), B), 128 threads: overlap-add of the (at most 4) windowed frames covering a sample,
// divided by the window envelope (torch.istft, centred, length = n_mid); frames come from pitch_iframe_kernel
__global__ void __launch_bounds__(128) pitch_istft_kernel(const AugClip* __restrict__ clips,
                                                          float* __restrict__ scratch) {
  const AugClip cl = clips[blockIdx.y];
  const long long n0 = (long long)blockIdx.x * 128;
  if (cl.kind != SSR_AUG_PITCH || n0 >= cl.n_mid) return;
  const float* fr = scratch + cl.fr_off;
  float* y = scratch + cl.y_off;
  const int q = blockIdx.x;
  const long long n = n0 + threadIdx.x;
  if (n >= cl.n_mid) return;
  const long long pos = n + PS_NFFT / 2;
  float ysum = 0.f, env = 0.f;
#pragma unroll
  for (int fi = 0; fi < 4; ++fi) {  // fixed order: bit-reproducible
    const int f = q - 1 + fi;
    if (f < 0 || f >= cl.frames2) continue;
    const int j = (int)(pos - (long long)f * PS_HOP);  // 0 .. 511 by construction
    const float wj = 0.5f - 0.5f * cospif((float)j * (1.0f / 256.0f));
    ysum += fr[(long long)f * PS_NFFT + j];  // already windowed
    env = fmaf(wj, wj, env);
  }
  y[n] = env > 1e-11f ? ysum / env : 0.f;  // beyond the last frame: zero padding
}

// grid (max frames2, B), 256 threads: one vocoder frame -> 512 windowed time samples (irfft of the Hermitian
// extension, 1/N, times the synthesis window)
__global__ void __launch_bounds__(256) pitch_iframe_kernel(const AugClip* __restrict__ clips,
                                                           float* __restrict__ scratch) {
  const AugClip cl = clips[blockIdx.y];
  const int f = blockIdx.x;
  if (cl.kind != SSR_AUG_PITCH || f >= cl.frames2) return;
  __shared__ float2 s[PS_NFFT];
  __shared__ float2 tw[PS_NFFT / 2];
  const float2* X = reinterpret_cast<const float2*>(scratch + cl.spec2_off) + (long long)f * PS_BINS;
  {
    float sn, cs;
    sincospif((float)threadIdx.x * (1.0f / 256.0f), &sn, &cs);
    tw[threadIdx.x] = make_float2(cs, -sn);
  }
  for (int k = threadIdx.x; k < PS_BINS; k += 256) {
    float2 v = X[k];
    if (k == 0 || k == PS_BINS - 1) v.y = 0.f;  // a real signal's DC / Nyquist bins are real (c2r ignores the rest)
    s[brev9(k)] = v;
    if (k > 0 && k < PS_BINS - 1) s[brev9(PS_NFFT - k)] = make_float2(v.x, -v.y);
  }
  __syncthreads();
  fft512<true>(s, tw);
  float* fr = scratch + cl.fr_off + (long long)f * PS_NFFT;
  for (int j = threadIdx.x; j < PS_NFFT; j += 256) {
    const float wj = 0.5f - 0.5f * cospif((float)j * (1.0f / 256.0f));
    fr[j] = s[j].x * (1.0f / PS_NFFT) * wj;
  }
}

// ------------------------------------------------------------------------------------------------ counter RNG
// Philox4x32-10 keyed by the op's seed (unique per clip, chosen by the host), counter = sample index / 4: any
// sharding or batching of the clips reproduces the same noise. Box-Muller on the four words gives four normals.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
  k[0] += 0x9E3779B9u;
  k[1] += 0xBB67AE85u;
}

__device__ __forceinline__ void philox_normal4(unsigned long long seed, uint32_t clip, uint32_t block4, float (&z)[4]) {
  uint32_t c[4] = {block4, clip, 0x5353525Fu, 0x42323030u};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
  const float u0 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u1 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(c[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u3 = ((float)(c[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  z[0] = r0 * c0;
  z[1] = r0 * s0;
  z[2] = r1 * c1;
  z[3] = r1 * s1;
}

__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.0f), 1.0f); }

// ------------------------------------------------------------------------------------------------ kernels
// Pass 1 (speed clips only): x -> mid at the perturbed rate.
__global__ void __launch_bounds__(256) aug_resample_fwd_kernel(const float* __restrict__ in, long long in_stride,
                                                               const AugClip* __restrict__ clips,
                                                               float* __restrict__ mid, long long mid_stride) {
  const AugClip cl = clips[blockIdx.y];
  if (cl.kind != SSR_AUG_SPEED) return;
  const float* x = in + (long long)blockIdx.y * in_stride;
  float* y = mid + (long long)blockIdx.y * mid_stride;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < cl.n_mid;
       n += (long long)gridDim.x * blockDim.x)
    y[n] = resample_one(x, cl.n_in, cl.o1, cl.n1, n);
}

// Pass 2: speed clips resample back (mid -> out at the original rate); the other kinds are elementwise. The final
// clamp to [-1, 1] (model_training_1.py:204) is fused. Tail samples up to out_stride are zero-filled so that the
// result can go to the encoder as a padded ragged batch.
__global__ void __launch_bounds__(256) aug_finish_kernel(const float* __restrict__ in, long long in_stride,
                                                         const AugClip* __restrict__ clips,
                                                         const float* __restrict__ mid, long long mid_stride,
                                                         const float* __restrict__ noise, long long noise_stride,
                                                         const float* __restrict__ pscratch,
                                                         float* __restrict__ out, long long out_stride) {
  const AugClip cl = clips[blockIdx.y];
  const float* x = in + (long long)blockIdx.y * in_stride;
  float* y = out + (long long)blockIdx.y * out_stride;
  const long long step = (long long)gridDim.x * blockDim.x;
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cl.kind == SSR_AUG_SPEED) {
    const float* m = mid + (long long)blockIdx.y * mid_stride;
    for (long long n = first; n < out_stride; n += step)
      y[n] = n < cl.n_out ? clamp1(resample_one(m, cl.n_mid, cl.n1, cl.o1, n)) : 0.f;
  } else if (cl.kind == SSR_AUG_PITCH) {
    // stretched signal -> resample int(sr / rate) -> sr, cropped / zero-padded to the input length
    const float* m = pscratch + cl.y_off;
    for (long long n = first; n < out_stride; n += step) {
      float v = 0.f;
      if (n < cl.n_out && n < cl.n_res) v = cl.o1 == cl.n1 ? m[n] : resample_one_f32(m, cl.n_mid, cl.o1, cl.n1, n);
      y[n] = clamp1(v);
    }
  } else if (cl.kind == SSR_AUG_NOISE && noise != nullptr) {
    // caller-supplied standard-normal noise (the reference's torch.randn_like stream): bit-exact replay of
    // `waveform + noise * noise_factor` (two roundings, no fused multiply-add)
    const float* z = noise + (long long)blockIdx.y * noise_stride;
    for (long long n = first; n < out_stride; n += step)
      y[n] = n < cl.n_out ? clamp1(__fadd_rn(x[n], __fmul_rn(z[n], cl.factor))) : 0.f;
  } else if (cl.kind == SSR_AUG_NOISE) {
    for (long long n4 = first; n4 * 4 < out_stride; n4 += step) {
      float z[4];
      philox_normal4(cl.seed, (uint32_t)(n4 >> 32), (uint32_t)n4, z);  // keyed by the op's seed only
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long n = n4 * 4 + i;
        if (n < out_stride) y[n] = n < cl.n_out ? clamp1(__fadd_rn(x[n], __fmul_rn(z[i], cl.factor))) : 0.f;
      }
    }
  } else {
    const float g = cl.kind == SSR_AUG_VOLUME ? cl.factor : 1.0f;
    for (long long n = first; n < out_stride; n += step) y[n] = n < cl.n_out ? clamp1(__fmul_rn(x[n], g)) : 0.f;
  }
}

static int gcd_i(int a, int b) {
  while (b) {
    const int t = a % b;
    a = b;
    b = t;
  }
  return a;
}

// torchaudio: target_length = ceil(float32(new * length / orig)) with the quotient formed in double precision.
static int resampled_len(long long n, int o, int nw) {
  const float f = (float)((double)nw * (double)n / (double)o);
  return (int)ceilf(f);
}

static int fail(char* err, int err_len, const char* msg) {
  if (err && err_len > 0) snprintf(err, (size_t)err_len, "%s", msg);
  return -1;
}

}  // namespace ssr

using namespace ssr;

extern "C" int32_t ssr_resample_length(int32_t n, int32_t orig_rate, int32_t new_rate) {
  if (n < 0 || orig_rate <= 0 || new_rate <= 0) return -1;
  if (orig_rate == new_rate) return n;
  const int g = gcd_i(orig_rate, new_rate);
  return resampled_len(n, orig_rate / g, new_rate / g);
}

extern "C" int32_t ssr_augment_out_length(const ssr_aug_op* op, int32_t n, int32_t sample_rate) {
  if (op == nullptr || n < 0 || sample_rate <= 0) return -1;
  if (op->kind != SSR_AUG_SPEED || op->new_rate == sample_rate) return n;
  if (op->new_rate <= 0) return -1;
  return ssr_resample_length(ssr_resample_length(n, sample_rate, op->new_rate), op->new_rate, sample_rate);
}

namespace ssr {

struct AugPlan {
  std::vector<AugClip> clips;
  long long max_mid = 0;      // speed: longest intermediate signal
  long long mid_stride = 0;
  long long max_frames = 0;   // pitch: most STFT frames (before / after the vocoder) / longest stretched signal
  long long max_frames2 = 0;
  long long max_stretch = 0;
  long long pitch_floats = 0; // pitch scratch, in floats
  bool any_speed = false, any_pitch = false;
  size_t plan_bytes = 0, mid_bytes = 0, need = 0;
};

// Host-side plan shared by ssr_augment and ssr_augment_work_bytes. Returns an error text or nullptr.
static const char* build_plan(const int32_t* n_in, int batch, const ssr_aug_op* ops, int sample_rate,
                              long long in_stride, AugPlan& P) {
  P.clips.assign((size_t)batch, AugClip());
  for (int b = 0; b < batch; ++b) {
    AugClip& c = P.clips[b];
    const ssr_aug_op& op = ops[b];
    memset(&c, 0, sizeof(c));
    c.kind = op.kind;
    c.n_in = n_in[b];
    c.o1 = c.n1 = 1;
    c.factor = op.factor;
    c.seed = op.seed;
    if (c.n_in < 0 || (in_stride >= 0 && c.n_in > in_stride)) return "ssr_augment: n_in out of range";
    if (op.kind < SSR_AUG_NONE || op.kind > SSR_AUG_PITCH) return "ssr_augment: unknown augmentation kind";
    if (op.kind == SSR_AUG_SPEED && op.new_rate == sample_rate) c.kind = SSR_AUG_NONE;  // Resample is the identity
    if (op.kind == SSR_AUG_PITCH && op.new_rate == 0) c.kind = SSR_AUG_NONE;            // n_steps == 0
    c.n_out = c.n_in;
    if (c.kind == SSR_AUG_SPEED) {
      if (op.new_rate <= 0) return "ssr_augment: bad new_rate";
      const int g = gcd_i(sample_rate, op.new_rate);
      c.o1 = sample_rate / g;
      c.n1 = op.new_rate / g;
      c.n_mid = resampled_len(c.n_in, c.o1, c.n1);
      c.n_out = resampled_len(c.n_mid, c.n1, c.o1);
      if (c.n_mid > P.max_mid) P.max_mid = c.n_mid;
      P.any_speed = true;
    } else if (c.kind == SSR_AUG_PITCH) {
      if (op.new_rate < -24 || op.new_rate > 24) return "ssr_augment: pitch n_steps out of range";
      if (c.n_in <= PS_NFFT / 2) return "ssr_augment: clip too short for the pitch kind (reflect padding needs > 256 samples)";
      c.rate = pow(2.0, -(double)op.new_rate / 12.0);
      c.frames = 1 + c.n_in / PS_HOP;
      c.frames2 = (int)ceil((double)c.frames / c.rate);
      c.n_mid = (int)nearbyint((double)c.n_in / c.rate);  // python round(): ties to even
      const int orig = (int)((double)sample_rate / c.rate);
      const int g = gcd_i(orig, sample_rate);
      c.o1 = orig / g;
      c.n1 = sample_rate / g;
      c.n_res = orig == sample_rate ? c.n_mid : resampled_len(c.n_mid, c.o1, c.n1);
      c.spec_off = P.pitch_floats;
      P.pitch_floats += 2LL * c.frames * PS_BINS;
      c.spec2_off = P.pitch_floats;
      P.pitch_floats += 2LL * c.frames2 * PS_BINS;
      c.fr_off = P.pitch_floats;
      P.pitch_floats += (long long)c.frames2 * PS_NFFT;
      c.y_off = P.pitch_floats;
      P.pitch_floats += ((long long)c.n_mid + 3) & ~3LL;
      if (c.frames > P.max_frames) P.max_frames = c.frames;
      if (c.frames2 > P.max_frames2) P.max_frames2 = c.frames2;
      if (c.n_mid > P.max_stretch) P.max_stretch = c.n_mid;
      P.any_pitch = true;
    }
  }
  P.mid_stride = (P.max_mid + 3) & ~3LL;
  P.plan_bytes = (sizeof(AugClip) * (size_t)batch + 255) & ~(size_t)255;
  P.mid_bytes = P.any_speed ? ((sizeof(float) * (size_t)P.mid_stride * (size_t)batch + 255) & ~(size_t)255) : 0;
  P.need = P.plan_bytes + P.mid_bytes + sizeof(float) * (size_t)P.pitch_floats;
  return nullptr;
}

}  // namespace ssr

extern "C" int ssr_augment(const float* audio_dev, int64_t in_stride, const int32_t* n_in, int32_t batch,
                           const ssr_aug_op* ops, int32_t sample_rate, const float* noise_dev, int64_t noise_stride,
                           void* work_dev, int64_t work_bytes, float* out_dev, int64_t out_stride, int32_t* n_out,
                           void* cuda_stream, char* err, int32_t err_len) {
  if (batch <= 0) return 0;
  if (!audio_dev || !n_in || !ops || !out_dev || !n_out) return fail(err, err_len, "ssr_augment: null argument");
  if (sample_rate <= 0) return fail(err, err_len, "ssr_augment: bad sample rate");
  cudaStream_t st = (cudaStream_t)cuda_stream;

  // host-side plan (pageable: cudaMemcpyAsync stages it before returning, so it may die with this frame)
  AugPlan P;
  if (const char* msg = build_plan(n_in, batch, ops, sample_rate, in_stride, P)) return fail(err, err_len, msg);
  for (int b = 0; b < batch; ++b) {
    if (P.clips[b].n_out > out_stride)
      return fail(err, err_len, "ssr_augment: out_stride too small for the resampled length");
    n_out[b] = P.clips[b].n_out;
  }
  if (!work_dev || (size_t)work_bytes < P.need) {
    char msg[160];
    snprintf(msg, sizeof msg, "ssr_augment: work buffer too small (%lld bytes given, %zu needed)",
             (long long)work_bytes, P.need);
    return fail(err, err_len, msg);
  }
  AugClip* plan_dev = reinterpret_cast<AugClip*>(work_dev);
  float* mid = reinterpret_cast<float*>(reinterpret_cast<char*>(work_dev) + P.plan_bytes);
  float* pscratch = reinterpret_cast<float*>(reinterpret_cast<char*>(work_dev) + P.plan_bytes + P.mid_bytes);
  cudaError_t ce =
      cudaMemcpyAsync(plan_dev, P.clips.data(), sizeof(AugClip) * (size_t)batch, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess && P.any_speed) {
    dim3 grid((unsigned)((P.max_mid + 255) / 256), (unsigned)batch);
    aug_resample_fwd_kernel<<<grid, 256, 0, st>>>(audio_dev, in_stride, plan_dev, mid, P.mid_stride);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess && P.any_pitch) {
    pitch_stft_kernel<<<dim3((unsigned)P.max_frames, (unsigned)batch), 256, 0, st>>>(audio_dev, in_stride, plan_dev,
                                                                                     pscratch);
    pitch_vocoder_kernel<<<dim3((PS_BINS + 63) / 64, (unsigned)batch), 64, 0, st>>>(plan_dev, pscratch);
    pitch_iframe_kernel<<<dim3((unsigned)P.max_frames2, (unsigned)batch), 256, 0, st>>>(plan_dev, pscratch);
    pitch_istft_kernel<<<dim3((unsigned)((P.max_stretch + 127) / 128), (unsigned)batch), 128, 0, st>>>(plan_dev,
                                                                                                      pscratch);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) {
    dim3 grid((unsigned)((out_stride + 255) / 256), (unsigned)batch);
    aug_finish_kernel<<<grid, 256, 0, st>>>(audio_dev, in_stride, plan_dev, mid, P.mid_stride, noise_dev,
                                            noise_stride, pscratch, out_dev, out_stride);
    ce = cudaGetLastError();
  }
  if (ce != cudaSuccess) {
    char msg[160];
    snprintf(msg, sizeof msg, "ssr_augment: %s", cudaGetErrorString(ce));
    return fail(err, err_len, msg);
  }
  return 0;
}

extern "C" int64_t ssr_augment_work_bytes(const int32_t* n_in, int32_t batch, const ssr_aug_op* ops,
                                          int32_t sample_rate) {
  if (batch <= 0) return 0;
  if (!n_in || !ops || sample_rate <= 0) return -1;
  AugPlan P;
  if (build_plan(n_in, batch, ops, sample_rate, -1, P) != nullptr) return -1;
  return (int64_t)P.need;
}
