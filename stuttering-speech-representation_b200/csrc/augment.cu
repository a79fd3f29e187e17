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
  int o1, n1;     // first pass: reduced orig / new rate
  float factor;   // noise std / gain
  unsigned long long seed;
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

extern "C" int ssr_augment(const float* audio_dev, int64_t in_stride, const int32_t* n_in, int32_t batch,
                           const ssr_aug_op* ops, int32_t sample_rate, const float* noise_dev, int64_t noise_stride,
                           void* work_dev, int64_t work_bytes, float* out_dev, int64_t out_stride, int32_t* n_out,
                           void* cuda_stream, char* err, int32_t err_len) {
  if (batch <= 0) return 0;
  if (!audio_dev || !n_in || !ops || !out_dev || !n_out) return fail(err, err_len, "ssr_augment: null argument");
  if (sample_rate <= 0) return fail(err, err_len, "ssr_augment: bad sample rate");
  cudaStream_t st = (cudaStream_t)cuda_stream;

  // host-side plan (pageable: cudaMemcpyAsync stages it before returning, so it may die with this frame)
  std::vector<AugClip> plan((size_t)batch);
  long long max_mid = 0;
  bool any_speed = false;
  for (int b = 0; b < batch; ++b) {
    AugClip& c = plan[b];
    const ssr_aug_op& op = ops[b];
    c.kind = op.kind;
    c.n_in = n_in[b];
    c.n_mid = 0;
    c.o1 = c.n1 = 1;
    c.factor = op.factor;
    c.seed = op.seed;
    if (c.n_in < 0 || c.n_in > in_stride) {
      return fail(err, err_len, "ssr_augment: n_in out of range");
    }
    if (op.kind < SSR_AUG_NONE || op.kind > SSR_AUG_VOLUME) {
      return fail(err, err_len, "ssr_augment: unknown augmentation kind");
    }
    if (op.kind == SSR_AUG_SPEED && op.new_rate == sample_rate) c.kind = SSR_AUG_NONE;  // Resample is the identity
    if (c.kind == SSR_AUG_SPEED) {
      if (op.new_rate <= 0) {
          return fail(err, err_len, "ssr_augment: bad new_rate");
      }
      const int g = gcd_i(sample_rate, op.new_rate);
      c.o1 = sample_rate / g;
      c.n1 = op.new_rate / g;
      c.n_mid = resampled_len(c.n_in, c.o1, c.n1);
      c.n_out = resampled_len(c.n_mid, c.n1, c.o1);
      if (c.n_mid > max_mid) max_mid = c.n_mid;
      any_speed = true;
    } else {
      c.n_out = c.n_in;
    }
    if (c.n_out > out_stride) {
      return fail(err, err_len, "ssr_augment: out_stride too small for the resampled length");
    }
    n_out[b] = c.n_out;
  }
  const long long mid_stride = (max_mid + 3) & ~3LL;
  const size_t plan_bytes = (sizeof(AugClip) * (size_t)batch + 255) & ~(size_t)255;
  const size_t need = plan_bytes + (any_speed ? sizeof(float) * (size_t)mid_stride * (size_t)batch : 0);
  if (!work_dev || (size_t)work_bytes < need) {
    char msg[160];
    snprintf(msg, sizeof msg, "ssr_augment: work buffer too small (%lld bytes given, %zu needed)",
             (long long)work_bytes, need);
    return fail(err, err_len, msg);
  }
  AugClip* plan_dev = reinterpret_cast<AugClip*>(work_dev);
  float* mid = reinterpret_cast<float*>(reinterpret_cast<char*>(work_dev) + plan_bytes);
  cudaError_t ce = cudaMemcpyAsync(plan_dev, plan.data(), sizeof(AugClip) * (size_t)batch, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess && any_speed) {
    dim3 grid((unsigned)((max_mid + 255) / 256), (unsigned)batch);
    aug_resample_fwd_kernel<<<grid, 256, 0, st>>>(audio_dev, in_stride, plan_dev, mid, mid_stride);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) {
    dim3 grid((unsigned)((out_stride + 255) / 256), (unsigned)batch);
    aug_finish_kernel<<<grid, 256, 0, st>>>(audio_dev, in_stride, plan_dev, mid, mid_stride, noise_dev, noise_stride,
                                            out_dev, out_stride);
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
  long long max_mid = 0;
  bool any = false;
  for (int b = 0; b < batch; ++b) {
    if (ops[b].kind == SSR_AUG_SPEED && ops[b].new_rate != sample_rate && ops[b].new_rate > 0) {
      const long long m = ssr_resample_length(n_in[b], sample_rate, ops[b].new_rate);
      if (m > max_mid) max_mid = m;
      any = true;
    }
  }
  const size_t plan_bytes = (sizeof(AugClip) * (size_t)batch + 255) & ~(size_t)255;
  const long long mid_stride = (max_mid + 3) & ~3LL;
  return (int64_t)(plan_bytes + (any ? sizeof(float) * (size_t)mid_stride * (size_t)batch : 0));
}
