// Front ends (HBM / CUDA-core bound stages).
//  * WavLM: waveform zero-mean/unit-variance (HF/models/wav2vec2/feature_extraction_wav2vec2.py:77-97) fused into
//    conv layer 0 (C_in = 1, k = 10, stride 5, no bias) + {LayerNorm over 512 channels | GroupNorm over time} +
//    erf-GELU (HF/models/wavlm/modeling_wavlm.py:682-751).
//  * Whisper: log-mel spectrogram (HF/models/whisper/feature_extraction_whisper.py:135-164): reflect-padded
//    400-point hann STFT as an fp32 register-tiled DFT, power, slaney mel filterbank, log10, per-clip max clamp,
//    (x + 4) / 4.  Frames that lie completely in the zero padding are never computed (their power is exactly 0).
#include "common.cuh"
#include "kernels.cuh"

// wavlm_conv0_kernel<MODE> holds the LayerNorm and the GroupNorm path; one of them is dead per instantiation.
#pragma nv_diag_suppress 128

namespace ssr {

namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------ waveform statistics
__global__ void __launch_bounds__(512)
wave_stats_kernel(const float* __restrict__ audio, long long ld, const int* __restrict__ n_samples,
                  float* __restrict__ stats, int do_normalize) {
  const int b = blockIdx.x;
  const int n = n_samples[b];
  __shared__ double red[16];
  __shared__ double s_mean;
  if (!do_normalize || n <= 0) {
    if (threadIdx.x == 0) {
      stats[2 * b] = 0.f;
      stats[2 * b + 1] = 1.f;
    }
    return;
  }
  const float* x = audio + (long long)b * ld;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 16; ++i) t += red[i];
    s_mean = t / (double)n;
  }
  __syncthreads();
  const double mean = s_mean;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)x[i] - mean;
    q += d * d;
  }
  for (int o = 16; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 16; ++i) t += red[i];
    const double var = t / (double)n;  // population variance, as numpy .var()
    stats[2 * b] = (float)mean;
    stats[2 * b + 1] = (float)(1.0 / sqrt(var + 1e-7));
  }
}

// ------------------------------------------------------------------------------------------ WavLM conv layer 0
// MODE 0: LayerNorm over channels + GELU   (feat_extract_norm = "layer")
// MODE 1: accumulate per-(clip, channel) sum / sum-of-squares over valid frames (GroupNorm pass 1)
// MODE 2: GroupNorm(512 groups) apply + GELU                                         (feat_extract_norm = "group")
// Block: 8 warps x 8 frames; a warp computes 4 frames at a time, a lane owns channel pairs {2*lane + 64*i}.
constexpr int C0_FRAMES = 64;

template <int MODE>
__global__ void __launch_bounds__(256)
wavlm_conv0_kernel(const Conv0Args a) {
  __shared__ __align__(16) float w_s[10 * 512];
  __shared__ float x_s[C0_FRAMES * 5 + 8];
  __shared__ float acc_s[(MODE == 1) ? 2 * 512 : 1];
  const int b = blockIdx.y;
  const int t_blk = blockIdx.x * C0_FRAMES;
  const int n = a.n_samples[b];
  const int T0 = n >= 10 ? (n - 10) / 5 + 1 : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < 5120; i += 256) {
    const int c = i / 10, k = i - c * 10;
    w_s[k * 512 + c] = a.w[i];
  }
  const float mean = a.stats[2 * b], rstd = a.stats[2 * b + 1];
  const float* x = a.audio + (long long)b * a.audio_ld;
  for (int i = threadIdx.x; i < C0_FRAMES * 5 + 5; i += 256) {
    const int idx = t_blk * 5 + i;
    x_s[i] = (idx < n) ? (x[idx] - mean) * rstd : 0.f;
  }
  if (MODE == 1)
    for (int i = threadIdx.x; i < 1024; i += 256) acc_s[i] = 0.f;
  __syncthreads();

  float gn_s[16], gn_q[16];
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) gn_s[i] = gn_q[i] = 0.f;
  }

#pragma unroll 1
  for (int round = 0; round < 2; ++round) {
    const int tl = warp * 8 + round * 4;  // first local frame of this group of 4
    float xv[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) xv[i] = x_s[tl * 5 + i];
    float acc[8][2][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][0][j] = acc[i][1][j] = 0.f;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 w = *reinterpret_cast<const float2*>(&w_s[k * 512 + 2 * lane + 64 * i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][0][j] = fmaf(w.x, xv[5 * j + k], acc[i][0][j]);
          acc[i][1][j] = fmaf(w.y, xv[5 * j + k], acc[i][1][j]);
        }
      }
    }
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (t_blk + tl + j < T0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            gn_s[2 * i] += acc[i][0][j];
            gn_s[2 * i + 1] += acc[i][1][j];
            gn_q[2 * i] += acc[i][0][j] * acc[i][0][j];
            gn_q[2 * i + 1] += acc[i][1][j] * acc[i][1][j];
          }
        }
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t_blk + tl + j;
      if (t >= a.slot0) continue;  // warp-uniform
      float mu = 0.f, rs = 1.f;
      if (MODE == 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += acc[i][0][j] + acc[i][1][j];
        mu = warp_sum_f(s) * (1.0f / 512.f);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d0 = acc[i][0][j] - mu, d1 = acc[i][1][j] - mu;
          q += d0 * d0 + d1 * d1;
        }
        rs = rsqrtf(warp_sum_f(q) * (1.0f / 512.f) + 1e-5f);
      }
      bf16* dst = a.out + ((long long)b * a.slot0 + t) * 512;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = 2 * lane + 64 * i;
        float y0, y1;
        if (MODE == 0) {
          const float2 g = __ldg(reinterpret_cast<const float2*>(a.gamma + c));
          const float2 be = __ldg(reinterpret_cast<const float2*>(a.beta + c));
          y0 = (acc[i][0][j] - mu) * rs * g.x + be.x;
          y1 = (acc[i][1][j] - mu) * rs * g.y + be.y;
        } else {
          // per-channel statistics over the clip's valid frames
          const double inv = T0 > 0 ? 1.0 / (double)T0 : 0.0;
          const double* ga = a.gn_acc + ((long long)b * 512 + c) * 2;
          const double m0 = ga[0] * inv, m1 = ga[2] * inv;
          const double v0 = fmax(ga[1] * inv - m0 * m0, 0.0), v1 = fmax(ga[3] * inv - m1 * m1, 0.0);
          const float r0 = (float)(1.0 / sqrt(v0 + 1e-5)), r1 = (float)(1.0 / sqrt(v1 + 1e-5));
          const float2 g = __ldg(reinterpret_cast<const float2*>(a.gamma + c));
          const float2 be = __ldg(reinterpret_cast<const float2*>(a.beta + c));
          y0 = (acc[i][0][j] - (float)m0) * r0 * g.x + be.x;
          y1 = (acc[i][1][j] - (float)m1) * r1 * g.y + be.y;
        }
        __nv_bfloat162 p = __floats2bfloat162_rn(gelu_fast(y0), gelu_fast(y1));
        *reinterpret_cast<__nv_bfloat162*>(dst + c) = p;
      }
    }
  }
  if (MODE == 1) {
    // fixed warp order -> the block partial is deterministic
    for (int w = 0; w < 8; ++w) {
      if (warp == w) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = 2 * lane + 64 * i;
          acc_s[2 * c + 0] += gn_s[2 * i];
          acc_s[2 * c + 1] += gn_q[2 * i];
          acc_s[2 * c + 2] += gn_s[2 * i + 1];
          acc_s[2 * c + 3] += gn_q[2 * i + 1];
        }
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < 1024; i += 256)
      atomicAdd(a.gn_acc + (long long)b * 1024 + i, (double)acc_s[i]);
  }
}

}  // namespace

int launch_wavlm_conv0(const Conv0Args& a, cudaStream_t st, std::string& err) {
  if (a.B <= 0) return 0;
  wave_stats_kernel<<<a.B, 512, 0, st>>>(a.audio, a.audio_ld, a.n_samples, a.stats, a.do_normalize);
  dim3 grid(ceil_div(a.slot0, C0_FRAMES), a.B);
  if (a.mode == 0) {
    wavlm_conv0_kernel<0><<<grid, 256, 0, st>>>(a);
  } else {
    cudaMemsetAsync(a.gn_acc, 0, sizeof(double) * 1024 * a.B, st);
    wavlm_conv0_kernel<1><<<grid, 256, 0, st>>>(a);
    wavlm_conv0_kernel<2><<<grid, 256, 0, st>>>(a);
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("wavlm conv0 launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

// ============================================================================================== Whisper log-mel
namespace {

constexpr int NFFT = 400, HOP = 160, NBIN = 201, NMEL = 80, NFRAMES = 3000, NSAMP = 480000;
constexpr int TW_LD = 448;          // padded twiddle columns (2 * 201 = 402 real)
constexpr int LM_FRAMES = 64;       // frames per CTA
constexpr int P_LD = 225;           // 224 bins computed (201 real), +1 pad -> conflict-free column walks
constexpr int LM_SMEM = (NFFT * LM_FRAMES + LM_FRAMES * P_LD) * 4;

__device__ __forceinline__ unsigned int float_to_ordered(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ int live_frames(int n) {
  if (n <= 0) return 0;
  if (n > NSAMP) n = NSAMP;
  int f = (n + 200 + HOP - 1) / HOP;  // frames whose window starts before sample n
  return f < NFRAMES ? f : NFRAMES;
}

// DFT (windowed, via the pre-windowed twiddle table) + power + mel + log10 for 64 frames of one clip.
__global__ void __launch_bounds__(256)
logmel_dft_kernel(const LogMelArgs a) {
  extern __shared__ __align__(16) float lm_smem[];
  float* fT = lm_smem;                      // [400][64] frame samples, transposed
  float* P = lm_smem + NFFT * LM_FRAMES;    // [64][225] power spectrum
  __shared__ float s_max[8];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * LM_FRAMES;
  int n = a.n_samples[b];
  if (n > NSAMP) n = NSAMP;
  const int nlive = live_frames(n);
  if (f0 >= nlive) return;
  const float* x = a.audio + (long long)b * a.audio_ld;

  // samples f0*160 - 200 ... : frame fr, tap k -> padded index (f0 + fr) * 160 + k - 200, reflected at both ends
  for (int i = threadIdx.x; i < NFFT * LM_FRAMES; i += 256) {
    const int fr = i & (LM_FRAMES - 1), k = i >> 6;
    int idx = (f0 + fr) * HOP + k - 200;
    if (idx < 0) idx = -idx;
    if (idx >= NSAMP) idx = 2 * (NSAMP - 1) - idx;
    fT[k * LM_FRAMES + fr] = (idx < n) ? x[idx] : 0.f;
  }
  __syncthreads();

  const int fg = threadIdx.x >> 4, cg = threadIdx.x & 15;  // 4 frames x 4 columns (= 2 bins: re, im, re, im)
#pragma unroll 1
  for (int chunk = 0; chunk < TW_LD / 64; ++chunk) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* tw = a.twiddle + chunk * 64 + cg * 4;
#pragma unroll 4
    for (int k = 0; k < NFFT; ++k) {
      const float4 f = *reinterpret_cast<const float4*>(&fT[k * LM_FRAMES + fg * 4]);
      const float4 w = __ldg(reinterpret_cast<const float4*>(tw + (long long)k * TW_LD));
      const float fv[4] = {f.x, f.y, f.z, f.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(fv[i], wv[j], acc[i][j]);
    }
    const int bin0 = chunk * 32 + cg * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      P[(fg * 4 + i) * P_LD + bin0] = acc[i][0] * acc[i][0] + acc[i][1] * acc[i][1];
      P[(fg * 4 + i) * P_LD + bin0 + 1] = acc[i][2] * acc[i][2] + acc[i][3] * acc[i][3];
    }
  }
  __syncthreads();

  // mel projection (banded), log10, per-clip running max
  float lmax = -INFINITY;
  const int fr = threadIdx.x & 63;
  for (int m = threadIdx.x >> 6; m < NMEL; m += 4) {
    const int lo = a.mel_lo[m], hi = a.mel_hi[m];
    const float* wrow = a.melw + m * NBIN;
    float s = 0.f;
    for (int f = lo; f <= hi; ++f) s = fmaf(__ldg(wrow + f), P[fr * P_LD + f], s);
    const float v = log10f(fmaxf(s, 1e-10f));
    if (f0 + fr < nlive) {
      a.logspec[((long long)b * NFRAMES + f0 + fr) * NMEL + m] = v;
      lmax = fmaxf(lmax, v);
    }
  }
  for (int o = 16; o >= 1; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = lmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = s_max[0];
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_max[i]);
    if (mx > -INFINITY) atomicMax(a.gmax + b, float_to_ordered(mx));
  }
}

__global__ void logmel_init_kernel(unsigned int* gmax, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) gmax[i] = float_to_ordered(-INFINITY);
}

// Clamp to (max - 8), (x + 4) / 4, and emit both layouts. One block per (32 frames, clip).
__global__ void __launch_bounds__(256)
logmel_finish_kernel(const LogMelArgs a) {
  __shared__ float tile[32][NMEL + 1];
  const int b = blockIdx.y, f0 = blockIdx.x * 32;
  int n = a.n_samples[b];
  const int nlive = live_frames(n);
  const float dead = log10f(1e-10f);  // value of every all-zero frame
  float gm = ordered_to_float(a.gmax[b]);
  if (nlive < NFRAMES) gm = fmaxf(gm, dead);
  const float floor_v = gm - 8.0f;
  for (int i = threadIdx.x; i < 32 * NMEL; i += 256) {
    const int fr = i / NMEL, m = i - fr * NMEL;
    const int t = f0 + fr;
    float v = dead;
    if (t < nlive) v = a.logspec[((long long)b * NFRAMES + t) * NMEL + m];
    v = (fmaxf(v, floor_v) + 4.0f) / 4.0f;
    tile[fr][m] = v;
    if (a.conv_in != nullptr && t < NFRAMES)
      a.conv_in[((long long)b * (NFRAMES + 2) + t + 1) * NMEL + m] = __float2bfloat16_rn(v);
  }
  if (a.mel_out == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * NMEL; i += 256) {
    const int m = i >> 5, fr = i & 31;
    const int t = f0 + fr;
    if (t < NFRAMES) a.mel_out[((long long)b * NMEL + m) * NFRAMES + t] = tile[fr][m];
  }
}

}  // namespace

int launch_logmel(const LogMelArgs& a, cudaStream_t st, std::string& err) {
  if (a.B <= 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(logmel_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM);
    if (ce != cudaSuccess) {
      err = std::string("cudaFuncSetAttribute(logmel): ") + cudaGetErrorString(ce);
      return -1;
    }
    attr_set = true;
  }
  logmel_init_kernel<<<ceil_div(a.B, 256), 256, 0, st>>>(a.gmax, a.B);
  int ms = a.max_samples > NSAMP ? NSAMP : a.max_samples;
  int max_live = ms <= 0 ? 0 : (ms + 200 + HOP - 1) / HOP;
  if (max_live > NFRAMES) max_live = NFRAMES;
  if (max_live > 0) {
    dim3 grid(ceil_div(max_live, LM_FRAMES), a.B);
    logmel_dft_kernel<<<grid, 256, LM_SMEM, st>>>(a);
  }
  dim3 g2(ceil_div(NFRAMES, 32), a.B);
  logmel_finish_kernel<<<g2, 256, 0, st>>>(a);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("logmel launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

}  // namespace ssr
