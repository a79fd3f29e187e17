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
// mean and 1 / sqrt(var + 1e-7) of one clip per CTA (population variance, as numpy .var()). One pass: sum and sum of
// squares in fp64 (a 48 000-sample clip of O(1) values leaves ~1e-12 of relative error in the variance even when the
// mean dominates), 128-bit loads, four independent accumulator pairs per thread. The kernel is latency-sized (192 KB
// per CTA): a two-pass scalar version took 67 us however few clips there were, which the host-entry pipeline paid
// once per chunk.
__global__ void __launch_bounds__(512)
wave_stats_kernel(const float* __restrict__ audio, long long ld, const int* __restrict__ n_samples,
                  float* __restrict__ stats, int do_normalize) {
  const int b = blockIdx.x;
  const int n = n_samples[b];
  __shared__ double red[2][16];
  if (!do_normalize || n <= 0) {
    if (threadIdx.x == 0) {
      stats[2 * b] = 0.f;
      stats[2 * b + 1] = 1.f;
    }
    return;
  }
  const float* x = audio + (long long)b * ld;
  double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
  // head up to the first 16-byte boundary, 128-bit body, scalar tail
  const int mis = (int)((reinterpret_cast<uintptr_t>(x) >> 2) & 3);
  const int head = min(n, (4 - mis) & 3);
  const int nq = (n - head) >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x + head);
  int i = threadIdx.x;
  for (; i + 3 * 512 < nq; i += 4 * 512) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(x4 + i + u * 512);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u] += ((double)v[u].x + (double)v[u].y) + ((double)v[u].z + (double)v[u].w);
      q[u] += ((double)v[u].x * v[u].x + (double)v[u].y * v[u].y) + ((double)v[u].z * v[u].z + (double)v[u].w * v[u].w);
    }
  }
  for (; i < nq; i += 512) {
    const float4 v = __ldg(x4 + i);
    s[0] += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    q[0] += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
  }
  if ((int)threadIdx.x < head) {
    const double v = (double)x[threadIdx.x];
    s[1] += v;
    q[1] += v * v;
  }
  const int tail0 = head + 4 * nq;
  if (tail0 + (int)threadIdx.x < n) {
    const double v = (double)x[tail0 + threadIdx.x];
    s[2] += v;
    q[2] += v * v;
  }
  double ss = (s[0] + s[1]) + (s[2] + s[3]), qq = (q[0] + q[1]) + (q[2] + q[3]);
  for (int o = 16; o >= 1; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    qq += __shfl_xor_sync(0xffffffffu, qq, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = ss;
    red[1][threadIdx.x >> 5] = qq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int k = 0; k < 16; ++k) {  // fixed order: deterministic
      ts += red[0][k];
      tq += red[1][k];
    }
    const double mean = ts / (double)n;
    const double var = fmax(tq / (double)n - mean * mean, 0.0);
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

// ------------------------------------------------------------------------------- conv layer 0 on the tensor cores
// LayerNorm variant (feat_extract_norm = "layer", WavLM-Large). The step is power-limited and this layer was 1.5e9
// warp instructions of CUDA-core work per 256 clips (10 FMA per output element before the norm even starts), so the
// 10-tap products move to mma.sync without giving up fp32 accuracy: x and w are each split into bf16 hi + lo parts
// and   x.w ~= xh.wh + xl.wh + xh.wl   is one K = 32 contraction   [xh(10) | xl(10) | xh(10) | 0 0] . [wh; wh; wl; 0]
// (relative error ~2^-16, the dropped xl.wl term). LayerNorm needs the whole 512-channel row, which would be 256
// accumulator registers per thread; instead the products are formed twice: pass 1 keeps only sum / sum of squares,
// pass 2 normalises, applies GELU and stores through a small per-warp staging tile (128-byte row segments).
constexpr int C0M_WARPS = 8;
constexpr int C0M_FRAMES = 16 * C0M_WARPS;  // frames per block iteration
constexpr int C0M_WP = 40;                  // bf16 pitch of a weight row (80 bytes: conflict-free fragment loads)
constexpr int C0M_SP = 72;                  // bf16 pitch of a staging row (64 channels + pad)
constexpr int C0M_XLEN = C0M_FRAMES * 5 + 8;
constexpr int C0M_OFF_GB = 512 * C0M_WP * 2;
constexpr int C0M_OFF_STAGE = C0M_OFF_GB + 2 * 512 * 4;
constexpr int C0M_OFF_X = C0M_OFF_STAGE + C0M_WARPS * 16 * C0M_SP * 2;
constexpr int C0M_OFF_WS = C0M_OFF_X + 2 * C0M_XLEN * 2;  // wbar[10] | G[10][10] (upper triangle, doubled off-diagonals)
constexpr int C0M_SMEM = C0M_OFF_WS + 112 * 4;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256)
wavlm_conv0_mma_kernel(const Conv0Args a, const int iters) {
  extern __shared__ __align__(16) unsigned char c0m_smem[];
  bf16* w_s = reinterpret_cast<bf16*>(c0m_smem);                                   // [channel][k = 0..31] (+ pad)
  float* gb_s = reinterpret_cast<float*>(c0m_smem + C0M_OFF_GB);                   // gamma | beta
  bf16* stage_all = reinterpret_cast<bf16*>(c0m_smem + C0M_OFF_STAGE);             // [warp][16][C0M_SP]
  bf16* xh_s = reinterpret_cast<bf16*>(c0m_smem + C0M_OFF_X);
  bf16* xl_s = xh_s + C0M_XLEN;
  float* ws_s = reinterpret_cast<float*>(c0m_smem + C0M_OFF_WS);
  const int b = blockIdx.y;
  const int n = a.n_samples[b];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  if (threadIdx.x < 110) ws_s[threadIdx.x] = a.wstat[threadIdx.x];

  // weights: k 0-9 -> hi, 10-19 -> hi, 20-29 -> lo, 30-31 -> 0
  for (int i = threadIdx.x; i < 512 * 32; i += 256) {
    const int ch = i >> 5, k = i & 31;
    float v = 0.f;
    if (k < 30) {
      const float w = a.w[ch * 10 + (k % 10)];
      const float hi = __bfloat162float(__float2bfloat16_rn(w));
      v = k < 20 ? hi : w - hi;
    }
    w_s[ch * C0M_WP + k] = __float2bfloat16_rn(v);
  }
  for (int i = threadIdx.x; i < 512; i += 256) {
    gb_s[i] = a.gamma[i];
    gb_s[512 + i] = a.beta[i];
  }
  const float mean = a.stats[2 * b], rstd = a.stats[2 * b + 1];
  const float* x = a.audio + (long long)b * a.audio_ld;

  for (int it = 0; it < iters; ++it) {
    const int t_blk = (blockIdx.x * iters + it) * C0M_FRAMES;
    if (t_blk >= a.slot0) break;  // block-uniform
    __syncthreads();              // previous iteration's readers of xh_s / xl_s are done
    for (int i = threadIdx.x; i < C0M_FRAMES * 5 + 5; i += 256) {
      const int idx = t_blk * 5 + i;
      const float v = (idx < n) ? (x[idx] - mean) * rstd : 0.f;
      const bf16 hi = __float2bfloat16_rn(v);
      xh_s[i] = hi;
      xl_s[i] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
    __syncthreads();

    // A fragments of this warp's 16 frames: rows g and g + 8, two k-steps
    const int f0 = warp * 16;
    uint32_t af[2][4];
    {
      auto pair = [&](int frame, int k) -> uint32_t {  // k even; the pair never straddles a segment (10, 20, 30 are even)
        if (k >= 30) return 0u;
        const bf16* src = k < 10 ? xh_s + frame * 5 + k : k < 20 ? xl_s + frame * 5 + (k - 10)
                                                                 : xh_s + frame * 5 + (k - 20);
        const uint32_t lo = __bfloat16_as_ushort(src[0]), hi = __bfloat16_as_ushort(src[1]);
        return lo | (hi << 16);
      };
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        af[ks][0] = pair(f0 + g, ks * 16 + 2 * c);
        af[ks][1] = pair(f0 + g + 8, ks * 16 + 2 * c);
        af[ks][2] = pair(f0 + g, ks * 16 + 2 * c + 8);
        af[ks][3] = pair(f0 + g + 8, ks * 16 + 2 * c + 8);
      }
    }
    auto tile = [&](int nt, float (&acc)[4]) {
      acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
      const bf16* wrow = w_s + (nt * 8 + g) * C0M_WP + 2 * c;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        mma16816(acc, af[ks], *reinterpret_cast<const uint32_t*>(wrow + ks * 16),
                 *reinterpret_cast<const uint32_t*>(wrow + ks * 16 + 8));
    };

    // Row statistics WITHOUT forming the 512 outputs: y_c = sum_k w[c][k] x[k] is linear in the ten window samples, so
    //   mean_c y   = wbar . x                      wbar[k]   = mean_c w[c][k]
    //   mean_c y^2 = x^T G x                       G[k][k']  = mean_c w[c][k] w[c][k']   (symmetric, 55 distinct)
    // Lane l < 16 does frame f0 + l in fp32 (65 FMAs); the rows of this lane's fragment (g, g + 8) fetch theirs by
    // shuffle. (This replaced a first pass over all 64 channel tiles: 128 MMAs + 512 CUDA-core instructions per warp.)
    float mu_l = 0.f, rs_l = 0.f;
    {
      const int fr = f0 + (lane & 15);
      float xv[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) xv[k] = __bfloat162float(xh_s[fr * 5 + k]) + __bfloat162float(xl_s[fr * 5 + k]);
      float e2 = 0.f;
#pragma unroll
      for (int k = 0; k < 10; ++k) {
        mu_l = fmaf(ws_s[k], xv[k], mu_l);
        float t = 0.f;
#pragma unroll
        for (int k2 = k; k2 < 10; ++k2) t = fmaf(ws_s[10 + k * 10 + k2], xv[k2], t);  // off-diagonals stored doubled
        e2 = fmaf(t, xv[k], e2);
      }
      rs_l = rsqrtf(fmaxf(e2 - mu_l * mu_l, 0.f) + 1e-5f);
    }
    const float mu0 = __shfl_sync(0xffffffffu, mu_l, g), mu1 = __shfl_sync(0xffffffffu, mu_l, g + 8);
    const float rs0 = __shfl_sync(0xffffffffu, rs_l, g), rs1 = __shfl_sync(0xffffffffu, rs_l, g + 8);

    // pass 2: normalise, GELU, store 64 channels (one 128-byte segment per frame) at a time
    bf16* stg = stage_all + warp * (16 * C0M_SP);
    for (int grp = 0; grp < 8; ++grp) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int nt = grp * 8 + j;
        float acc[4];
        tile(nt, acc);
        const int ch = nt * 8 + 2 * c;
        const float2 ga = *reinterpret_cast<const float2*>(&gb_s[ch]);
        const float2 be = *reinterpret_cast<const float2*>(&gb_s[512 + ch]);
        const float y00 = gelu_fast(fmaf((acc[0] - mu0) * rs0, ga.x, be.x));
        const float y01 = gelu_fast(fmaf((acc[1] - mu0) * rs0, ga.y, be.y));
        const float y10 = gelu_fast(fmaf((acc[2] - mu1) * rs1, ga.x, be.x));
        const float y11 = gelu_fast(fmaf((acc[3] - mu1) * rs1, ga.y, be.y));
        *reinterpret_cast<__nv_bfloat162*>(stg + g * C0M_SP + j * 8 + 2 * c) = __floats2bfloat162_rn(y00, y01);
        *reinterpret_cast<__nv_bfloat162*>(stg + (g + 8) * C0M_SP + j * 8 + 2 * c) = __floats2bfloat162_rn(y10, y11);
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 4; ++r) {  // 16 rows x 128 bytes = 128 uint4, four per lane
        const int idx = r * 32 + lane, row = idx >> 3, seg = idx & 7;
        const int t = t_blk + f0 + row;
        if (t < a.slot0)
          *reinterpret_cast<uint4*>(a.out + ((long long)b * a.slot0 + t) * 512 + grp * 64 + seg * 8) =
              *reinterpret_cast<const uint4*>(stg + row * C0M_SP + seg * 8);
      }
      __syncwarp();
    }
  }
}

static bool conv0_fma_forced() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSR_CONV0_FMA");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

}  // namespace

int launch_wavlm_conv0(const Conv0Args& a, cudaStream_t st, std::string& err) {
  if (a.B <= 0) return 0;
  wave_stats_kernel<<<a.B, 512, 0, st>>>(a.audio, a.audio_ld, a.n_samples, a.stats, a.do_normalize);
  dim3 grid(ceil_div(a.slot0, C0_FRAMES), a.B);
  if (a.mode == 0 && !conv0_fma_forced()) {
    // each block walks `iters` groups of 128 frames so that the 40 KB weight image is built once per ~512 frames
    const int groups = ceil_div(a.slot0, C0M_FRAMES);
    const int iters = groups >= 4 ? 4 : groups;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(wavlm_conv0_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C0M_SMEM);
      attr_set = true;
    }
    wavlm_conv0_mma_kernel<<<dim3(ceil_div(groups, iters), a.B), 256, C0M_SMEM, st>>>(a, iters);
  } else if (a.mode == 0) {
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

// ---------------------------------------------------------------------------------------- folded-DFT log-mel kernel
// The 400-point real DFT of a frame, decimated twice by hand so that what is left are four dense ~100 x 101 real
// contractions per frame (40.4 k multiply-adds instead of 161 k for the dense 400 x 402 form):
//   fold in k (cos even / sin odd about k = 200; the periodic hann window is symmetric, so it folds into the tables)
//       e[k] = s[k] + s[400-k],  o[k] = s[k] - s[400-k]            (k = 1..199; e[0] = s[0], e[200] = s[200])
//       Re X[f] = sum_k e[k] w[k] cos(2 pi f k / 400),   Im X[f] = -sum_k o[k] w[k] sin(2 pi f k / 400)
//   fold in f (f -> 200 - f flips the sign of the odd-k terms of Re and of the even-k terms of Im)
//       Ae / Ao = even-k / odd-k parts of Re,  Be / Bo = the same of Im, for f = 0..100 only
//       |X[f]|^2 = (Ae + Ao)^2 + (Be + Bo)^2,    |X[200 - f]|^2 = (Ae - Ao)^2 + (Bo - Be)^2
// These run as fp32 FMAs (a thread owns 4 frames x 8 f x {Ae, Be, Ao, Bo} = 128 accumulators; folded samples and
// table rows come from shared memory as 128-bit loads): fp32 throughout keeps the low-level bins of tonal input as
// accurate as the reference's fp32 FFT, which split-bf16 tensor-core products do not (2^-17 of the LARGEST product).
// The mel filterbank then is the small tensor-core GEMM: power [64 x 208] x filterbank^T [208 x 80] on mma.sync with
// power and weights each split into bf16 hi + lo (three products, relative error ~2^-17 of non-negative terms: no
// cancellation), only over the 16-bin steps where the 8 mels of a tile are non-zero. log10, the per-clip running
// maximum and a coalesced store of the [64 frames x 80] tile follow in the same kernel; the clamp against the
// clip-wide maximum needs every frame of the clip and stays in logmel_finish_kernel (its input is L2-resident).
constexpr int LF = 64;                        // frames per CTA
constexpr int LX = (LF - 1) * HOP + NFFT;     // staged samples of those frames (overlapping windows): 10480
constexpr int LJ = 108;                       // pitch of a folded-sample row: >= 104, 4 * odd (mod 32) -> the float4
                                              // loads of 8 consecutive frames are bank-conflict free
constexpr int LNF = 104;                      // f columns: 0..100 live, 13 thread groups of 8
constexpr int LKC = 8;                        // table rows per staged chunk
constexpr int LNCHUNK = LNF / LKC;            // 13 chunks per parity
constexpr int LTHREADS = 224;                 // 16 frame groups x 13 f groups = 208 working threads, 7 warps
constexpr int LP_LD = 216;                    // pitch of the power tile: >= 208, 8 * odd (mod 32)
constexpr int L_OFF_A = 0;                                // region A: staged samples -> table chunks -> log tile
constexpr int L_OFF_B = LX;                               // region B: folded samples (4 x [64][108]) -> power tile
constexpr int L_TAB_CHUNK = 2 * LKC * LNF;                // floats per chunk (cos rows, then sin rows)
constexpr int LM2_SMEM = (LX + 4 * LF * LJ) * 4;
static_assert(2 * L_TAB_CHUNK <= LX && LF * NMEL <= LX, "region A is reused for the table ring and the log tile");
static_assert(LF * LP_LD <= 4 * LF * LJ, "region B is reused for the power tile");

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(LTHREADS, 1)
logmel_folded_kernel(const LogMelArgs a) {
  extern __shared__ __align__(16) float lm2[];
  float* xs = lm2 + L_OFF_A;
  float* fold = lm2 + L_OFF_B;  // [4][LF][LJ]: Ee (e, even k), Oe (o, even k), Eo, Oo
  __shared__ float s_max[LTHREADS / 32];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * LF;
  int n = a.n_samples[b];
  if (n > NSAMP) n = NSAMP;
  const int nlive = live_frames(n);
  if (f0 >= nlive) return;
  const float* x = a.audio + (long long)b * a.audio_ld;
  const int tid = threadIdx.x;

  // ---- stage the samples of the 64 frames: padded index base + i, reflected at both ends of the 30 s window, zero
  //      beyond the clip. Whole aligned quads inside the clip travel global -> shared as 16-byte cp.async (all of a
  //      thread's copies in flight at once); the edges (reflection, clip end, unaligned rows) go element by element.
  {
    const int base = f0 * HOP - 200;  // multiple of 8
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int q = tid; q < LX / 4; q += LTHREADS) {
      const int i0 = base + 4 * q;
      if (vec_ok && i0 >= 0 && i0 + 3 < n) {
        cp_async16(xs + 4 * q, x + i0);
      } else {
        float t[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int idx = i0 + e;
          if (idx < 0) idx = -idx;
          if (idx >= NSAMP) idx = 2 * (NSAMP - 1) - idx;
          t[e] = (idx < n) ? __ldg(x + idx) : 0.f;
        }
        *reinterpret_cast<float4*>(xs + 4 * q) = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();

  // ---- fold: row m of buffer {0: Ee, 1: Oe, 2: Eo, 3: Oo}, column j <-> k = 2 j + parity. Pad columns are zero
  //      (the matching table rows are zero as well; 0 x garbage would not be). One warp per (parity, frame) row.
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int row = warp; row < 2 * LF; row += LTHREADS / 32) {
      const int par = row >= LF, m = row - par * LF;
      const float* xm = xs + m * HOP;
      float* de = fold + ((2 * par) * LF + m) * LJ;
      float* dO = de + LF * LJ;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = lane + 32 * jj;
        if (j < LJ) {
          const int k = 2 * j + par;
          float e = 0.f, o = 0.f;
          if (k <= 200) {
            const float s1 = xm[k];
            const float s2 = xm[NFFT - k];  // k = 0 reads xm[400], one past the frame (for the last frame: the first
                                            // word of region B); the value is discarded by `edge`
            const bool edge = (k == 0) || (k == 200);
            e = edge ? s1 : s1 + s2;
            o = edge ? 0.f : s1 - s2;
          }
          de[j] = e;
          dO[j] = o;
        }
      }
    }
  }
  __syncthreads();  // xs is dead from here on: region A becomes the table ring

  // ---- the four contractions. Thread (fg, ng): frames fg + 16 i (i < 4), f = 8 ng + c (c < 8).
  const int fg = tid & 15, ng = tid >> 4;
  const bool worker = ng < LNCHUNK;  // 13 f groups; the remaining 16 threads only help with the copies
  float* ring = lm2 + L_OFF_A;
  auto issue_chunk = [&](int c) {  // c = 0..25: parity c / 13, rows (c % 13) * 8 .. + 7 of the cos and the sin table
    const int par = c / LNCHUNK, r0 = (c - par * LNCHUNK) * LKC;
    const float* gc = a.dft_tab + ((size_t)(2 * par) * LNF + r0) * LNF;       // cos rows (contiguous 8 x 104)
    const float* gs = a.dft_tab + ((size_t)(2 * par + 1) * LNF + r0) * LNF;   // sin rows
    float* dst = ring + (c & 1) * L_TAB_CHUNK;
    constexpr int PIECES = LKC * LNF / 4;  // 208 16-byte pieces per table
    for (int p = tid; p < 2 * PIECES; p += LTHREADS) {
      const int t = p >= PIECES, q = p - t * PIECES;
      cp_async16(dst + t * (LKC * LNF) + 4 * q, (t ? gs : gc) + 4 * q);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float acc[2][2][4][8];  // [parity][cos|sin][frame][f]
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[p][t][i][c] = 0.f;

  issue_chunk(0);
#pragma unroll
  for (int par = 0; par < 2; ++par) {
    const float* Ebuf = fold + (2 * par) * LF * LJ;
    const float* Obuf = fold + (2 * par + 1) * LF * LJ;
#pragma unroll 1
    for (int cc = 0; cc < LNCHUNK; ++cc) {
      const int c = par * LNCHUNK + cc;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();  // chunk c has landed for everyone; everyone is done with chunk c - 1
      if (c + 1 < 2 * LNCHUNK) issue_chunk(c + 1);
      if (worker) {
        const float* tc = ring + (c & 1) * L_TAB_CHUNK + ng * 8;
        const float* ts = tc + LKC * LNF;
#pragma unroll
        for (int jb = 0; jb < LKC; jb += 4) {
          float4 e4[4], o4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            e4[i] = *reinterpret_cast<const float4*>(Ebuf + (fg + 16 * i) * LJ + cc * LKC + jb);
            o4[i] = *reinterpret_cast<const float4*>(Obuf + (fg + 16 * i) * LJ + cc * LKC + jb);
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float4 c0 = *reinterpret_cast<const float4*>(tc + (jb + jj) * LNF);
            const float4 c1 = *reinterpret_cast<const float4*>(tc + (jb + jj) * LNF + 4);
            const float4 s0 = *reinterpret_cast<const float4*>(ts + (jb + jj) * LNF);
            const float4 s1 = *reinterpret_cast<const float4*>(ts + (jb + jj) * LNF + 4);
            const float cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            const float sw[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float ev = jj == 0 ? e4[i].x : jj == 1 ? e4[i].y : jj == 2 ? e4[i].z : e4[i].w;
              const float ov = jj == 0 ? o4[i].x : jj == 1 ? o4[i].y : jj == 2 ? o4[i].z : o4[i].w;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                acc[par][0][i][q] = fmaf(ev, cw[q], acc[par][0][i][q]);
                acc[par][1][i][q] = fmaf(ov, sw[q], acc[par][1][i][q]);
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();  // every thread is done with the folded samples and the table ring

  // ---- power tile P[frame][bin] (bins 201..215 zero: the mel contraction runs over K = 208)
  float* P = lm2 + L_OFF_B;
  for (int i = tid; i < LF * (LP_LD - NBIN); i += LTHREADS) {
    const int m = i / (LP_LD - NBIN), cpad = i - m * (LP_LD - NBIN);
    P[m * LP_LD + NBIN + cpad] = 0.f;
  }
  if (worker) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = fg + 16 * i;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int f = ng * 8 + q;
        if (f <= 100) {
          const float ae = acc[0][0][i][q], be = acc[0][1][i][q], ao = acc[1][0][i][q], bo = acc[1][1][i][q];
          const float re1 = ae + ao, im1 = be + bo, re2 = ae - ao, im2 = bo - be;
          P[m * LP_LD + f] = fmaf(re1, re1, im1 * im1);
          if (f < 100) P[m * LP_LD + 200 - f] = fmaf(re2, re2, im2 * im2);
        }
      }
    }
  }
  __syncthreads();

  // ---- mel filterbank on the tensor cores + log10 into the staged [64][80] tile
  float* ltile = lm2 + L_OFF_A;
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, c = lane & 3;
    for (int t = warp; t < (LF / 16) * (NMEL / 8); t += LTHREADS / 32) {
      const int mt = t / (NMEL / 8), nt = t - mt * (NMEL / 8);
      const int ks0 = __ldg(a.mel_band + 2 * nt), ks1 = __ldg(a.mel_band + 2 * nt + 1);
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      const float* p0 = P + (mt * 16 + g) * LP_LD + 2 * c;
      const float* p1 = p0 + 8 * LP_LD;
      const bf16* wh = a.melw_hi + (nt * 8 + g) * 208 + 2 * c;
      const bf16* wl = a.melw_lo + (nt * 8 + g) * 208 + 2 * c;
      for (int ks = ks0; ks < ks1; ++ks) {
        const float2 v[4] = {*reinterpret_cast<const float2*>(p0 + ks * 16),
                             *reinterpret_cast<const float2*>(p1 + ks * 16),
                             *reinterpret_cast<const float2*>(p0 + ks * 16 + 8),
                             *reinterpret_cast<const float2*>(p1 + ks * 16 + 8)};
        uint32_t ah[4], al[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(v[r].x, v[r].y);
          const __nv_bfloat162 l = __floats2bfloat162_rn(v[r].x - __low2float(h), v[r].y - __high2float(h));
          ah[r] = *reinterpret_cast<const uint32_t*>(&h);
          al[r] = *reinterpret_cast<const uint32_t*>(&l);
        }
        const uint32_t bh0 = __ldg(reinterpret_cast<const unsigned int*>(wh + ks * 16));
        const uint32_t bh1 = __ldg(reinterpret_cast<const unsigned int*>(wh + ks * 16 + 8));
        const uint32_t bl0 = __ldg(reinterpret_cast<const unsigned int*>(wl + ks * 16));
        const uint32_t bl1 = __ldg(reinterpret_cast<const unsigned int*>(wl + ks * 16 + 8));
        mma16816(d, al, bh0, bh1);  // small terms first
        mma16816(d, ah, bl0, bl1);
        mma16816(d, ah, bh0, bh1);
      }
      const int r0 = mt * 16 + g, col = nt * 8 + 2 * c;
      *reinterpret_cast<float2*>(ltile + r0 * NMEL + col) =
          make_float2(log10f(fmaxf(d[0], 1e-10f)), log10f(fmaxf(d[1], 1e-10f)));
      *reinterpret_cast<float2*>(ltile + (r0 + 8) * NMEL + col) =
          make_float2(log10f(fmaxf(d[2], 1e-10f)), log10f(fmaxf(d[3], 1e-10f)));
    }
  }
  __syncthreads();

  // ---- coalesced store of the live frames (the tile is contiguous in logspec) + per-clip running maximum
  {
    const int live = min(LF, nlive - f0);
    float4* dst = reinterpret_cast<float4*>(a.logspec + ((long long)b * NFRAMES + f0) * NMEL);
    const float4* src = reinterpret_cast<const float4*>(ltile);
    float lmax = -INFINITY;
    for (int i = tid; i < live * (NMEL / 4); i += LTHREADS) {
      const float4 v = src[i];
      dst[i] = v;
      lmax = fmaxf(lmax, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
    for (int o = 16; o >= 1; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if ((tid & 31) == 0) s_max[tid >> 5] = lmax;
    __syncthreads();
    if (tid == 0) {
      float mx = s_max[0];
      for (int i = 1; i < LTHREADS / 32; ++i) mx = fmaxf(mx, s_max[i]);
      if (mx > -INFINITY) atomicMax(a.gmax + b, float_to_ordered(mx));
    }
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
    if (ce == cudaSuccess)
      ce = cudaFuncSetAttribute(logmel_folded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LM2_SMEM);
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
  if (max_live > 0 && a.dense) {
    dim3 grid(ceil_div(max_live, LM_FRAMES), a.B);
    logmel_dft_kernel<<<grid, 256, LM_SMEM, st>>>(a);
  } else if (max_live > 0) {
    dim3 grid(ceil_div(max_live, LF), a.B);
    logmel_folded_kernel<<<grid, LTHREADS, LM2_SMEM, st>>>(a);
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
