// Row-wise (HBM-bound) kernels: LayerNorm (+erf-GELU, +WavLM relative-position gate), time mean-pool,
// pooled-partials finalize, positional-conv post-processing.
// Reference arithmetic: torch.nn.LayerNorm(eps=1e-5) call sites HF/models/wavlm/modeling_wavlm.py:703-727,100-105,
// 314-373,513; gate HF/models/wavlm/modeling_wavlm.py:167-180; pool REF/WavLM_embeddings.py:321,
// REF/whisper_embeddings_large.py:278.
#include "common.cuh"
#include "ptx.cuh"
#include "kernels.cuh"

namespace ssr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float gelu_erf_r(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---------------------------------------------------------------------------------------------- LayerNorm
// One warp per row; the row lives in registers (NV float4 per lane, D = NV * 128). Two-pass statistics in fp32.
template <int NV, bool IN_BF16>
__global__ void __launch_bounds__(256)
layernorm_kernel(const LayerNormArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Rows are visited from the END of the matrix: the producer (a GEMM epilogue walking its M tiles upwards) wrote the
  // last rows most recently, so the first CTAs find their fp32 input in L2 (the matrix is larger than L2), and the
  // bf16 rows written last here are the ones the consuming GEMM reads first.
  const long long row = a.reverse ? a.rows - 1 - ((long long)blockIdx.x * 8 + warp) : (long long)blockIdx.x * 8 + warp;
  ptx::griddep_wait();    // programmatic dependent launch: the rows come from the previous kernel
  ptx::griddep_launch();
  if (row >= a.rows || row < 0) return;
  constexpr int D = NV * 128;
  float v[NV][4];
  if (IN_BF16) {
    const uint2* src = reinterpret_cast<const uint2*>(a.in_bf16 + row * a.ld_in);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      uint2 u = src[i * 32 + lane];
      __nv_bfloat162 p0 = *reinterpret_cast<__nv_bfloat162*>(&u.x);
      __nv_bfloat162 p1 = *reinterpret_cast<__nv_bfloat162*>(&u.y);
      v[i][0] = __low2float(p0);
      v[i][1] = __high2float(p0);
      v[i][2] = __low2float(p1);
      v[i][3] = __high2float(p1);
    }
  } else {
    const float4* src = reinterpret_cast<const float4*>(a.in_f32 + row * a.ld_in);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 f = src[i * 32 + lane];
      v[i][0] = f.x;
      v[i][1] = f.y;
      v[i][2] = f.z;
      v[i][3] = f.w;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = v[i][j] - mean;
      q += d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + a.eps);
  const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
  const float4* b4 = reinterpret_cast<const float4*>(a.beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(g4 + i * 32 + lane), b = __ldg(b4 + i * 32 + lane);
    v[i][0] = (v[i][0] - mean) * rstd * g.x + b.x;
    v[i][1] = (v[i][1] - mean) * rstd * g.y + b.y;
    v[i][2] = (v[i][2] - mean) * rstd * g.z + b.z;
    v[i][3] = (v[i][3] - mean) * rstd * g.w + b.w;
    if (a.gelu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] = gelu_fast(v[i][j]);
    }
  }
  if (a.out_f32 != nullptr) {
    float4* dst = reinterpret_cast<float4*>(a.out_f32 + row * a.ld_out32);
#pragma unroll
    for (int i = 0; i < NV; ++i) dst[i * 32 + lane] = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
  }
  if (a.out_bf16 != nullptr) {
    uint2* dst = reinterpret_cast<uint2*>(a.out_bf16 + row * a.ld_out16);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[i][0], v[i][1]);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(v[i][2], v[i][3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0);
      u.y = *reinterpret_cast<uint32_t*>(&p1);
      dst[i * 32 + lane] = u;
    }
  }
  if (a.gate_out != nullptr) {
    // WavLM gated relative-position bias, per head h (64 channels = 16 lanes x 4 values of one float4 slot):
    //   ga = sigmoid(wa . x_h + ba), gb = sigmoid(wb . x_h + bb), gate = ga * (gb * const_h - 1) + 2
    // where wa / wb are the sums of rows 0-3 / 4-7 of gru_rel_pos_linear (the reference sums 4 outputs each).
    const int sub = lane & 15;
    const float4 wa = __ldg(reinterpret_cast<const float4*>(a.gate_wa) + sub);
    const float4 wb = __ldg(reinterpret_cast<const float4*>(a.gate_wb) + sub);
    // The 16-lane sums of slot i are kept by lane `sub == i`, so that the sigmoids of all 2 * NV heads are evaluated
    // once, in parallel, instead of by two live lanes per slot.
    float my_da = 0.f, my_db = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float da = v[i][0] * wa.x + v[i][1] * wa.y + v[i][2] * wa.z + v[i][3] * wa.w;
      float db = v[i][0] * wb.x + v[i][1] * wb.y + v[i][2] * wb.z + v[i][3] * wb.w;
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) {
        da += __shfl_xor_sync(0xffffffffu, da, o);
        db += __shfl_xor_sync(0xffffffffu, db, o);
      }
      if (sub == i) {
        my_da = da;
        my_db = db;
      }
    }
    if (sub < NV) {
      const int h = sub * 2 + (lane >> 4);
      const float ga = 1.0f / (1.0f + expf(-(my_da + a.gate_ba)));
      const float gb = 1.0f / (1.0f + expf(-(my_db + a.gate_bb)));
      a.gate_out[row * a.n_heads + h] = ga * (gb * __ldg(a.gate_const + h) - 1.0f) + 2.0f;
    }
  }
}

int g_ln_reverse = 1;

int launch_layernorm(const LayerNormArgs& a_in, cudaStream_t st, std::string& err) {
  if (a_in.rows <= 0) return 0;
  LayerNormArgs a = a_in;
  a.reverse = g_ln_reverse;
  const int grid = (int)((a.rows + 7) / 8);
  const bool bf = a.in_bf16 != nullptr;
#define SSR_LN_CASE(NV)                                                   \
  case NV * 128:                                                          \
    if (bf)                                                               \
      launch_pdl(layernorm_kernel<NV, true>, dim3(grid), dim3(256), 0, st, a);  \
    else                                                                  \
      launch_pdl(layernorm_kernel<NV, false>, dim3(grid), dim3(256), 0, st, a); \
    break;
  switch (a.D) {
    SSR_LN_CASE(2)
    SSR_LN_CASE(3)
    SSR_LN_CASE(4)
    SSR_LN_CASE(6)
    SSR_LN_CASE(8)
    SSR_LN_CASE(10)
    default:
      err = "layernorm: unsupported width " + std::to_string(a.D) + " (supported: 256, 384, 512, 768, 1024, 1280)";
      return -1;
  }
#undef SSR_LN_CASE
  if (a.gate_out != nullptr && a.n_heads * 64 != a.D) {
    err = "layernorm: gate needs head_dim 64";
    return -1;
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("layernorm launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- time mean-pool
// out[b, c] = mean_{t < len_b} x[b * slot + t, c]; sequential fp32 sum in a fixed order (deterministic).
__global__ void __launch_bounds__(128)
pool_mean_kernel(const float* __restrict__ x, int slot, int D, const int* __restrict__ lens, float* __restrict__ out,
                 long long out_stride) {
  const int b = blockIdx.y;
  const int c = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (c >= D) return;
  const int len = lens[b];
  const float4* p = reinterpret_cast<const float4*>(x + (long long)b * slot * D + c);
  const int stride4 = D / 4;
  float4 s0 = make_float4(0, 0, 0, 0), s1 = s0;
  int t = 0;
  // Eight rows in flight per thread (a 1500-row Whisper column walk with two in flight ran at 1.5 TB/s: ncu,
  // profiles/r02_pool_mean_kernel_ncu.json); the additions keep the order of the two-row loop below, bit for bit.
  for (; t + 7 < len; t += 8) {
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = p[(long long)(t + i) * stride4];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      s0.x += v[i].x; s0.y += v[i].y; s0.z += v[i].z; s0.w += v[i].w;
      s1.x += v[i + 1].x; s1.y += v[i + 1].y; s1.z += v[i + 1].z; s1.w += v[i + 1].w;
    }
  }
  for (; t + 1 < len; t += 2) {
    const float4 a = p[(long long)t * stride4], q = p[(long long)(t + 1) * stride4];
    s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
    s1.x += q.x; s1.y += q.y; s1.z += q.z; s1.w += q.w;
  }
  if (t < len) {
    const float4 a = p[(long long)t * stride4];
    s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
  }
  const float inv = len > 0 ? 1.0f / (float)len : 0.f;
  float4 r = make_float4((s0.x + s1.x) * inv, (s0.y + s1.y) * inv, (s0.z + s1.z) * inv, (s0.w + s1.w) * inv);
  *reinterpret_cast<float4*>(out + (long long)b * out_stride + c) = r;
}

int launch_pool_mean(const float* x, int B, int slot, int D, const int* lens, float* out, long long out_stride,
                     cudaStream_t st, std::string& err) {
  if (D % 4) {
    err = "pool: D must be a multiple of 4";
    return -1;
  }
  dim3 grid(ceil_div(D / 4, 128), B);
  pool_mean_kernel<<<grid, 128, 0, st>>>(x, slot, D, lens, out, out_stride);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("pool launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

// Reduce the per-32-row-group partial column sums written by the GEMM epilogue (fixed order -> deterministic).
__global__ void __launch_bounds__(128)
pool_finalize_kernel(const float* __restrict__ part, int slot, int D, const int* __restrict__ lens,
                     float* __restrict__ out, long long out_stride, long long part_layer_stride,
                     long long out_layer_stride) {
  // blockIdx.z = layer: the partial sums of several layers are reduced by one launch
  part += (long long)blockIdx.z * part_layer_stride;
  out += (long long)blockIdx.z * out_layer_stride;
  const int b = blockIdx.y;
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= D) return;
  const int len = lens[b];
  float s = 0.f;
  if (len > 0) {
    const long long r0 = (long long)b * slot, r1 = r0 + len - 1;
    for (long long g = r0 / 32; g <= r1 / 32; ++g) {
      const int b_first = (int)((g * 32) / slot);
      const int seg = b - b_first;  // 0 or 1 because slot >= 32
      s += part[(g * 2 + seg) * D + c];
    }
    s *= 1.0f / (float)len;
  }
  out[(long long)b * out_stride + c] = s;
}

int launch_pool_finalize(const float* part, int B, int slot, int D, const int* lens, float* out, long long out_stride,
                         cudaStream_t st, std::string& err, int n_layers, long long part_layer_stride,
                         long long out_layer_stride) {
  if (n_layers <= 0) return 0;
  dim3 grid(ceil_div(D, 128), B, n_layers);
  pool_finalize_kernel<<<grid, 128, 0, st>>>(part, slot, D, lens, out, out_stride, part_layer_stride,
                                             out_layer_stride);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("pool finalize launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- posconv helpers
// Copy the projected features [B*slot, D] (fp32) into the zero-padded, group-padded bf16 layout the positional
// conv GEMM reads: xp[b, 64 + t, g * 64 + c] for c < gw (gw = D / 16 channels per group), zero elsewhere
// (rows outside [64, 64 + len_b) and pad channels are zero: the conv's zero padding, HF modeling_wavlm.py:52-58).
__global__ void posconv_pack_kernel(const float* __restrict__ x, int slot, int D, int gw, const int* __restrict__ lens,
                                    bf16* __restrict__ xp, int pslot) {
  const int b = blockIdx.y;
  const int pr = blockIdx.x;  // padded row
  const int t = pr - 64;
  const bool live = t >= 0 && t < lens[b];
  bf16* dst = xp + ((long long)b * pslot + pr) * 1024;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const int g = i >> 6, c = i & 63;
    float v = 0.f;
    if (live && c < gw) v = x[((long long)b * slot + t) * D + g * gw + c];
    dst[i] = __float2bfloat16_rn(v);
  }
}

int launch_posconv_pack(const float* x, int B, int slot, int D, const int* lens, bf16* xp, int pslot, cudaStream_t st,
                        std::string& err) {
  if (D % 16 || D / 16 > 64) {
    err = "posconv: hidden size must be 16 groups of <= 64 channels";
    return -1;
  }
  dim3 grid(pslot, B);
  posconv_pack_kernel<<<grid, 256, 0, st>>>(x, slot, D, D / 16, lens, xp, pslot);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("posconv pack launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

// h[b*slot + t, g*gw + c] = x + gelu(conv[b*pslot + t, g*64 + c] + bias)   (HF modeling_wavlm.py:82-90, 404-405)
__global__ void posconv_finish_kernel(const float* __restrict__ conv, int pslot, const float* __restrict__ bias,
                                      const float* __restrict__ x, int slot, int D, int gw,
                                      const int* __restrict__ lens, float* __restrict__ h) {
  const int b = blockIdx.y, t = blockIdx.x;
  const bool live = t < lens[b];
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const int g = i / gw, c = i - g * gw;
    const long long o = ((long long)b * slot + t) * D + i;
    float v = 0.f;
    if (live) v = x[o] + gelu_erf_r(conv[((long long)b * pslot + t) * 1024 + g * 64 + c] + bias[i]);
    h[o] = v;
  }
}

int launch_posconv_finish(const float* conv, int pslot, const float* bias, const float* x, int B, int slot, int D,
                          const int* lens, float* h, cudaStream_t st, std::string& err) {
  dim3 grid(slot, B);
  posconv_finish_kernel<<<grid, 256, 0, st>>>(conv, pslot, bias, x, slot, D, D / 16, lens, h);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("posconv finish launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

}  // namespace ssr
