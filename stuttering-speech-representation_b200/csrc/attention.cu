// Multi-head self-attention, head_dim 64, flash-style (online softmax, scores never leave the SM).
// bf16 tensor cores via mma.sync.m16n8k16 (round-1 implementation; the tcgen05/TMEM version is the planned upgrade),
// fp32 softmax, fp32 accumulation.
//
// Replaces: WavLM  F.multi_head_attention_forward -> SDPA with the gated relative-position bias as an additive
//           float mask (HF/models/wavlm/modeling_wavlm.py:147-241; bias table :243-271; gate :167-180);
//           Whisper SDPA with pre-scaled q and no mask (HF/models/whisper/modeling_whisper.py:284-357).
// The [B*H, T, T] gated bias of the reference is never materialised: score[i,j] += gate[i,h] * table[h][j-i].
#include "common.cuh"
#include "kernels.cuh"

namespace ssr {

namespace {

constexpr int QB = 64;   // query rows per CTA (4 warps x 16)
constexpr int KB = 64;   // keys per iteration
constexpr int HD = 64;   // head dim
constexpr int LDS = 72;  // padded smem row (bf16 elements): 144 B, conflict-free for ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// Copy a [64 x 64] bf16 tile (row pitch ld elements in global) into padded smem; rows >= rows_avail are zero-filled.
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, long long ld, int rows_avail) {
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int r = i >> 3, ch = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < rows_avail) v = *reinterpret_cast<const uint4*>(src + (long long)r * ld + ch * 8);
    *reinterpret_cast<uint4*>(dst + r * LDS + ch * 8) = v;
  }
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(128)
attention_kernel(const AttentionArgs a) {
  __shared__ __align__(16) bf16 sQ[QB * LDS];
  __shared__ __align__(16) bf16 sK[KB * LDS];
  __shared__ __align__(16) bf16 sV[KB * LDS];
  __shared__ float sRel[128];

  const int qblk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int len = min(a.lens[b], a.slot);
  const int q0 = qblk * QB;
  const long long ld = 3LL * a.D;
  const bf16* base = a.qkv + (long long)b * a.slot * ld;

  load_tile(sQ, base + (long long)q0 * ld + h * HD, ld, a.slot - q0);
  __syncthreads();

  uint32_t qf[4][4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldsm_x4(qf[kk], sQ + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + kk * 16 + (lane >> 4) * 8);

  const int i0 = q0 + warp * 16 + g;  // this thread's two query rows: i0 and i0 + 8 (clip-local)
  float gate0 = 0.f, gate1 = 0.f;
  if (HAS_BIAS) {
    if (i0 < a.slot) gate0 = a.gate[((long long)b * a.slot + i0) * a.H + h];
    if (i0 + 8 < a.slot) gate1 = a.gate[((long long)b * a.slot + i0 + 8) * a.H + h];
  }

  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[n][j] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  constexpr float LOG2E = 1.4426950408889634f;

  const int nkb = (len + KB - 1) / KB;
  for (int kb = 0; kb < nkb; ++kb) {
    const int k0 = kb * KB;
    __syncthreads();  // previous iteration's reads of sK / sV / sRel are complete
    load_tile(sK, base + (long long)k0 * ld + a.D + h * HD, ld, a.slot - k0);
    load_tile(sV, base + (long long)k0 * ld + 2 * a.D + h * HD, ld, a.slot - k0);
    if (HAS_BIAS) {
      // rel = j - i for i in [q0, q0+63], j in [k0, k0+63]  ->  index (j - k0) - (i - q0) + 63 in [0, 126]
      if (threadIdx.x < 127) {
        const int rel = (k0 - q0) + (int)threadIdx.x - 63;
        const int idx = rel + a.rel_center;
        sRel[threadIdx.x] = (idx >= 0 && idx < a.rel_stride) ? a.relbias[(long long)h * a.rel_stride + idx] : 0.f;
      }
    }
    __syncthreads();

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
      for (int kk2 = 0; kk2 < 2; ++kk2) {
        uint32_t kf[4];
        ldsm_x4(kf, sK + (n * 8 + (lane & 7)) * LDS + kk2 * 32 + (lane >> 3) * 8);
        mma_bf16(s[n], qf[kk2 * 2 + 0], kf[0], kf[1]);
        mma_bf16(s[n], qf[kk2 * 2 + 1], kf[2], kf[3]);
      }
    }
    // ---- bias + key mask ----
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jl = n * 8 + tq * 2 + (j & 1);       // key index within block
        const int il = warp * 16 + g + (j >> 1) * 8;   // query index within block
        float v = s[n][j];
        if (HAS_BIAS) v += ((j >> 1) ? gate1 : gate0) * sRel[jl - il + 63];
        if (k0 + jl >= len) v = -INFINITY;
        s[n][j] = v;
      }
    }
    // ---- online softmax ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
      mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float mu0 = (mn0 == -INFINITY) ? 0.f : mn0, mu1 = (mn1 == -INFINITY) ? 0.f : mn1;
    const float sc0 = exp2f((m0 - mu0) * LOG2E), sc1 = exp2f((m1 - mu1) * LOG2E);
    m0 = mn0;
    m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      s[n][0] = exp2f((s[n][0] - mu0) * LOG2E);
      s[n][1] = exp2f((s[n][1] - mu0) * LOG2E);
      s[n][2] = exp2f((s[n][2] - mu1) * LOG2E);
      s[n][3] = exp2f((s[n][3] - mu1) * LOG2E);
      rs0 += s[n][0] + s[n][1];
      rs1 += s[n][2] + s[n][3];
    }
    l0 = l0 * sc0 + rs0;
    l1 = l1 * sc1 + rs1;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      o[n][0] *= sc0;
      o[n][1] *= sc0;
      o[n][2] *= sc1;
      o[n][3] *= sc1;
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pf[4];
      pf[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        uint32_t vf[4];
        ldsm_x4_t(vf, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + n2 * 16 + (lane >> 4) * 8);
        mma_bf16(o[2 * n2 + 0], pf, vf[0], vf[1]);
        mma_bf16(o[2 * n2 + 1], pf, vf[2], vf[3]);
      }
    }
  }

  // ---- finalize: divide by the row sums (quad-reduced), stage through smem, coalesced 16-byte stores ----
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = l0 > 0.f ? 1.0f / l0 : 0.f, inv1 = l1 > 0.f ? 1.0f / l1 : 0.f;
  bf16* sO = sQ + warp * 16 * LDS;  // warp-private rows of the Q tile (Q fragments already live in registers)
  __syncwarp();
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    *reinterpret_cast<uint32_t*>(sO + g * LDS + n * 8 + tq * 2) = pack_bf16(o[n][0] * inv0, o[n][1] * inv0);
    *reinterpret_cast<uint32_t*>(sO + (g + 8) * LDS + n * 8 + tq * 2) = pack_bf16(o[n][2] * inv1, o[n][3] * inv1);
  }
  __syncwarp();
  for (int i = lane; i < 16 * 8; i += 32) {
    const int r = i >> 3, ch = i & 7;
    const int qi = q0 + warp * 16 + r;
    if (qi < a.slot) {
      const uint4 v = *reinterpret_cast<const uint4*>(sO + r * LDS + ch * 8);
      *reinterpret_cast<uint4*>(a.out + ((long long)b * a.slot + qi) * a.D + h * HD + ch * 8) = v;
    }
  }
}

}  // namespace

int launch_attention(const AttentionArgs& a, cudaStream_t st, std::string& err) {
  if (a.D != a.H * HD) {
    err = "attention: head_dim must be 64";
    return -1;
  }
  if (a.B <= 0 || a.slot <= 0) return 0;
  dim3 grid(ceil_div(a.slot, QB), a.H, a.B);
  if (a.gate != nullptr)
    attention_kernel<true><<<grid, 128, 0, st>>>(a);
  else
    attention_kernel<false><<<grid, 128, 0, st>>>(a);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("attention launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

}  // namespace ssr
