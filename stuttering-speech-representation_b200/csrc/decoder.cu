// Whisper decoder, single-token probe: the reference runs ONE decoder step with input_ids = [[0]]
// over the encoder output and reads every decoder layer's hidden state (REF/whisper_embeddings_large.py:257-262,
// 286-297; HF/models/whisper/modeling_whisper.py:449-506, 691-796).
//
// With one query token per clip the cross-attention is re-associated so that the encoder states are never projected:
//     scores[h, j] = q_h . (Wk_h enc_j)          = (Wk_h^T q_h) . enc_j        =: q'_h . enc_j
//     out_h        = sum_j p[h, j] (Wv_h enc_j + bv_h) = Wv_h (sum_j p[h, j] enc_j) + bv_h
// i.e. two HBM-bound passes over enc [1500, D] per clip and layer (~0.16 GFLOP) instead of the K / V projections of
// all 1500 positions (9.8 GFLOP per clip and layer in the reference). The token-level Linear layers (M = batch rows)
// reuse the tcgen05 GEMM. These kernels are the small fp32 CUDA-core pieces in between.
#include "common.cuh"
#include "kernels.cuh"

namespace ssr {

namespace {

// h[b, :] = v[:]  (decoder hidden_states[0] = embed_tokens[0] + embed_positions[0], identical for every clip)
__global__ void bcast_rows_kernel(const float* __restrict__ v, float* __restrict__ out, int D, long long ld) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < D) out[(long long)b * ld + c] = v[c];
}

// q'[b, h, d] = sum_e q[b, h*64 + e] * Wk[h*64 + e, d]        (Wk: bf16 [D, D], row = k_proj output channel)
__global__ void __launch_bounds__(256)
dec_qproj_kernel(const float* __restrict__ q, const bf16* __restrict__ wk, float* __restrict__ qp, int D) {
  __shared__ float qs[64];
  const int h = blockIdx.x, b = blockIdx.y, H = gridDim.x;
  if (threadIdx.x < 64) qs[threadIdx.x] = q[(long long)b * D + h * 64 + threadIdx.x];
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += 256) {
    float acc = 0.f;
#pragma unroll 8
    for (int e = 0; e < 64; ++e) acc = fmaf(qs[e], __bfloat162float(wk[(long long)(h * 64 + e) * D + d]), acc);
    qp[((long long)b * H + h) * D + d] = acc;
  }
}

// scores[b, h, j] = enc[b, j, :] . q'[b, h, :]   ; a warp owns 4 encoder rows, q' of the clip sits in shared memory.
constexpr int SC_ROWS = 32;  // rows per block (8 warps x 4)
template <int HMAX>
__global__ void __launch_bounds__(256)
dec_scores_kernel(const bf16* __restrict__ enc, const float* __restrict__ qp, float* __restrict__ scores, int T, int D,
                  int H) {
  extern __shared__ float qsm[];  // [H][D]
  const int b = blockIdx.y;
  const int j0 = blockIdx.x * SC_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < H * D; i += 256) qsm[i] = qp[(long long)b * H * D + i];
  __syncthreads();
  float acc[4][HMAX];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int h = 0; h < HMAX; ++h) acc[r][h] = 0.f;
  const int jr = j0 + warp * 4;
  const bf16* e0 = enc + ((long long)b * T + jr) * D;
  for (int d = lane; d < D; d += 32) {
    float ev[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) ev[r] = (jr + r < T) ? __bfloat162float(e0[(long long)r * D + d]) : 0.f;
#pragma unroll
    for (int h = 0; h < HMAX; ++h) {
      if (h < H) {
        const float w = qsm[h * D + d];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][h] = fmaf(ev[r], w, acc[r][h]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int h = 0; h < HMAX; ++h) {
      float v = acc[r][h];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && h < H && jr + r < T) scores[((long long)b * H + h) * T + jr + r] = v;
    }
}

// in-place softmax over j for every (clip, head)
__global__ void __launch_bounds__(256)
dec_softmax_kernel(float* __restrict__ s, int T) {
  __shared__ float red[8];
  float* row = s + (long long)blockIdx.x * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < T; j += 256) m = fmaxf(m, row[j]);
  for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float l = 0.f;
  for (int j = threadIdx.x; j < T; j += 256) {
    const float p = expf(row[j] - m);
    row[j] = p;
    l += p;
  }
  for (int o = 16; o >= 1; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if (lane == 0) red[warp] = l;
  __syncthreads();
  l = 0.f;
  for (int i = 0; i < 8; ++i) l += red[i];
  const float inv = 1.0f / l;
  for (int j = threadIdx.x; j < T; j += 256) row[j] *= inv;
}

// ctx_part[seg, b, h, d] = sum_{j in segment} p[b, h, j] * enc[b, j, d] ; a thread owns one column d and all heads.
constexpr int CTX_SEGS = 4;
template <int HMAX>
__global__ void __launch_bounds__(256)
dec_ctx_kernel(const bf16* __restrict__ enc, const float* __restrict__ p, float* __restrict__ part, int T, int D, int H,
               int B) {
  __shared__ float ps[HMAX][64];
  const int b = blockIdx.z, seg = blockIdx.y;
  const int d = blockIdx.x * 256 + threadIdx.x;
  const int seg_len = (T + CTX_SEGS - 1) / CTX_SEGS;
  const int ja = seg * seg_len, jb = min(T, ja + seg_len);
  float acc[HMAX];
#pragma unroll
  for (int h = 0; h < HMAX; ++h) acc[h] = 0.f;
  for (int j0 = ja; j0 < jb; j0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < H * 64; i += 256) {
      const int h = i >> 6, jj = i & 63;
      ps[h][jj] = (j0 + jj < jb) ? p[((long long)b * H + h) * T + j0 + jj] : 0.f;
    }
    __syncthreads();
    if (d < D) {
      const int n = min(64, jb - j0);
      for (int jj = 0; jj < n; ++jj) {
        const float e = __bfloat162float(enc[((long long)b * T + j0 + jj) * D + d]);
#pragma unroll
        for (int h = 0; h < HMAX; ++h)
          if (h < H) acc[h] = fmaf(ps[h][jj], e, acc[h]);
      }
    }
  }
  if (d < D) {
#pragma unroll
    for (int h = 0; h < HMAX; ++h)
      if (h < H) part[(((long long)seg * B + b) * H + h) * D + d] = acc[h];
  }
}

// cv[b, h*64 + e] = Wv[h*64 + e, :] . (sum_seg ctx_part[seg, b, h, :]) + bv[h*64 + e]   -> bf16 (next GEMM's A operand)
__global__ void __launch_bounds__(256)
dec_vproj_kernel(const float* __restrict__ part, const bf16* __restrict__ wv, const float* __restrict__ bv,
                 bf16* __restrict__ out, int D, int B) {
  extern __shared__ float ctx[];  // [D]
  const int h = blockIdx.x, b = blockIdx.y, H = gridDim.x;
  for (int d = threadIdx.x; d < D; d += 256) {
    float s = 0.f;
    for (int seg = 0; seg < CTX_SEGS; ++seg) s += part[(((long long)seg * B + b) * H + h) * D + d];
    ctx[d] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = warp; e < 64; e += 8) {
    const bf16* wr = wv + (long long)(h * 64 + e) * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(__bfloat162float(wr[d]), ctx[d], acc);
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[(long long)b * D + h * 64 + e] = __float2bfloat16_rn(acc + bv[h * 64 + e]);
  }
}

}  // namespace

int launch_bcast_rows(const float* v, float* out, int B, int D, long long ld, cudaStream_t st, std::string& err) {
  dim3 grid(ceil_div(D, 256), B);
  bcast_rows_kernel<<<grid, 256, 0, st>>>(v, out, D, ld);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("bcast_rows launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

int launch_dec_cross_attention(const DecCrossArgs& a, cudaStream_t st, std::string& err) {
  if (a.H > 20 || a.D != a.H * 64) {
    err = "decoder cross-attention: supports head_dim 64 and at most 20 heads";
    return -1;
  }
  static bool attr_set = false;
  const int sc_smem = a.H * a.D * 4;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(dec_scores_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1280 * 4);
    if (ce != cudaSuccess) {
      err = std::string("cudaFuncSetAttribute(dec_scores_kernel): ") + cudaGetErrorString(ce);
      return -1;
    }
    attr_set = true;
  }
  dec_qproj_kernel<<<dim3(a.H, a.B), 256, 0, st>>>(a.q, a.wk, a.qp, a.D);
  dec_scores_kernel<20><<<dim3(ceil_div(a.T, SC_ROWS), a.B), 256, sc_smem, st>>>(a.enc, a.qp, a.scores, a.T, a.D, a.H);
  dec_softmax_kernel<<<a.B * a.H, 256, 0, st>>>(a.scores, a.T);
  dec_ctx_kernel<20><<<dim3(ceil_div(a.D, 256), CTX_SEGS, a.B), 256, 0, st>>>(a.enc, a.scores, a.ctx_part, a.T, a.D,
                                                                             a.H, a.B);
  dec_vproj_kernel<<<dim3(a.H, a.B), 256, a.D * 4, st>>>(a.ctx_part, a.wv, a.bv, a.out, a.D, a.B);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("decoder cross-attention launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

}  // namespace ssr
