// Whisper decoder, single-token probe: the reference runs ONE decoder step with input_ids = [[0]]
// over the encoder output and reads every decoder layer's hidden state (REF/whisper_embeddings_large.py:257-262,
// 286-297; HF/models/whisper/modeling_whisper.py:449-506, 691-796).
//
// With one query token per clip the cross-attention is re-associated so that the encoder states are never projected:
//     scores[h, j] = q_h . (Wk_h enc_j)          = (Wk_h^T q_h) . enc_j        =: q'_h . enc_j
//     out_h        = sum_j p[h, j] (Wv_h enc_j + bv_h) = Wv_h (sum_j p[h, j] enc_j) + bv_h
// i.e. attention of H query vectors of width D against K = V = enc [1500, D] (~0.16 GFLOP per clip and layer)
// instead of the K / V projections of all 1500 positions (9.8 GFLOP per clip and layer in the reference).
//
// dec_xattn_kernel does scores, softmax and the weighted sum in ONE pass over enc (HBM-bound: 3.84 MB per clip and
// layer, read once): 16-row chunks of enc stream through a 3-stage TMA ring (128B-swizzled boxes); per chunk the
// 16 warps split D for S = E Q' (mma.sync m16n8k16, bf16, fp32 accumulate; partials reduced through shared memory in
// fixed order), an online softmax over the chunk updates the running max / sum per head, and ctx^T[d, h] += E^T P is
// accumulated in registers (A operand = the same shared-memory tile through ldmatrix.trans). A clip's rows are split
// over a few CTAs (so that the grid covers the SMs); dec_vproj_kernel merges the splits and applies Wv, bv.
// The token-level Linear layers (M = batch rows) reuse the tcgen05 GEMM.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace ssr {

namespace {

using namespace ptx;

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

// ------------------------------------------------------------------------------------------------ token GEMV
// out[m, n] = act(A[m, :] . W[n, :] + bias[n]) (+ resid[m, n]) for a handful of token rows (M <= 16: the per-clip and
// small-batch decoder probe). A 128 x 256 tensor-core tile would hold one live row and a couple of CTAs would stream
// the whole weight matrix; here every warp owns one output column and the grid covers the SMs, so the weights stream
// at HBM speed (the only traffic that matters: 3.3 - 13 MB of bf16 per Linear).
constexpr int GV_MMAX = 16;
__global__ void __launch_bounds__(256)
dec_gemv_kernel(const bf16* __restrict__ A, int M, int K, const bf16* __restrict__ W, int N,
                const float* __restrict__ bias, int act, const float* resid, int ldr, float* out_f32, int ldo32,
                bf16* __restrict__ out_bf16, int ldo16) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  float acc[GV_MMAX];
#pragma unroll
  for (int m = 0; m < GV_MMAX; ++m) acc[m] = 0.f;
  const bf16* wrow = W + (long long)n * K;
  for (int k0 = lane * 8; k0 < K; k0 += 256) {
    const uint4 wv = *reinterpret_cast<const uint4*>(wrow + k0);
    const __nv_bfloat162* w2 = reinterpret_cast<const __nv_bfloat162*>(&wv);
    float wf[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      wf[2 * i] = __low2float(w2[i]);
      wf[2 * i + 1] = __high2float(w2[i]);
    }
#pragma unroll
    for (int m = 0; m < GV_MMAX; ++m) {
      if (m < M) {
        const uint4 av = *reinterpret_cast<const uint4*>(A + (long long)m * K + k0);
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&av);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[m] = fmaf(wf[2 * i], __low2float(a2[i]), acc[m]);
          acc[m] = fmaf(wf[2 * i + 1], __high2float(a2[i]), acc[m]);
        }
      }
    }
  }
  float mine = 0.f;  // lane m keeps row m's sum
#pragma unroll
  for (int m = 0; m < GV_MMAX; ++m) {
    if (m < M) {  // warp-uniform
      float v = acc[m];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == m) mine = v;
    }
  }
  if (lane < M) {
    float v = mine + (bias ? bias[n] : 0.f);
    if (act == ACT_GELU) v = gelu_fast(v);
    if (resid) v += resid[(long long)lane * ldr + n];
    if (out_f32) out_f32[(long long)lane * ldo32 + n] = v;
    if (out_bf16) out_bf16[(long long)lane * ldo16 + n] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------ fused pass
constexpr int XC_ROWS = 16;    // encoder rows per chunk (= MMA M of the score GEMM, K of the context GEMM)
constexpr int XC_STAGES = 3;
constexpr int XC_WARPS = 16;
constexpr int XC_THREADS = XC_WARPS * 32;
constexpr int XC_HP = 24;      // heads padded to three n8 tiles
constexpr int XC_PP = 36;      // byte pitch of a P^T row ([h][16 t] bf16)
constexpr int XC_BOX = XC_ROWS * 128;  // one TMA box: 16 rows x 64 bf16

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct XattnSmem {
  // byte offsets inside dynamic shared memory for a given D
  int stage, qs, red, ps, misc, total, qpitch;
  __host__ __device__ explicit XattnSmem(int D) {
    stage = 0;
    const int stage_bytes = (D / 64) * XC_BOX;
    qpitch = D * 2 + 16;
    qs = XC_STAGES * stage_bytes;
    red = qs + XC_HP * qpitch;
    ps = red + XC_WARPS * XC_ROWS * XC_HP * 4;
    misc = ps + ((XC_HP * XC_PP + 15) & ~15);
    total = misc + 512;
  }
};

// grid = (S, B): CTA (s, b) owns encoder rows [s*R, min(T, (s+1)*R)) of clip b, R a multiple of 16.
// part[(b*S + s)][h][d] = sum_j exp(score_j - m) enc[j, d]; ml[(b*S + s)][h] = (m, l).
template <int MT>  // D = 256 * MT: each warp owns 16*MT columns of enc
__global__ void __launch_bounds__(XC_THREADS, 1)
dec_xattn_kernel(const __grid_constant__ CUtensorMap tme, const float* __restrict__ qp, float* __restrict__ part,
                 float* __restrict__ ml, int T, int H, int R) {
  constexpr int D = 256 * MT;
  constexpr int W = 16 * MT;
  constexpr int STAGE_BYTES = (D / 64) * XC_BOX;
  extern __shared__ __align__(1024) uint8_t smem[];
  const XattnSmem L(D);
  uint8_t* qs = smem + L.qs;
  float* red = reinterpret_cast<float*>(smem + L.red);
  uint8_t* ps = smem + L.ps;
  float* m_run = reinterpret_cast<float*>(smem + L.misc);
  float* l_run = m_run + XC_HP;
  float* alpha_s = l_run + XC_HP;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.misc + 3 * XC_HP * 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, c = lane & 3;
  const int s = blockIdx.x, S = gridDim.x, b = blockIdx.y;
  const int t_begin = s * R, t_end = min(T, t_begin + R);
  const int n_chunks = t_end > t_begin ? (t_end - t_begin + XC_ROWS - 1) / XC_ROWS : 0;
  const long long row_base = (long long)b * T + t_begin;

  if (tid == 0) {
    if (smem_u32(smem) & 1023) __trap();
    prefetch_tmap(&tme);
    for (int i = 0; i < XC_STAGES; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  // q' of this clip -> bf16 [24][D] (padding heads zero), running statistics
  for (int i = tid; i < XC_HP * D; i += XC_THREADS) {
    const int h = i / D, d = i - h * D;
    const float v = h < H ? qp[((long long)b * H + h) * D + d] : 0.f;
    *reinterpret_cast<bf16*>(qs + h * L.qpitch + d * 2) = __float2bfloat16_rn(v);
  }
  if (tid < XC_HP) {
    m_run[tid] = -INFINITY;
    l_run[tid] = 0.f;
  }
  __syncthreads();

  auto issue = [&](int chunk) {  // one thread: all boxes of a chunk into stage chunk % XC_STAGES
    const int st = chunk % XC_STAGES;
    mbar_arrive_expect_tx(&full[st], STAGE_BYTES);
    for (int kb = 0; kb < D / 64; ++kb)
      tma_load_2d(smem + st * STAGE_BYTES + kb * XC_BOX, &tme, &full[st], kb * 64,
                  (int)(row_base + (long long)chunk * XC_ROWS));
  };
  if (tid == 0)
    for (int i = 0; i < XC_STAGES && i < n_chunks; ++i) issue(i);

  float acc[MT][3][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;

  for (int ch = 0; ch < n_chunks; ++ch) {
    const int st = ch % XC_STAGES;
    const uint32_t stage_addr = smem_u32(smem + st * STAGE_BYTES);
    mbar_wait(&full[st], (ch / XC_STAGES) & 1);

    // ---- phase 1: partial scores over this warp's columns: S[t, h] += E[t, d] q'[h, d]
    float sc[3][4];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < MT; ++ks) {
      const int k0 = warp * W + ks * 16;
      const int row = lane & 15;
      const int kc = ((k0 & 63) >> 3) + (lane >> 4);
      uint32_t a[4];
      ldmatrix_x4(stage_addr + (k0 >> 6) * XC_BOX + row * 128 + ((kc ^ (row & 7)) << 4), a);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const uint8_t* qrow = qs + (nt * 8 + g) * L.qpitch + (k0 + 2 * c) * 2;
        mma_bf16_16816(sc[nt], a, *reinterpret_cast<const uint32_t*>(qrow),
                       *reinterpret_cast<const uint32_t*>(qrow + 16));
      }
    }
    {
      float* r = red + warp * (XC_ROWS * XC_HP);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const int h = nt * 8 + 2 * c;
        r[g * XC_HP + h] = sc[nt][0];
        r[g * XC_HP + h + 1] = sc[nt][1];
        r[(g + 8) * XC_HP + h] = sc[nt][2];
        r[(g + 8) * XC_HP + h + 1] = sc[nt][3];
      }
    }
    __syncthreads();  // (A) partial scores visible; every warp has finished phase 2 of the previous chunk
    if (tid == 0 && ch >= 1 && ch - 1 + XC_STAGES < n_chunks) issue(ch - 1 + XC_STAGES);  // refill the freed stage

    // ---- online softmax over the chunk: thread (h, t) with h = tid / 16, t = tid % 16 (a half-warp per head)
    if (tid < XC_HP * XC_ROWS) {
      const int h = tid >> 4, t = tid & 15;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < XC_WARPS; ++w) v += red[w * (XC_ROWS * XC_HP) + t * XC_HP + h];
      const bool valid = t_begin + ch * XC_ROWS + t < t_end;
      v = valid ? v : -INFINITY;
      float cm = v;
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
      const float m_old = m_run[h];
      const float m_new = fmaxf(m_old, cm);  // finite: every chunk holds at least one valid row
      const float p = valid ? __expf(v - m_new) : 0.f;
      float ls = p;
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
      *reinterpret_cast<bf16*>(ps + h * XC_PP + t * 2) = __float2bfloat16_rn(p);
      __syncwarp();
      if (t == 0) {
        const float al = __expf(m_old - m_new);  // 0 on the first chunk (m_old = -inf)
        alpha_s[h] = al;
        m_run[h] = m_new;
        l_run[h] = l_run[h] * al + ls;
      }
    }
    __syncthreads();  // (B) P^T and alpha visible

    // ---- phase 2: ctx^T[d, h] = alpha_h * ctx^T[d, h] + sum_t E[t, d] P[t, h]
    uint32_t pb[3][2];
    float al[3][2];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      const uint8_t* prow = ps + (nt * 8 + g) * XC_PP + (2 * c) * 2;
      pb[nt][0] = *reinterpret_cast<const uint32_t*>(prow);
      pb[nt][1] = *reinterpret_cast<const uint32_t*>(prow + 16);
      al[nt][0] = alpha_s[nt * 8 + 2 * c];
      al[nt][1] = alpha_s[nt * 8 + 2 * c + 1];
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int d0 = warp * W + mt * 16;
      const int mat = lane >> 3, r = lane & 7;
      const int trow = r + ((mat & 2) ? 8 : 0);
      const int dc = ((d0 & 63) >> 3) + (mat & 1);
      uint32_t a[4];
      ldmatrix_x4_trans(stage_addr + (d0 >> 6) * XC_BOX + trow * 128 + ((dc ^ (trow & 7)) << 4), a);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        acc[mt][nt][0] *= al[nt][0];
        acc[mt][nt][1] *= al[nt][1];
        acc[mt][nt][2] *= al[nt][0];
        acc[mt][nt][3] *= al[nt][1];
        mma_bf16_16816(acc[mt][nt], a, pb[nt][0], pb[nt][1]);
      }
    }
  }

  // ---- partial results of this split
  const long long slot = (long long)b * S + s;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int d = warp * W + mt * 16 + g;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      const int h = nt * 8 + 2 * c;
      float* p0 = part + (slot * H + h) * D + d;
      if (h < H) {
        p0[0] = acc[mt][nt][0];
        p0[8] = acc[mt][nt][2];
      }
      if (h + 1 < H) {
        p0[D] = acc[mt][nt][1];
        p0[D + 8] = acc[mt][nt][3];
      }
    }
  }
  __syncthreads();
  if (tid < H) {
    ml[(slot * H + tid) * 2] = m_run[tid];
    ml[(slot * H + tid) * 2 + 1] = l_run[tid];
  }
}

// cv[b, h*64 + e] = Wv[h*64 + e, :] . ctx[b, h, :] + bv[h*64 + e]  -> bf16 (next GEMM's A operand), where ctx merges
// the S splits: ctx = sum_s exp(m_s - M) part_s / sum_s exp(m_s - M) l_s.
__global__ void __launch_bounds__(256)
dec_vproj_kernel(const float* __restrict__ part, const float* __restrict__ ml, int S, const bf16* __restrict__ wv,
                 const float* __restrict__ bv, bf16* __restrict__ out, int D) {
  extern __shared__ float ctx[];  // [D] + weights [S]
  float* wgt = ctx + D;
  const int h = blockIdx.x, b = blockIdx.y, H = gridDim.x;
  if (threadIdx.x == 0) {
    float M = -INFINITY;
    for (int s = 0; s < S; ++s) M = fmaxf(M, ml[(((long long)b * S + s) * H + h) * 2]);
    float Lsum = 0.f;
    for (int s = 0; s < S; ++s) {
      const float* e = ml + (((long long)b * S + s) * H + h) * 2;
      const float w = e[1] > 0.f ? __expf(e[0] - M) : 0.f;  // an empty split has l = 0, m = -inf
      wgt[s] = w;
      Lsum += w * e[1];
    }
    const float inv = 1.0f / Lsum;
    for (int s = 0; s < S; ++s) wgt[s] *= inv;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += 256) {
    float v = 0.f;
    for (int s = 0; s < S; ++s)
      if (wgt[s] != 0.f) v = fmaf(wgt[s], part[(((long long)b * S + s) * H + h) * D + d], v);
    ctx[d] = v;
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

template <int MT>
int launch_xattn(const CUtensorMap& tme, const DecCrossArgs& a, int S, int R, cudaStream_t st, std::string& err) {
  const XattnSmem L(256 * MT);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(dec_xattn_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (ce != cudaSuccess) {
      err = std::string("cudaFuncSetAttribute(dec_xattn_kernel): ") + cudaGetErrorString(ce);
      return -1;
    }
    attr_set = true;
  }
  dec_xattn_kernel<MT><<<dim3(S, a.B), XC_THREADS, L.total, st>>>(tme, a.qp, a.ctx_part, a.ml, a.T, a.H, R);
  return 0;
}

}  // namespace

bool dec_gemv_applicable(int M, int K) { return M >= 1 && M <= GV_MMAX && K % 8 == 0; }

int launch_dec_gemv(const bf16* A, int M, int K, const bf16* W, int N, const EpiParams& ep, cudaStream_t st,
                    std::string& err) {
  if (!dec_gemv_applicable(M, K) || ep.in_slot != 0 || ep.pool_part != nullptr || ep.resid_by_t) {
    err = "decoder GEMV: unsupported shape or epilogue";
    return -1;
  }
  dec_gemv_kernel<<<ceil_div(N, 8), 256, 0, st>>>(A, M, K, W, N, ep.bias, ep.act, ep.resid, ep.ldr, ep.out_f32,
                                                   ep.ldo32, ep.out_bf16, ep.ldo16);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("decoder GEMV launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

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

int dec_cross_splits(int B, int T, int num_sms) {
  int S = num_sms / (B > 0 ? B : 1);
  const int max_s = ceil_div(T, 4 * XC_ROWS);  // at least 4 chunks per split
  if (S > max_s) S = max_s;
  return S < 1 ? 1 : S;
}

int launch_dec_cross_attention(const DecCrossArgs& a, cudaStream_t st, std::string& err) {
  if (a.H > 20 || a.D != a.H * 64 || a.D % 256 != 0 || a.D > 1280) {
    err = "decoder cross-attention: supports head_dim 64, at most 20 heads, d_model a multiple of 256";
    return -1;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int S = dec_cross_splits(a.B, a.T, num_sms);
  const int R = ceil_div(ceil_div(a.T, S), XC_ROWS) * XC_ROWS;
  CUtensorMap tme;
  if (make_tmap_2d(&tme, a.enc, (unsigned long long)a.D, (unsigned long long)a.B * a.T, (unsigned long long)a.D,
                   XC_ROWS, err))
    return -1;
  dec_qproj_kernel<<<dim3(a.H, a.B), 256, 0, st>>>(a.q, a.wk, a.qp, a.D);
  int rc = 0;
  switch (a.D / 256) {
    case 1: rc = launch_xattn<1>(tme, a, S, R, st, err); break;
    case 2: rc = launch_xattn<2>(tme, a, S, R, st, err); break;
    case 3: rc = launch_xattn<3>(tme, a, S, R, st, err); break;
    case 4: rc = launch_xattn<4>(tme, a, S, R, st, err); break;
    default: rc = launch_xattn<5>(tme, a, S, R, st, err); break;
  }
  if (rc) return rc;
  dec_vproj_kernel<<<dim3(a.H, a.B), 256, (a.D + S) * 4, st>>>(a.ctx_part, a.ml, S, a.wv, a.bv, a.out, a.D);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("decoder cross-attention launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

}  // namespace ssr
