// bf16 GEMM for sm_100a:  C[M,N] = A[M,K] · W[N,K]^T  with a fused epilogue (bias, erf-GELU, fp32 residual,
// slot-remapped stores, fused per-clip time pooling partial sums).
//
// Main path: persistent warp-specialised kernel — TMA (cp.async.bulk.tensor, 128B swizzle) -> smem ring ->
// tcgen05.mma (one issuing thread, fp32 accumulators in TMEM, double buffered) -> tcgen05.ld epilogue warps.
// Every Linear layer of the WavLM / Whisper encoders and every Conv1d (as an implicit GEMM: the im2col matrix of a
// channels-last signal is a 2-D view with row stride = conv_stride * C, which a TMA tensor map expresses directly)
// runs through this kernel.  Reference arithmetic being replaced: torch F.linear / F.conv1d calls inside
// HF/models/wavlm/modeling_wavlm.py:93-105,188-241,288-295,682-789 and HF/models/whisper/modeling_whisper.py:284-414,619-625.
//
// A second, deliberately naive SIMT kernel with an independent scalar epilogue exists only for bring-up
// cross-checks (SSR_DEBUG_SIMT_GEMM=1); it is never selected otherwise.
#include "common.cuh"
#include "ptx.cuh"

namespace ssr {

using namespace ptx;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ------------------------------------------------------------------------------------------------ row routing
struct RowInfo {
  bool live;
  int b, t;
  long long orow;
};

__device__ __forceinline__ RowInfo route_row(const EpiParams& e, int M, int r) {
  RowInfo ri;
  if (e.in_slot > 0) {
    ri.b = r / e.in_slot;
    ri.t = r - ri.b * e.in_slot;
    int lim = e.valid;
    if (e.lens != nullptr && r < M) lim = min(lim, __ldg(e.lens + ri.b));
    ri.live = (r < M) && (ri.t < lim);
    ri.orow = (long long)ri.b * e.out_slot + ri.t + e.out_off;
  } else {
    ri.live = r < M;
    ri.orow = r;
    if (e.pool_slot > 0) {
      ri.b = r / e.pool_slot;
      ri.t = r - ri.b * e.pool_slot;
      if (e.lens != nullptr && r < M) ri.live = ri.live && (ri.t < __ldg(e.lens + ri.b));
    } else {
      ri.b = 0;
      ri.t = r;
    }
  }
  return ri;
}

// Sum 32 per-lane vectors of 32 columns across the warp: afterwards lane l holds the column-(l) sum in v[0].
__device__ __forceinline__ void warp_colsum32(float (&v)[32], uint32_t lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      float send = upper ? v[i] : v[i + off];
      float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// One 32-column chunk of one output row, values in registers.
__device__ __forceinline__ void epi_chunk(const EpiParams& e, int N, const RowInfo& ri, int group, int col0,
                                          uint32_t (&raw)[32], uint32_t lane) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);

  if (e.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = __ldg(b4 + i);
      v[4 * i + 0] += b.x;
      v[4 * i + 1] += b.y;
      v[4 * i + 2] += b.z;
      v[4 * i + 3] += b.w;
    }
  }
  if (e.act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  }
  if (e.resid != nullptr && ri.live) {
    const long long rr = e.resid_by_t ? (long long)ri.t : ri.orow;
    const float4* r4 = reinterpret_cast<const float4*>(e.resid + rr * e.ldr + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = r4[i];
      v[4 * i + 0] += b.x;
      v[4 * i + 1] += b.y;
      v[4 * i + 2] += b.z;
      v[4 * i + 3] += b.w;
    }
  }
  if (ri.live) {
    if (e.out_f32 != nullptr) {
      float4* o4 = reinterpret_cast<float4*>(e.out_f32 + ri.orow * e.ldo32 + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
    if (e.out_bf16 != nullptr) {
      uint4* o4 = reinterpret_cast<uint4*>(e.out_bf16 + ri.orow * e.ldo16 + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * i + 0], v[8 * i + 1]);
        __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
        __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&p0);
        u.y = *reinterpret_cast<uint32_t*>(&p1);
        u.z = *reinterpret_cast<uint32_t*>(&p2);
        u.w = *reinterpret_cast<uint32_t*>(&p3);
        o4[i] = u;
      }
    }
  }
  if (e.pool_part != nullptr) {
    // segment 0: rows of the clip that owns the group's first row; segment 1: rows of the following clip.
    const int b_first = __shfl_sync(0xffffffffu, ri.b, 0);
    const bool in0 = ri.live && (ri.b == b_first);
    const bool in1 = ri.live && (ri.b == b_first + 1);
    const unsigned any0 = __ballot_sync(0xffffffffu, in0);
    const unsigned any1 = __ballot_sync(0xffffffffu, in1);
    if (any0) {
      float s[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i] = in0 ? v[i] : 0.0f;
      warp_colsum32(s, lane);
      e.pool_part[((long long)group * 2 + 0) * N + col0 + lane] = s[0];
    }
    if (any1) {
      float s[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i] = in1 ? v[i] : 0.0f;
      warp_colsum32(s, lane);
      e.pool_part[((long long)group * 2 + 1) * N + col0 + lane] = s[0];
    }
  }
}

// ------------------------------------------------------------------------------------------------ tcgen05 kernel
constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;

template <int BN>
struct TcCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + BAR_BYTES + 1024;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + Cfg::B_STAGE_BYTES));
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const int total_tiles = p.num_m_tiles * p.num_n_tiles;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    // ===================== TMA producer =====================
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.num_n_tiles;
      const int n_tile = tile - m_tile * p.num_n_tiles;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], A_STAGE_BYTES + Cfg::B_STAGE_BYTES);
        if (p.a_mode == 0)
          tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, &full[s], kb * BK, m_tile * BM);
        else
          tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, &full[s], n_tile * BN, m_tile * BM + kb);
        tma_load_2d(sB + s * Cfg::B_STAGE_BYTES, &tmB, &full[s], kb * BK, n_tile * BN);
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (threadIdx.x == 32) {
    // ===================== MMA issuer (single thread) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], accph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem_u32(sA + s * A_STAGE_BYTES));
        const uint64_t db = umma_desc_sw128(smem_u32(sB + s * Cfg::B_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advancing K by 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);  // frees the smem slot once these MMAs have consumed it
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
      umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
      acc ^= 1;
      if (acc == 0) accph ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps: TMEM -> registers -> global =====================
    const uint32_t ew = warp - 4;  // == warp % 4: the TMEM lane quadrant this warp may read
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.num_n_tiles;
      const int n_tile = tile - m_tile * p.num_n_tiles;
      mbar_wait(&tfull[acc], accph);
      tc_fence_after();
      const int r = m_tile * BM + ew * 32 + lane;
      const RowInfo ri = route_row(p.epi, p.M, r);
      const int group = m_tile * 4 + ew;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((ew * 32u) << 16) + acc * BN + c * 32, raw);
        tmem_wait_ld();
        epi_chunk(p.epi, p.N, ri, group, n_tile * BN + c * 32, raw, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) accph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ SIMT debug kernel
__device__ __forceinline__ void epi_scalar(const EpiParams& e, const RowInfo& ri, int c, float acc) {
  if (!ri.live) return;
  float v = acc;
  if (e.bias) v += e.bias[c];
  if (e.act == ACT_GELU) v = gelu_erf(v);
  if (e.resid) v += e.resid[(e.resid_by_t ? (long long)ri.t : ri.orow) * e.ldr + c];
  if (e.out_f32) e.out_f32[ri.orow * e.ldo32 + c] = v;
  if (e.out_bf16) e.out_bf16[ri.orow * e.ldo16 + c] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const bf16* __restrict__ A, long long lda, long long a_rows, const bf16* __restrict__ W,
                 const GemmParams p) {
  __shared__ float As[16][65];
  __shared__ float Ws[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int rr = i >> 4, kk = i & 15;
      const int k = k0 + kk;
      float a = 0.f, w = 0.f;
      if (k < p.K) {
        long long row = m0 + rr;
        long long off;
        if (p.a_mode == 0) {
          off = row * lda + k;
        } else {
          row += k >> 6;
          off = row * lda + n0 + (k & 63);
        }
        if (m0 + rr < p.M && row < a_rows) a = __bfloat162float(A[off]);
        if (n0 + rr < p.N) w = __bfloat162float(W[(long long)(n0 + rr) * p.K + k]);
      }
      As[kk][rr] = a;
      Ws[kk][rr] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = As[kk][ty * 4 + i];
        w[i] = Ws[kk][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * w[j];
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    const RowInfo ri = route_row(p.epi, p.M, r);
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c < p.N) epi_scalar(p.epi, ri, c, acc[i][j]);
    }
  }
}

// Debug-path pooling partials (the SIMT kernel does not fuse them): recompute from the stored fp32 output.
__global__ void pool_part_from_out_kernel(const GemmParams p) {
  const EpiParams& e = p.epi;
  const int group = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= p.N) return;
  const RowInfo r0 = route_row(e, p.M, group * 32);
  float s0 = 0.f, s1 = 0.f;
  bool a0 = false, a1 = false;
  for (int i = 0; i < 32; ++i) {
    const RowInfo ri = route_row(e, p.M, group * 32 + i);
    if (!ri.live) continue;
    const float v = e.out_f32[ri.orow * e.ldo32 + c];
    if (ri.b == r0.b) {
      s0 += v;
      a0 = true;
    } else if (ri.b == r0.b + 1) {
      s1 += v;
      a1 = true;
    }
  }
  if (a0) e.pool_part[((long long)group * 2 + 0) * p.N + c] = s0;
  if (a1) e.pool_part[((long long)group * 2 + 1) * p.N + c] = s1;
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn(std::string& err) {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (ce != cudaSuccess || p == nullptr || qres != cudaDriverEntryPointSuccess) {
    err = std::string("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: ") + cudaGetErrorString(ce);
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// 2-D bf16 tensor map: dim0 (contiguous) x dim1 with row pitch `pitch_elems`; box = 64 x box_rows; 128B swizzle.
int make_tmap_2d(CUtensorMap* m, const void* base, unsigned long long dim0, unsigned long long dim1,
                        unsigned long long pitch_elems, unsigned box_rows, std::string& err) {
  PFN_encodeTiled enc = get_encode_fn(err);
  if (!enc) return -1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((pitch_elems * 2) & 15) != 0) {
    err = "tensor map: base or pitch not 16-byte aligned";
    return -1;
  }
  cuuint64_t gdim[2] = {dim0, dim1};
  cuuint64_t gstr[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled failed (CUresult %d) dim0=%llu dim1=%llu pitch=%llu box_rows=%u", (int)r, dim0,
             dim1, pitch_elems, box_rows);
    err = buf;
    return -1;
  }
  return 0;
}

template <int BN>
static int launch_tc(const GemmOp& op, GemmParams& p, cudaStream_t stream, int num_sms, std::string& err) {
  using Cfg = TcCfg<BN>;
  CUtensorMap tmA, tmB;
  if (op.a_mode == 0) {
    if (make_tmap_2d(&tmA, op.A, (unsigned long long)op.K, (unsigned long long)op.a_rows,
                     (unsigned long long)op.lda, BM, err))
      return -1;
  } else {
    if (make_tmap_2d(&tmA, op.A, (unsigned long long)op.a_cols, (unsigned long long)op.a_rows,
                     (unsigned long long)op.lda, BM, err))
      return -1;
  }
  if (make_tmap_2d(&tmB, op.W, (unsigned long long)op.K, (unsigned long long)op.N, (unsigned long long)op.K, BN, err))
    return -1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce =
        cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (ce != cudaSuccess) {
      err = std::string("cudaFuncSetAttribute(gemm_tc_kernel): ") + cudaGetErrorString(ce);
      return -1;
    }
    attr_set = true;
  }
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int grid = total < num_sms ? total : num_sms;
  gemm_tc_kernel<BN><<<grid, 256, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("gemm_tc_kernel launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

int launch_gemm(const GemmOp& op, cudaStream_t stream, bool simt, int num_sms, std::string& err) {
  GemmParams p;
  p.M = op.M;
  p.N = op.N;
  p.K = op.K;
  p.a_mode = op.a_mode;
  p.epi = op.epi;
  if (op.M <= 0 || op.N <= 0 || op.K <= 0) {
    err = "gemm: empty problem";
    return -1;
  }
  if (op.a_mode == 1 && (op.K % 64 != 0 || op.N % 64 != 0)) {
    err = "gemm: positional-conv mode needs K and N multiples of 64";
    return -1;
  }
  if (op.epi.pool_part != nullptr) {
    const int slot = op.epi.in_slot > 0 ? op.epi.in_slot : op.epi.pool_slot;
    if (slot < 32) {
      err = "gemm: fused pooling needs at least 32 rows per clip";
      return -1;
    }
  }
  if (simt) {
    p.num_m_tiles = ceil_div(op.M, 64);
    p.num_n_tiles = ceil_div(op.N, 64);
    p.num_kb = 0;
    dim3 grid(p.num_n_tiles, p.num_m_tiles);
    gemm_simt_kernel<<<grid, 256, 0, stream>>>(op.A, op.lda, op.a_rows, op.W, p);
    if (op.epi.pool_part != nullptr) {
      if (op.epi.out_f32 == nullptr) {
        err = "gemm(simt): pooling needs an fp32 output";
        return -1;
      }
      dim3 g2(ceil_div(op.M, 32), ceil_div(op.N, 128));
      pool_part_from_out_kernel<<<g2, 128, 0, stream>>>(p);
    }
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) {
      err = std::string("gemm_simt_kernel launch: ") + cudaGetErrorString(ce);
      return -1;
    }
    return 0;
  }
  if (op.N % 64 != 0) {
    err = "gemm: N must be a multiple of 64";
    return -1;
  }
  if ((op.epi.ldo32 & 3) || (op.epi.ldo16 & 7) || (op.epi.ldr & 3)) {
    err = "gemm: output / residual leading dimensions must keep 16-byte row alignment";
    return -1;
  }
  p.num_m_tiles = ceil_div(op.M, BM);
  p.num_kb = (op.a_mode == 0) ? ceil_div(op.K, BK) : op.K / BK;
  if (op.a_mode == 1) {
    p.num_n_tiles = op.N / 64;
    return launch_tc<64>(op, p, stream, num_sms, err);
  }
  if (op.N % 256 == 0) {
    p.num_n_tiles = op.N / 256;
    return launch_tc<256>(op, p, stream, num_sms, err);
  }
  if (op.N % 128 == 0) {
    p.num_n_tiles = op.N / 128;
    return launch_tc<128>(op, p, stream, num_sms, err);
  }
  p.num_n_tiles = op.N / 64;
  return launch_tc<64>(op, p, stream, num_sms, err);
}

}  // namespace ssr
