// bf16 GEMM for sm_100a:  C[M,N] = A[M,K] · W[N,K]^T  with a fused epilogue (bias, erf-GELU, fp32 residual,
// slot-remapped stores, fused per-clip time pooling partial sums).
//
// Main path: persistent warp-specialised kernels — TMA (cp.async.bulk.tensor, 128B swizzle) -> smem ring ->
// tcgen05.mma (one issuing thread, fp32 accumulators in TMEM, double buffered) -> tcgen05.ld epilogue warps.
//   * gemm_tc2_kernel : CTA pair (cta_group::2), 256 x 256 output tile per SM pair — every large GEMM (N % 256 == 0)
//   * gemm_tc_kernel  : single CTA, 128 x {64,128,256} tile — the positional conv (block-diagonal, N tile 64) and
//                       odd shapes
// Every Linear layer of the WavLM / Whisper encoders and every Conv1d (as an implicit GEMM: the im2col matrix of a
// channels-last signal is a 2-D view with row stride = conv_stride * C, which a TMA tensor map expresses directly)
// runs through these kernels.  Reference arithmetic being replaced: torch F.linear / F.conv1d calls inside
// HF/models/wavlm/modeling_wavlm.py:93-105,188-241,288-295,682-789 and HF/models/whisper/modeling_whisper.py:284-414,619-625.
//
// A deliberately naive SIMT kernel with an independent scalar epilogue exists only for bring-up cross-checks
// (engine option "simt_gemm"); it is never selected otherwise.
#include <stdlib.h>

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

// ------------------------------------------------------------------------------------------------ epilogue
// Per-warp staging area. The accumulator arrives thread-per-row (TMEM lane = output row); global memory wants
// row-contiguous accesses. A 32 x 32 fp32 chunk is transposed through `tile` (pitch 36 floats: 16-byte aligned,
// bank-spread) so that every global load / store instruction touches whole 128-byte (fp32) or 64-byte (bf16) row
// segments. `orow` / `rrow` hold the routed output / residual row of the warp's 32 rows (-1 = dead row) and `bias`
// the tile's bias slice, fetched once per tile BEFORE the accumulator is waited for (no load latency per chunk).
constexpr int STG_LD = 36;
struct __align__(16) EpiStage {
  float tile[32 * STG_LD];
  float bias[256];
  int orow[32];
  int rrow[32];
};

// Coalesced residual fetch for one 32-column chunk (4 rows x 128 contiguous bytes per instruction), issued one chunk
// ahead of its use so that the HBM latency overlaps the previous chunk's work.
__device__ __forceinline__ void load_resid(const EpiParams& e, const EpiStage& st, int col0, uint32_t lane,
                                           float4 (&x)[8]) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = st.rrow[it * 4 + (lane >> 3)];
    x[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rr >= 0) x[it] = *reinterpret_cast<const float4*>(e.resid + (long long)rr * e.ldr + col0 + (lane & 7) * 4);
  }
}

// One 32-column chunk of the warp's 32 output rows. `raw`: this thread's row of the accumulator; `x`: pre-fetched
// residual chunk (coalesced layout); `bcol`: offset of the chunk inside the staged bias slice.
__device__ __forceinline__ void epi_chunk(const EpiParams& e, int N, const RowInfo& ri, int group, int col0, int bcol,
                                          uint32_t (&raw)[32], uint32_t lane, EpiStage& st, const float4 (&x)[8]) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);

  if (e.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = *reinterpret_cast<const float4*>(&st.bias[bcol + i * 4]);  // warp-wide broadcast
      v[4 * i + 0] += b.x;
      v[4 * i + 1] += b.y;
      v[4 * i + 2] += b.z;
      v[4 * i + 3] += b.w;
    }
  }
  if (e.act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
  }
  if (e.resid != nullptr) {
#pragma unroll
    for (int it = 0; it < 8; ++it)
      *reinterpret_cast<float4*>(&st.tile[(it * 4 + (lane >> 3)) * STG_LD + (lane & 7) * 4]) = x[it];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = *reinterpret_cast<const float4*>(&st.tile[lane * STG_LD + i * 4]);
      v[4 * i + 0] += b.x;
      v[4 * i + 1] += b.y;
      v[4 * i + 2] += b.z;
      v[4 * i + 3] += b.w;
    }
    __syncwarp();
  }
  if (e.out_f32 != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(&st.tile[lane * STG_LD + i * 4]) =
          make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = it * 4 + (lane >> 3);
      const int orow = st.orow[row];
      if (orow >= 0)
        *reinterpret_cast<float4*>(e.out_f32 + (long long)orow * e.ldo32 + col0 + (lane & 7) * 4) =
            *reinterpret_cast<const float4*>(&st.tile[row * STG_LD + (lane & 7) * 4]);
    }
    __syncwarp();
  }
  if (e.out_bf16 != nullptr) {
    // bf16 rows are 64 bytes; staged with an 80-byte pitch (20 floats)
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
      *reinterpret_cast<uint4*>(&st.tile[lane * 20 + i * 4]) = u;
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 8 + (lane >> 2);
      const int orow = st.orow[row];
      if (orow >= 0)
        *reinterpret_cast<uint4*>(e.out_bf16 + (long long)orow * e.ldo16 + col0 + (lane & 3) * 8) =
            *reinterpret_cast<const uint4*>(&st.tile[row * 20 + (lane & 3) * 4]);
    }
    __syncwarp();
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

// Epilogue of one output tile for one warp: 32 rows (TMEM lane quadrant `quad`, first row r0) x the 32-column chunks
// c = split, split + NSPLIT, ... of the tile's BN columns (NSPLIT warps share a quadrant).
template <int BN, int NSPLIT>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, EpiStage& stg, uint32_t tmem_acc, uint32_t quad,
                                              int r0, int col_base, uint32_t lane, int split, uint64_t* tfull,
                                              uint32_t parity) {
  const RowInfo ri = route_row(p.epi, p.M, r0 + (int)lane);
  const int group = r0 >> 5;
  stg.orow[lane] = ri.live ? (int)ri.orow : -1;
  stg.rrow[lane] = ri.live ? (p.epi.resid_by_t ? ri.t : (int)ri.orow) : -1;
  if (p.epi.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < BN / 32; ++i) stg.bias[i * 32 + lane] = __ldg(p.epi.bias + col_base + i * 32 + lane);
  }
  __syncwarp();
  const bool has_resid = p.epi.resid != nullptr;
  float4 x[8], xn[8];
  if (has_resid) load_resid(p.epi, stg, col_base + split * 32, lane, x);  // in flight while the MMAs finish
  mbar_wait(tfull, parity);
  __syncwarp();
  tc_fence_after();
#pragma unroll 1
  for (int c = split; c < BN / 32; c += NSPLIT) {
    if (has_resid && c + NSPLIT < BN / 32) load_resid(p.epi, stg, col_base + (c + NSPLIT) * 32, lane, xn);
    uint32_t raw[32];
    tmem_ld_32x32(tmem_acc + ((quad * 32u) << 16) + c * 32, raw);
    tmem_wait_ld();
    epi_chunk(p.epi, p.N, ri, group, col_base + c * 32, c * 32, raw, lane, stg, x);
    if (has_resid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = xn[i];
    }
  }
  tc_fence_before();
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ single-CTA kernel
constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int EPI_WARPS1 = 4;

template <int BN>
struct TcCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int STAGE_OFF = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + ((BAR_BYTES + 15) & ~15);
  static constexpr int SMEM_BYTES = STAGE_OFF + EPI_WARPS1 * (int)sizeof(EpiStage) + 1024;
};

// Warp roles: 0-3 epilogue (TMEM lane quadrant = warp id), 4 TMA producer, 5 MMA issuer (+ TMEM allocation).
template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  // Declared 1024-byte aligned (128B-swizzle atoms) and used directly, so that the compiler keeps every access in the
  // shared address space (LDS/STS); an integer round-trip to align the base would demote them to generic LD/ST.
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + Cfg::B_STAGE_BYTES));
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  EpiStage* stages = reinterpret_cast<EpiStage*>(smem + Cfg::STAGE_OFF);

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
      mbar_init(&tempty[i], EPI_WARPS1);
    }
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 128) {
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
  } else if (threadIdx.x == 160) {
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
  } else if (warp < EPI_WARPS1) {
    // ===================== epilogue warps =====================
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.num_n_tiles;
      const int n_tile = tile - m_tile * p.num_n_tiles;
      epilogue_tile<BN, 1>(p, stages[warp], tmem_base + acc * BN, warp, m_tile * BM + (int)warp * 32, n_tile * BN,
                           lane, 0, &tfull[acc], accph);
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) accph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ CTA-pair kernel
// cta_group::2 variant for the large GEMMs: a cluster of two CTAs (one SM pair) owns a 256 x 256 output tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 output columns), so the per-SM
// fabric traffic per MMA cycle drops by a third compared with the single-CTA 128 x 256 tile (32 KB instead of 48 KB
// per 64-deep k-block). The leader CTA's single MMA thread issues M=256 instructions that read both CTAs' shared
// memory and write both CTAs' TMEM; commits are multicast to both CTAs' barriers.
// Warp roles: 0-7 epilogue (two warps per TMEM lane quadrant, alternating 32-column chunks), 8 TMA producer,
// 9 MMA issuer (+ TMEM allocation). Eight epilogue warps double the thread-level parallelism that hides the
// latencies of the per-chunk TMEM load / smem transpose / global access chain.
constexpr int B2_STAGE_BYTES = 128 * BK * 2;  // half of a 256-column B tile
constexpr int STAGES2 = 5;
constexpr int EPI_WARPS2 = 8;
constexpr int TC2_THREADS = (EPI_WARPS2 + 2) * 32;
constexpr int TC2_BAR_BYTES = ((2 * STAGES2 + 4) * 8 + 16 + 15) & ~15;
constexpr int TC2_STAGE_OFF = STAGES2 * (A_STAGE_BYTES + B2_STAGE_BYTES) + TC2_BAR_BYTES;
constexpr int TC2_SMEM_BYTES = TC2_STAGE_OFF + EPI_WARPS2 * (int)sizeof(EpiStage) + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int BN = 256;
  // Declared 1024-byte aligned (128B-swizzle atoms) and used directly, so that the compiler keeps every access in the
  // shared address space (LDS/STS); an integer round-trip to align the base would demote them to generic LD/ST.
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES2 * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES2 * (A_STAGE_BYTES + B2_STAGE_BYTES));
  uint64_t* full = bars;                      // leader's copy is the live one (both CTAs' TMA bytes land there)
  uint64_t* empty = bars + STAGES2;           // per CTA (multicast commit)
  uint64_t* tfull = bars + 2 * STAGES2;       // per CTA (multicast commit)
  uint64_t* tempty = bars + 2 * STAGES2 + 2;  // leader's copy: epilogue warps of both CTAs arrive on it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 4);
  EpiStage* stages = reinterpret_cast<EpiStage*>(smem + TC2_STAGE_OFF);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int total_tiles = p.num_m_tiles * p.num_n_tiles;  // num_m_tiles counts 256-row tiles here
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * EPI_WARPS2);
    }
    fence_mbar_init();
  }
  if (warp == EPI_WARPS2 + 1) {
    tmem_alloc_cg2(tmem_slot, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above is independent of earlier kernels; from here on activations / residuals of the predecessor are
  // read and its inputs may be overwritten (programmatic dependent launch, see ptx.cuh).
  griddep_wait();
  griddep_launch();

  if (threadIdx.x == EPI_WARPS2 * 32) {
    // ===================== TMA producer (both CTAs; bytes are credited to the leader's barrier) =================
    int s = 0;
    uint32_t ph = 0;
    for (int tile = cid; tile < total_tiles; tile += ncl) {
      const int m_tile = tile / p.num_n_tiles;
      const int n_tile = tile - m_tile * p.num_n_tiles;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * (A_STAGE_BYTES + B2_STAGE_BYTES));
        tma_load_2d_cg2(sA + s * A_STAGE_BYTES, &tmA, &full[s], kb * BK, m_tile * 256 + (int)rank * 128);
        tma_load_2d_cg2(sB + s * B2_STAGE_BYTES, &tmB, &full[s], kb * BK, n_tile * BN + (int)rank * 128);
        if (++s == STAGES2) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (threadIdx.x == (EPI_WARPS2 + 1) * 32 && rank == 0) {
    // ===================== MMA issuer (leader CTA only) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = cid; tile < total_tiles; tile += ncl) {
      mbar_wait(&tempty[acc], accph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem_u32(sA + s * A_STAGE_BYTES));
        const uint64_t db = umma_desc_sw128(smem_u32(sB + s * B2_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit_cg2(&empty[s], 3);
        if (++s == STAGES2) {
          s = 0;
          ph ^= 1;
        }
      }
      umma_commit_cg2(&tfull[acc], 3);
      acc ^= 1;
      if (acc == 0) accph ^= 1;
    }
  } else if (warp < EPI_WARPS2) {
    // ===================== epilogue warps (both CTAs, 128 rows each) =====================
    const uint32_t quad = warp & 3;
    const int split = (int)(warp >> 2);
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = cid; tile < total_tiles; tile += ncl) {
      const int m_tile = tile / p.num_n_tiles;
      const int n_tile = tile - m_tile * p.num_n_tiles;
      epilogue_tile<BN, 2>(p, stages[warp], tmem_base + acc * BN, quad,
                           m_tile * 256 + (int)rank * 128 + (int)quad * 32, n_tile * BN, lane, split, &tfull[acc],
                           accph);
      if (lane == 0) mbar_arrive_cluster(&tempty[acc], 0);  // the leader's MMA thread owns accumulator reuse
      acc ^= 1;
      if (acc == 0) accph ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still touch its smem / TMEM
  if (warp == EPI_WARPS2 + 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ positional conv
// WavLM's grouped positional convolution (16 groups x 64 channels, 128 taps, zero padding 64;
// HF/models/wavlm/modeling_wavlm.py:48-90) as a Toeplitz GEMM:  out[t, g, co] = sum_tap sum_ci x[t + tap, g, ci] w.
// A work item is (clip, 256-row tile, group). Its input window x[t0 .. t0+383, g*64 .. +64] is staged in shared
// memory ONCE (48 KB, three TMA boxes); the A operand of tap `tap` is that same window shifted down by `tap` rows,
// which costs nothing: the UMMA descriptor's start address simply moves by tap * 128 B. Measured on B200: the
// 128B swizzle is applied to ABSOLUTE shared-memory address bits (both by TMA when writing and by the tensor core
// when reading), so a start address that is only 128B- (not 1024B-) aligned needs no base-offset correction
// (base_offset = tap & 7 or its complement give wrong results; 0 matches the generic-GEMM cross-check bit for bit).
// Only the group's weights stream (8 KB per tap) through a TMA ring. Compared with re-loading a shifted A tile per
// tap this removes ~2/3 of the L2 traffic.
constexpr int PC_XWIN_BYTES = 3 * 128 * 128;  // 384 rows x 64 bf16
constexpr int PC_TAP_BYTES = 64 * 64 * 2;  // one tap's [64 out x 64 in] weight tile
constexpr int PC_TAPS = 4;                 // taps per ring stage: the MMA thread pays one barrier wait per 32 MMAs
constexpr int PC_W_BYTES = PC_TAPS * PC_TAP_BYTES;
constexpr int PC_WSTAGES = 3;  // 96 KB of weight tiles in flight
constexpr int PC_BAR_OFF = 2 * PC_XWIN_BYTES + PC_WSTAGES * PC_W_BYTES;
constexpr int PC_STAGE_OFF = PC_BAR_OFF + 512;  // (8 + 2 * PC_WSTAGES) mbarriers + the TMEM base slot
constexpr int PC_SMEM_BYTES = PC_STAGE_OFF + 4 * (int)sizeof(EpiStage);

struct PosConvParams {
  int B, pslot, n_mt, n_items;  // rows per clip in X, 256-row tiles per clip, total items
  GemmParams g;                 // M = B * pslot (flat X rows), N = 1024, epilogue
};

__global__ void __launch_bounds__(192, 1)
posconv_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const PosConvParams pc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint8_t* sX = smem;
  uint8_t* sW = smem + 2 * PC_XWIN_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PC_BAR_OFF);
  uint64_t* x_full = bars;                      // [2]
  uint64_t* x_empty = bars + 2;                 // [2]
  uint64_t* w_full = bars + 4;                  // [PC_WSTAGES]
  uint64_t* w_empty = bars + 4 + PC_WSTAGES;    // [PC_WSTAGES]
  uint64_t* tfull = bars + 4 + 2 * PC_WSTAGES;  // [2]
  uint64_t* tempty = tfull + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  EpiStage* stages = reinterpret_cast<EpiStage*>(smem + PC_STAGE_OFF);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    for (int i = 0; i < PC_WSTAGES; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (clip b, row tile mt, group g); rowbase = first flat X row of the window = first flat output row
  auto decode = [&](int idx, int& g, int& rowbase) {
    g = idx & 15;
    const int bm = idx >> 4;
    const int b = bm / pc.n_mt, mt = bm - b * pc.n_mt;
    rowbase = b * pc.pslot + mt * 256;
  };

  if (threadIdx.x == 128) {
    // ===================== TMA producer =====================
    uint32_t n = 0, nw = 0;
    for (int idx = blockIdx.x; idx < pc.n_items; idx += gridDim.x, ++n) {
      int g, rowbase;
      decode(idx, g, rowbase);
      const int xb = n & 1;
      mbar_wait(&x_empty[xb], ((n >> 1) & 1) ^ 1);
      mbar_arrive_expect_tx(&x_full[xb], PC_XWIN_BYTES);
#pragma unroll
      for (int i = 0; i < 3; ++i)
        tma_load_2d(sX + xb * PC_XWIN_BYTES + i * 128 * 128, &tmX, &x_full[xb], g * 64, rowbase + i * 128);
      for (int tap0 = 0; tap0 < 128; tap0 += PC_TAPS, ++nw) {
        const int s = nw % PC_WSTAGES;
        mbar_wait(&w_empty[s], ((nw / PC_WSTAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&w_full[s], PC_W_BYTES);
#pragma unroll
        for (int t = 0; t < PC_TAPS; ++t)
          tma_load_2d(sW + s * PC_W_BYTES + t * PC_TAP_BYTES, &tmW, &w_full[s], (tap0 + t) * 64, g * 64);
      }
    }
  } else if (threadIdx.x == 160) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    uint32_t n = 0, nw = 0;
    for (int idx = blockIdx.x; idx < pc.n_items; idx += gridDim.x, ++n) {
      const int xb = n & 1, acc = n & 1;
      mbar_wait(&tempty[acc], ((n >> 1) & 1) ^ 1);
      mbar_wait(&x_full[xb], (n >> 1) & 1);
      tc_fence_after();
      const uint64_t dx = umma_desc_sw128(smem_u32(sX + xb * PC_XWIN_BYTES));
      for (int tap0 = 0; tap0 < 128; tap0 += PC_TAPS, ++nw) {
        const int s = nw % PC_WSTAGES;
        mbar_wait(&w_full[s], (nw / PC_WSTAGES) & 1);
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < PC_TAPS; ++t) {
          const int tap = tap0 + t;
          const uint64_t dw = umma_desc_sw128(smem_u32(sW + s * PC_W_BYTES + t * PC_TAP_BYTES));
          // window shifted by `tap` rows: +tap*128 B in the descriptor's start-address field (16-byte units)
          const uint64_t da = dx + (uint64_t)(tap * 8);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + acc * 128 + half * 64, da + (uint64_t)(half * 1024 + 2 * k), dw + 2 * k, idesc,
                        (tap | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&w_empty[s]);
      }
      umma_commit(&x_empty[xb]);
      umma_commit(&tfull[acc]);
    }
  } else if (warp < 4) {
    // ===================== epilogue warps =====================
    uint32_t n = 0;
    for (int idx = blockIdx.x; idx < pc.n_items; idx += gridDim.x, ++n) {
      int g, rowbase;
      decode(idx, g, rowbase);
      const int acc = n & 1;
      const uint32_t par = (n >> 1) & 1;
#pragma unroll 1
      for (int half = 0; half < 2; ++half)
        epilogue_tile<64, 1>(pc.g, stages[warp], tmem_base + acc * 128 + half * 64, warp,
                             rowbase + half * 128 + (int)warp * 32, g * 64, lane, 0, &tfull[acc], par);
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int launch_posconv(const PosConvOp& op, cudaStream_t stream, int num_sms, std::string& err) {
  CUtensorMap tmX, tmW;
  if (make_tmap_2d(&tmX, op.X, 1024ULL, (unsigned long long)op.x_rows, 1024ULL, 128, err)) return -1;
  if (make_tmap_2d(&tmW, op.W, 8192ULL, 1024ULL, 8192ULL, 64, err)) return -1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(posconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM_BYTES);
    if (ce != cudaSuccess) {
      err = std::string("cudaFuncSetAttribute(posconv_tc_kernel): ") + cudaGetErrorString(ce);
      return -1;
    }
    attr_set = true;
  }
  PosConvParams pc;
  pc.B = op.B;
  pc.pslot = op.pslot;
  pc.n_mt = ceil_div(op.rows_per_clip, 256);
  pc.n_items = op.B * pc.n_mt * 16;
  pc.g.M = op.B * op.pslot;
  pc.g.N = 1024;
  pc.g.K = 8192;
  pc.g.a_mode = 1;
  pc.g.num_m_tiles = pc.g.num_n_tiles = pc.g.num_kb = 0;
  pc.g.epi = op.epi;
  if ((op.epi.ldo32 & 3) || (op.epi.ldo16 & 7) || (op.epi.ldr & 3)) {
    err = "posconv: output / residual leading dimensions must keep 16-byte row alignment";
    return -1;
  }
  const int grid = pc.n_items < num_sms ? pc.n_items : num_sms;
  posconv_tc_kernel<<<grid, 192, PC_SMEM_BYTES, stream>>>(tmX, tmW, pc);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("posconv_tc_kernel launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
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
  gemm_tc_kernel<BN><<<grid, 192, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("gemm_tc_kernel launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

static int launch_tc2(const GemmOp& op, GemmParams& p, cudaStream_t stream, int num_sms, std::string& err) {
  CUtensorMap tmA, tmB;
  if (make_tmap_2d(&tmA, op.A, (unsigned long long)op.K, (unsigned long long)op.a_rows, (unsigned long long)op.lda,
                   128, err))
    return -1;
  if (make_tmap_2d(&tmB, op.W, (unsigned long long)op.K, (unsigned long long)op.N, (unsigned long long)op.K, 128,
                   err))
    return -1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES);
    if (ce != cudaSuccess) {
      err = std::string("cudaFuncSetAttribute(gemm_tc2_kernel): ") + cudaGetErrorString(ce);
      return -1;
    }
    attr_set = true;
  }
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int max_clusters = num_sms / 2;
  const int grid = 2 * (total < max_clusters ? total : max_clusters);
  cudaError_t ce = launch_pdl(gemm_tc2_kernel, dim3(grid), dim3(TC2_THREADS), TC2_SMEM_BYTES, stream, tmA, tmB, p);
  if (ce == cudaSuccess) ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("gemm_tc2_kernel launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

static bool force_single_cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSR_GEMM_1CTA");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
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
  // Few rows (the per-clip calls: M = 150 for one 3 s clip): 256-column tiles would leave all but N / 256 SM pairs
  // idle while each of them walks the whole K dimension; 64-column single-CTA tiles spread the weight matrix over
  // N / 64 SMs instead. The cut-over is where the big tiles start to fill the machine.
  if ((long long)ceil_div(op.M, 256) * (op.N / 64) <= 2LL * num_sms && op.M <= 1024) {
    p.num_n_tiles = op.N / 64;
    return launch_tc<64>(op, p, stream, num_sms, err);
  }
  // Less than one wave of 256 x 256 pair tiles (e.g. one 30 s Whisper window, M = 1500, N = 1280: 30 tiles for 74 SM
  // pairs): 128 x 128 single-CTA tiles fill the machine better.
  if (op.N % 256 == 0 && (long long)ceil_div(op.M, 256) * (op.N / 256) < num_sms / 2 && !force_single_cta()) {
    p.num_n_tiles = op.N / 128;
    return launch_tc<128>(op, p, stream, num_sms, err);
  }
  if (op.N % 256 == 0 && !force_single_cta()) {
    p.num_m_tiles = ceil_div(op.M, 256);
    p.num_n_tiles = op.N / 256;
    return launch_tc2(op, p, stream, num_sms, err);
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
