// Conv1d (implicit GEMM, N = 512 output channels) + LayerNorm over the 512 channels + erf-GELU in ONE kernel, for the
// WavLM-Large feature encoder layers 1-6 (HF/models/wavlm/modeling_wavlm.py:703-727: conv -> transpose -> LayerNorm ->
// transpose -> GELU). The plain path runs gemm_tc2_kernel (raw conv output to HBM as bf16) and then a row kernel that
// reads it back, normalises, applies GELU and writes it again; here the normalisation happens on the fp32 accumulators.
//
// A LayerNorm row spans 512 channels = two 256-column accumulator tiles, so a 4-CTA cluster works on one 256-row
// M tile: CTA pair {0,1} (cta_group::2, M = 256) owns channels 0-255, pair {2,3} channels 256-511. Mainloop as in
// gemm_tc2_kernel (TMA 128B swizzle -> 5-stage ring -> tcgen05.mma, accumulators double-buffered in TMEM). Epilogue:
//   pass 1  tcgen05.ld the accumulator, per-row sum / sum of squares over the CTA's 256 columns;
//   exchange the 128 row partials go to the CTA holding the same rows of the other channel half (rank ^ 2) through
//           distributed shared memory (st.shared::cluster + mbarrier arrive with cluster-scope release / acquire);
//   pass 2  tcgen05.ld again, (x - mean) * rstd * gamma + beta, GELU, bf16, coalesced store via a staging tile.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace ssr {

using namespace ptx;

namespace {

constexpr int BK = 64;
constexpr int A_BYTES = 128 * BK * 2;  // this CTA's 128 rows of the A tile
constexpr int B_BYTES = 128 * BK * 2;  // this CTA's half of the pair's 256-column B tile
constexpr int STAGES = 5;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int OFF_BAR = STAGES * (A_BYTES + B_BYTES);
constexpr int N_BARS = 2 * STAGES + 6;                       // full, empty, tfull[2], tempty[2], xfull[2]
constexpr int OFF_GB = OFF_BAR + ((N_BARS * 8 + 16 + 15) & ~15);  // gamma[256] | beta[256] of this CTA's channels
constexpr int OFF_PART = OFF_GB + 2 * 256 * 4;               // [2 accumulators][2 column halves][128 rows] float2
constexpr int OFF_XPART = OFF_PART + 4 * 128 * 8;            // [2 accumulators][128 rows] float2 (partner's partials)
constexpr int OFF_STAGE = OFF_XPART + 2 * 128 * 8;           // per-warp bf16 staging: 32 rows x 80 bytes
constexpr int STAGE_BYTES = 32 * 80;
constexpr int SMEM_BYTES = OFF_STAGE + EPI_WARPS * STAGE_BYTES;

struct LnGemmParams {
  int M, num_m_tiles, num_kb;
  const float* gamma;
  const float* beta;
  float eps;
  bf16* out;  // [M, 512]
};

__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 20)) __trap();  // protocol bug: fail the launch instead of hanging the box
  }
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const LnGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full = bars;                     // live copy: the pair leader's
  uint64_t* empty = bars + STAGES;           // per CTA (multicast commit)
  uint64_t* tfull = bars + 2 * STAGES;       // per CTA (multicast commit)
  uint64_t* tempty = bars + 2 * STAGES + 2;  // live copy: the pair leader's
  uint64_t* xfull = bars + 2 * STAGES + 4;   // per CTA: the partner's row partials have landed in xpart[acc]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  float* gam = reinterpret_cast<float*>(smem + OFF_GB);
  float* bet = gam + 256;
  float2* part = reinterpret_cast<float2*>(smem + OFF_PART);
  float2* xpart = reinterpret_cast<float2*>(smem + OFF_XPART);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0..3
  const uint32_t prank = rank & 1;          // position inside the CTA pair
  const uint32_t leader = rank & ~1u;       // the pair's leader CTA
  const uint32_t pair = rank >> 1;          // channel half: 0 -> 0..255, 1 -> 256..511
  const int cid = blockIdx.x >> 2, ncl = gridDim.x >> 2;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * EPI_WARPS);
      mbar_init(&xfull[i], 4);  // one arrival per row quadrant of the partner CTA
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 256; i += THREADS) {
    gam[i] = p.gamma[pair * 256 + i];
    bet[i] = p.beta[pair * 256 + i];
  }
  if (warp == EPI_WARPS + 1) {
    tmem_alloc_cg2(tmem_slot, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == EPI_WARPS * 32) {
    // ===================== TMA producer =====================
    int s = 0;
    uint32_t ph = 0;
    for (int tile = cid; tile < p.num_m_tiles; tile += ncl) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        if (prank == 0) mbar_arrive_expect_tx(&full[s], 2 * (A_BYTES + B_BYTES));
        tma_load_2d_cg2(sA + s * A_BYTES, &tmA, &full[s], kb * BK, tile * 256 + (int)prank * 128);
        tma_load_2d_cg2(sB + s * B_BYTES, &tmB, &full[s], kb * BK, (int)pair * 256 + (int)prank * 128);
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (threadIdx.x == (EPI_WARPS + 1) * 32 && prank == 0) {
    // ===================== MMA issuer (pair leader) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(256, 256);
    const uint16_t mask = (uint16_t)(3u << leader);
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = cid; tile < p.num_m_tiles; tile += ncl) {
      mbar_wait(&tempty[acc], accph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem_u32(sA + s * A_BYTES));
        const uint64_t db = umma_desc_sw128(smem_u32(sB + s * B_BYTES));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit_cg2(&empty[s], mask);
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
      umma_commit_cg2(&tfull[acc], mask);
      acc ^= 1;
      if (acc == 0) accph ^= 1;
    }
  } else if (warp < EPI_WARPS) {
    // ===================== epilogue warps =====================
    const uint32_t quad = warp & 3;     // TMEM lane quadrant = 32 rows
    const int half = (int)(warp >> 2);  // which 128 of this CTA's 256 columns
    const int row = (int)quad * 32 + (int)lane;
    uint8_t* stg = smem + OFF_STAGE + warp * STAGE_BYTES;
    const uint32_t partner = rank ^ 2u;
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = cid; tile < p.num_m_tiles; tile += ncl) {
      const uint32_t tacc = tmem_base + acc * 256 + ((quad * 32u) << 16) + half * 128;
      mbar_wait(&tfull[acc], accph);
      __syncwarp();
      tc_fence_after();
      // ---- pass 1: row partial sums over this warp's 128 columns
      float s = 0.f, q = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tacc + c * 32, raw);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = __uint_as_float(raw[i]);
          s += v;
          q = fmaf(v, v, q);
        }
      }
      part[(acc * 2 + half) * 128 + row] = make_float2(s, q);
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");  // epilogue warps only
      const float2 p0 = part[acc * 256 + row], p1 = part[acc * 256 + 128 + row];
      const float cs = p0.x + p1.x, cq = p0.y + p1.y;  // this CTA's 256 columns
      if (half == 0) {
        st_cluster_f2(mapa_u32(&xpart[acc * 128 + row], partner), cs, cq);
        __syncwarp();
        if (lane == 0) mbar_arrive_release_cluster(mapa_u32(&xfull[acc], partner));
      }
      mbar_wait_acquire_cluster(&xfull[acc], accph);
      const float2 xp = xpart[acc * 128 + row];
      const float mean = (cs + xp.x) * (1.0f / 512.f);
      const float var = fmaxf((cq + xp.y) * (1.0f / 512.f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      const float shift = -mean * rstd;
      // ---- pass 2: normalise, GELU, bf16, coalesced store
      const long long grow0 = (long long)tile * 256 + (long long)prank * 128 + (long long)quad * 32;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tacc + c * 32, raw);
        tmem_wait_ld();
        const int col0 = half * 128 + c * 32;
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float g0 = gam[col0 + i], g1 = gam[col0 + i + 1];
          const float y0 = fmaf(fmaf(__uint_as_float(raw[i]), rstd, shift), g0, bet[col0 + i]);
          const float y1 = fmaf(fmaf(__uint_as_float(raw[i + 1]), rstd, shift), g1, bet[col0 + i + 1]);
          __nv_bfloat162 pk = __floats2bfloat162_rn(gelu_fast(y0), gelu_fast(y1));
          packed[i >> 1] = *reinterpret_cast<uint32_t*>(&pk);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(stg + lane * 80 + i * 16) =
              make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = it * 8 + (int)(lane >> 2);
          const long long grow = grow0 + r;
          if (grow < p.M)
            *reinterpret_cast<uint4*>(p.out + grow * 512 + (long long)pair * 256 + col0 + (lane & 3) * 8) =
                *reinterpret_cast<const uint4*>(stg + r * 80 + (lane & 3) * 16);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[acc], leader);
      acc ^= 1;
      if (acc == 0) accph ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while a peer can still touch its smem / TMEM
  if (warp == EPI_WARPS + 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

}  // namespace

// A: channels-last bf16 signal viewed as overlapping rows (row pitch lda, row length K); W: [512, K] bf16.
int launch_gemm_ln(const bf16* A, long long lda, long long a_rows, const bf16* W, int M, int K, const float* gamma,
                   const float* beta, float eps, bf16* out, cudaStream_t st, int num_sms, std::string& err) {
  if (M <= 0) return 0;
  if (K % BK != 0) {
    err = "gemm_ln: K must be a multiple of 64";
    return -1;
  }
  CUtensorMap tmA, tmB;
  if (make_tmap_2d(&tmA, A, (unsigned long long)K, (unsigned long long)a_rows, (unsigned long long)lda, 128, err))
    return -1;
  if (make_tmap_2d(&tmB, W, (unsigned long long)K, 512ULL, (unsigned long long)K, 128, err)) return -1;
  static int max_clusters = -1;
  if (max_clusters < 0) {
    cudaError_t ce = cudaFuncSetAttribute(gemm_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (ce != cudaSuccess) {
      err = std::string("cudaFuncSetAttribute(gemm_ln_kernel): ") + cudaGetErrorString(ce);
      return -1;
    }
    // how many 4-CTA clusters fit at once (GPC boundaries can leave a few SMs out)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(num_sms / 4 * 4));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_ln_kernel, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = num_sms / 4 - 3;
    }
    max_clusters = n < num_sms / 4 ? n : num_sms / 4;
  }
  LnGemmParams p;
  p.M = M;
  p.num_m_tiles = ceil_div(M, 256);
  p.num_kb = K / BK;
  p.gamma = gamma;
  p.beta = beta;
  p.eps = eps;
  p.out = out;
  const int clusters = p.num_m_tiles < max_clusters ? p.num_m_tiles : max_clusters;
  gemm_ln_kernel<<<4 * clusters, THREADS, SMEM_BYTES, st>>>(tmA, tmB, p);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("gemm_ln_kernel launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

}  // namespace ssr
