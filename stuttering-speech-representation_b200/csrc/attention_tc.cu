// Flash-style multi-head self-attention on the 5th-gen tensor cores (sm_100a), head_dim 64.
//
//   per CTA: one (clip, head, 128-query tile); loop over 128-key blocks
//     warp 4      TMA producer: Q once, then K_j / V_j tiles (double buffered) straight out of the fused qkv matrix
//     warp 5      single-thread tcgen05.mma issuer:  S = Q K_j^T  -> TMEM[0,128) ;  PV_j = P_j V_j -> TMEM[128,192)
//                 (driver warps carry the highest warp ids: the sub-partition arbiter favours them over softmax warps)
//     warps 0-3   softmax: thread = query row (TMEM lane). tcgen05.ld S, + gated relative-position bias, key mask,
//                 online max / sum in fp32, P (bf16) written to shared memory in the UMMA 128B-swizzled K-major
//                 layout; after PV_j completes the partial product is folded into the fp32 O accumulator kept in
//                 registers (so no TMEM read-modify-write is needed for the online-softmax rescale).
//   Two CTAs are co-resident per SM (112 KB smem, 256 TMEM columns each): while one runs its softmax on the CUDA
//   cores the other owns the tensor pipe.
//
// Reference arithmetic: see attention.cu (same math; that mma.sync kernel is kept for slot < 32 and as a cross-check).
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace ssr {

using namespace ptx;

namespace {

constexpr int QT = 128;   // queries per CTA
constexpr int KBLK = 128; // keys per block
constexpr int HD = 64;
constexpr int TILE_BYTES = 128 * 64 * 2;  // one [128 x 64] bf16 tile = 16 KB
constexpr int SM_Q = 0;
constexpr int SM_KV = TILE_BYTES;                  // 2 stages x (K, V)
constexpr int SM_P = SM_KV + 4 * TILE_BYTES;       // [128 x 128] bf16 as two K-major sub-tiles
constexpr int SM_BAR = SM_P + 2 * TILE_BYTES;
constexpr int ATT_SMEM = SM_BAR + 128;
constexpr int TMEM_COLS = 256;
constexpr int TM_S = 0, TM_O = 128;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MN-major (here: V, [keys x 64 d], d contiguous) 128B-swizzled operand: 8-row (K) groups are 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // LBO: next 64-wide MN atom (unused, N = 64)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO: next group of 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(192, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm, const AttentionArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int len = min(a.lens[b], a.slot);
  const int q0 = qt * QT;
  if (q0 >= len) return;  // rows past the clip's live frames are never consumed (see engine.cu: slot layout)
  const int nkb = (len + KBLK - 1) / KBLK;
  const int row0 = b * a.slot;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* bar_q = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* bar_s = bars + 5;
  uint64_t* bar_p = bars + 6;
  uint64_t* bar_o = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    prefetch_tmap(&tm);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (threadIdx.x == 128) {
    // ============================ TMA producer ============================
    mbar_arrive_expect_tx(bar_q, TILE_BYTES);
    tma_load_2d(smem + SM_Q, &tm, bar_q, h * HD, row0 + q0);
    for (int j = 0; j < nkb; ++j) {
      const int s = j & 1;
      mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
      mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
      tma_load_2d(smem + SM_KV + s * 2 * TILE_BYTES, &tm, &kv_full[s], a.D + h * HD, row0 + j * KBLK);
      tma_load_2d(smem + SM_KV + s * 2 * TILE_BYTES + TILE_BYTES, &tm, &kv_full[s], 2 * a.D + h * HD,
                  row0 + j * KBLK);
    }
  } else if (threadIdx.x == 160) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64) | (1u << 16);  // B (= V) is MN-major
    const uint64_t dq = umma_desc_sw128(smem_u32(smem + SM_Q));
    const uint64_t dp0 = umma_desc_sw128(smem_u32(smem + SM_P));
    const uint64_t dp1 = umma_desc_sw128(smem_u32(smem + SM_P + TILE_BYTES));
    mbar_wait(bar_q, 0);
    for (int j = 0; j < nkb; ++j) {
      const int s = j & 1;
      mbar_wait(&kv_full[s], (j >> 1) & 1);
      tc_fence_after();
      const uint64_t dk = umma_desc_sw128(smem_u32(smem + SM_KV + s * 2 * TILE_BYTES));
      const uint64_t dv = umma_desc_sw128_mn(smem_u32(smem + SM_KV + s * 2 * TILE_BYTES + TILE_BYTES));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + TM_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
      umma_commit(bar_s);
      mbar_wait(bar_p, j & 1);  // P_j is in shared memory (and O_{j-1} has been read back)
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t dp = (k < 4 ? dp0 : dp1) + 2 * (k & 3);
        // V: 16 keys per MMA = two 8-row groups = 2048 bytes
        umma_bf16(tmem + TM_O, dp, dv + (uint64_t)(k * 2048 >> 4), idesc_pv, k != 0);
      }
      umma_commit(&kv_empty[s]);
      umma_commit(bar_o);
    }
  } else if (warp < 4) {
    // ============================ softmax / output warps ============================
    const uint32_t quad = warp;                   // TMEM lane quadrant accessible to this warp
    const int il = quad * 32 + lane;              // query row inside the tile
    const int i = q0 + il;                        // query index inside the clip
    const uint32_t lane_addr = (quad * 32u) << 16;
    float gate = 0.f;
    const float* rel = nullptr;
    if (HAS_BIAS) {
      if (i < a.slot) gate = a.gate[((long long)row0 + i) * a.H + h];
      rel = a.relbias + (long long)h * a.rel_stride + a.rel_center - i;  // rel[j] = table[h][j - i]
    }
    float o[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = smem + SM_P + il * 128;
    const uint32_t sw = il & 7;

    for (int j = 0; j < nkb; ++j) {
      const int k0 = j * KBLK;
      const bool need_mask = k0 + KBLK > len;
      mbar_wait(bar_s, j & 1);
      __syncwarp();
      tc_fence_after();
      // ---- pass 1: block row maximum ----
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem + lane_addr + TM_S + c * 32, raw);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          float v = __uint_as_float(raw[k]);
          const int jg = k0 + c * 32 + k;
          if (HAS_BIAS) v = fmaf(gate, __ldg(rel + jg), v);
          if (need_mask && jg >= len) v = -INFINITY;
          m_blk = fmaxf(m_blk, v);
        }
      }
      const float m_new = fmaxf(m_run, m_blk);
      const float mu = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2_approx((m_run - mu) * LOG2E);
      const float mu2 = mu * LOG2E;
      m_run = m_new;
      // ---- pass 2: probabilities -> bf16 P tile in shared memory (128B-swizzled K-major) ----
      float l_blk = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem + lane_addr + TM_S + c * 32, raw);
        tmem_wait_ld();
        uint32_t packed[16];
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          float v0 = __uint_as_float(raw[k]), v1 = __uint_as_float(raw[k + 1]);
          const int jg = k0 + c * 32 + k;
          if (HAS_BIAS) {
            v0 = fmaf(gate, __ldg(rel + jg), v0);
            v1 = fmaf(gate, __ldg(rel + jg + 1), v1);
          }
          if (need_mask) {
            if (jg >= len) v0 = -INFINITY;
            if (jg + 1 >= len) v1 = -INFINITY;
          }
          const float p0 = ex2_approx(fmaf(v0, LOG2E, -mu2));
          const float p1 = ex2_approx(fmaf(v1, LOG2E, -mu2));
          l_blk += p0 + p1;
          __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
          packed[k >> 1] = *reinterpret_cast<uint32_t*>(&pk);
        }
        // chunk c covers keys [c*32, c*32+32): sub-tile c/2, 16-byte columns (c%2)*4 .. +3, XOR-swizzled by row%8
        uint8_t* sub = prow + (c >> 1) * TILE_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t col = ((c & 1) * 4 + q) ^ sw;
          *reinterpret_cast<uint4*>(sub + col * 16) =
              make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
      }
      l_run = l_run * alpha + l_blk;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      // ---- fold PV_j into the register accumulator ----
      mbar_wait(bar_o, j & 1);
      __syncwarp();
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem + lane_addr + TM_O + c * 32, raw);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) o[c * 32 + k] = fmaf(o[c * 32 + k], alpha, __uint_as_float(raw[k]));
      }
      tc_fence_before();
    }
    // ---- normalise and store this row ----
    if (i < a.slot) {
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(a.out + ((long long)row0 + i) * a.D + h * HD);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 pk = __floats2bfloat162_rn(o[q * 8 + 2 * e] * inv, o[q * 8 + 2 * e + 1] * inv);
          w[e] = *reinterpret_cast<uint32_t*>(&pk);
        }
        dst[q] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

}  // namespace

int launch_attention_tc(const AttentionArgs& a, cudaStream_t st, std::string& err) {
  if (a.D != a.H * HD) {
    err = "attention: head_dim must be 64";
    return -1;
  }
  if (a.B <= 0 || a.slot <= 0) return 0;
  CUtensorMap tm;
  if (make_tmap_2d(&tm, a.qkv, 3ULL * a.D, (unsigned long long)a.B * a.slot, 3ULL * a.D, 128, err)) return -1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t c1 =
        cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    cudaError_t c2 =
        cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (c1 != cudaSuccess || c2 != cudaSuccess) {
      err = std::string("cudaFuncSetAttribute(attention_tc_kernel): ") +
            cudaGetErrorString(c1 != cudaSuccess ? c1 : c2);
      return -1;
    }
    attr_set = true;
  }
  dim3 grid(ceil_div(a.slot, QT), a.H, a.B);
  if (a.gate != nullptr)
    attention_tc_kernel<true><<<grid, 192, ATT_SMEM, st>>>(tm, a);
  else
    attention_tc_kernel<false><<<grid, 192, ATT_SMEM, st>>>(tm, a);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("attention_tc launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

}  // namespace ssr
