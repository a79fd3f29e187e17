// Flash-style multi-head self-attention on the 5th-gen tensor cores (sm_100a), head_dim 64.
//
// Persistent, warp-specialised kernel; a work item is one (clip, head, 128-query tile); each item loops over
// 128-key blocks:
//     warp 4      TMA producer: Q per item, K_j / V_j tiles through a 2-stage ring that runs ahead across items,
//                 all straight out of the fused qkv activation matrix (no head split / transpose pass)
//     warp 5      single-thread tcgen05.mma issuer:  S = Q K_j^T  -> TMEM[0,128) ;  PV_j = P_j V_j -> TMEM[128,192)
//                 (driver warps carry the highest warp ids: the sub-partition arbiter favours them)
//     warps 0-3   softmax: thread = query row (TMEM lane). tcgen05.ld S, + gated relative-position bias, key mask,
//                 online max / sum in fp32, P (bf16) written to shared memory in the UMMA 128B-swizzled K-major
//                 layout; after PV_j completes the partial product is folded into the fp32 O accumulator kept in
//                 registers (so no TMEM read-modify-write is needed for the online-softmax rescale).
//   Two CTAs are co-resident per SM (112 KB smem, 256 TMEM columns each): while one runs its softmax on the CUDA
//   cores the other owns the tensor pipe. Padding is not computed: the last key block uses an MMA N / K extent
//   rounded to 16 live keys, softmax touches only the 32-column chunks that hold live keys, and warps whose 32 query
//   rows are all beyond the clip's length only keep the barrier protocol going.
//
// Reference arithmetic: see attention.cu (same math; that mma.sync kernel is kept as a cross-check).
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace ssr {

using namespace ptx;

namespace {

constexpr int QT = 128;    // queries per item
constexpr int KBLK = 128;  // keys per block
constexpr int HD = 64;
constexpr int TILE_BYTES = 128 * 64 * 2;  // one [128 x 64] bf16 tile = 16 KB
constexpr int SM_Q = 0;
constexpr int SM_KV = TILE_BYTES;             // 2 stages x (K, V)
constexpr int SM_P = SM_KV + 4 * TILE_BYTES;  // [128 x 128] bf16 as two K-major sub-tiles
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

struct Item {
  int b, h, q0, len, nkb;
  bool valid;
};

// Items are ordered query-tile-major (all first tiles, then all second tiles, ...) so that the statically strided
// persistent CTAs each see the same mix of full and partial tiles.
__device__ __forceinline__ Item decode_item(const AttentionArgs& a, int idx) {
  Item it;
  const int bh = a.B * a.H;
  const int qt = idx / bh;
  const int rem = idx - qt * bh;
  it.b = rem / a.H;
  it.h = rem - it.b * a.H;
  it.q0 = qt * QT;
  it.len = min(__ldg(a.lens + it.b), a.slot);
  it.valid = it.q0 < it.len;  // tiles past the clip's live frames are never consumed (engine.cu: slot layout)
  it.nkb = (it.len + KBLK - 1) / KBLK;
  return it;
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(192, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm, const AttentionArgs a, const int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;   // [2]
  uint64_t* kv_empty = bars + 4;  // [2]
  uint64_t* bar_s = bars + 6;
  uint64_t* bar_p = bars + 7;
  uint64_t* bar_o = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    prefetch_tmap(&tm);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
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
    uint32_t n_item = 0, n_kv = 0;
    Item nxt = decode_item(a, blockIdx.x);
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
      const Item it = nxt;
      // the next item's length load is issued now and consumed one iteration later (keeps it off the critical path)
      if (idx + (int)gridDim.x < n_items) nxt = decode_item(a, idx + gridDim.x);
      if (!it.valid) continue;
      const int row0 = it.b * a.slot;
      mbar_wait(q_empty, (n_item & 1) ^ 1);
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(smem + SM_Q, &tm, q_full, it.h * HD, row0 + it.q0);
      for (int j = 0; j < it.nkb; ++j, ++n_kv) {
        const int s = n_kv & 1;
        mbar_wait(&kv_empty[s], ((n_kv >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
        tma_load_2d(smem + SM_KV + s * 2 * TILE_BYTES, &tm, &kv_full[s], a.D + it.h * HD, row0 + j * KBLK);
        tma_load_2d(smem + SM_KV + s * 2 * TILE_BYTES + TILE_BYTES, &tm, &kv_full[s], 2 * a.D + it.h * HD,
                    row0 + j * KBLK);
      }
      ++n_item;
    }
  } else if (threadIdx.x == 160) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64) | (1u << 16);  // B (= V) is MN-major
    const uint64_t dq = umma_desc_sw128(smem_u32(smem + SM_Q));
    const uint64_t dp0 = umma_desc_sw128(smem_u32(smem + SM_P));
    const uint64_t dp1 = umma_desc_sw128(smem_u32(smem + SM_P + TILE_BYTES));
    uint32_t n_item = 0, n_kv = 0;
    Item nxt = decode_item(a, blockIdx.x);
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
      const Item it = nxt;
      if (idx + (int)gridDim.x < n_items) nxt = decode_item(a, idx + gridDim.x);
      if (!it.valid) continue;
      mbar_wait(q_full, n_item & 1);
      for (int j = 0; j < it.nkb; ++j, ++n_kv) {
        const int s = n_kv & 1;
        const int nlive = min(KBLK, it.len - j * KBLK);
        const int n16 = (nlive + 15) >> 4;  // live keys in units of 16
        mbar_wait(&kv_full[s], (n_kv >> 1) & 1);
        tc_fence_after();
        const uint64_t dk = umma_desc_sw128(smem_u32(smem + SM_KV + s * 2 * TILE_BYTES));
        const uint64_t dv = umma_desc_sw128_mn(smem_u32(smem + SM_KV + s * 2 * TILE_BYTES + TILE_BYTES));
        const uint32_t idesc_s = umma_idesc_bf16(128, n16 * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + TM_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        if (j == it.nkb - 1) umma_commit(q_empty);  // Q tile may be overwritten once these MMAs retire
        umma_commit(bar_s);
        mbar_wait(bar_p, n_kv & 1);  // P_j is in shared memory (and O_{j-1} has been read back)
        tc_fence_after();
        for (int k = 0; k < n16; ++k) {
          const uint64_t dp = (k < 4 ? dp0 : dp1) + 2 * (k & 3);
          // V: 16 keys per MMA = two 8-row groups = 2048 bytes
          umma_bf16(tmem + TM_O, dp, dv + (uint64_t)(k * 2048 >> 4), idesc_pv, k != 0);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(bar_o);
      }
      ++n_item;
    }
  } else if (warp < 4) {
    // ============================ softmax / output warps ============================
    const uint32_t quad = warp;  // TMEM lane quadrant accessible to this warp
    const int il = quad * 32 + lane;
    const uint32_t lane_addr = (quad * 32u) << 16;
    uint8_t* prow = smem + SM_P + il * 128;
    const uint32_t sw = il & 7;
    uint32_t n_kv = 0;
    Item nxt = decode_item(a, blockIdx.x);
    float gate_nxt = 0.f;
    if (HAS_BIAS && nxt.q0 + il < a.slot)
      gate_nxt = a.gate[((long long)nxt.b * a.slot + nxt.q0 + il) * a.H + nxt.h];
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
      const Item it = nxt;
      const float gate = gate_nxt;
      // prefetch the next item's length and this row's gate: consumed one iteration later
      if (idx + (int)gridDim.x < n_items) {
        nxt = decode_item(a, idx + gridDim.x);
        if (HAS_BIAS && nxt.q0 + il < a.slot)
          gate_nxt = a.gate[((long long)nxt.b * a.slot + nxt.q0 + il) * a.H + nxt.h];
      }
      if (!it.valid) continue;
      const int row0 = it.b * a.slot;
      const int i = it.q0 + il;                                   // query index inside the clip
      const bool warp_live = it.q0 + (int)quad * 32 < it.len;      // at least one live query row in this warp
      const float* rel = nullptr;
      if (HAS_BIAS) rel = a.relbias + (long long)it.h * a.rel_stride + a.rel_center - i;  // rel[j] = table[h][j - i]
      float o[64];
#pragma unroll
      for (int d = 0; d < 64; ++d) o[d] = 0.f;
      float m_run = -INFINITY, l_run = 0.f;

      for (int j = 0; j < it.nkb; ++j, ++n_kv) {
        const int k0 = j * KBLK;
        const int nlive = min(KBLK, it.len - k0);
        const int nch = (nlive + 31) >> 5;  // 32-column chunks that hold live keys
        const bool need_mask = (nlive & 31) != 0;
        mbar_wait(bar_s, n_kv & 1);
        __syncwarp();
        tc_fence_after();
        float alpha = 1.f;
        if (warp_live) {
          // ---- pass 1: block row maximum ----
          float m_blk = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < nch; ++c) {
            uint32_t raw[32];
            tmem_ld_32x32(tmem + lane_addr + TM_S + c * 32, raw);
            tmem_wait_ld();
            const bool mask_here = need_mask && (c == nch - 1);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              float v = __uint_as_float(raw[k]);
              const int jg = k0 + c * 32 + k;
              if (HAS_BIAS) v = fmaf(gate, __ldg(rel + jg), v);
              if (mask_here && jg >= it.len) v = -INFINITY;
              m_blk = fmaxf(m_blk, v);
            }
          }
          const float m_new = fmaxf(m_run, m_blk);
          const float mu = (m_new == -INFINITY) ? 0.f : m_new;
          alpha = ex2_approx((m_run - mu) * LOG2E);
          const float mu2 = mu * LOG2E;
          m_run = m_new;
          // ---- pass 2: probabilities -> bf16 P tile in shared memory (128B-swizzled K-major) ----
          float l_blk = 0.f;
#pragma unroll 1
          for (int c = 0; c < nch; ++c) {
            uint32_t raw[32];
            tmem_ld_32x32(tmem + lane_addr + TM_S + c * 32, raw);
            tmem_wait_ld();
            const bool mask_here = need_mask && (c == nch - 1);
            uint32_t packed[16];
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              float v0 = __uint_as_float(raw[k]), v1 = __uint_as_float(raw[k + 1]);
              const int jg = k0 + c * 32 + k;
              if (HAS_BIAS) {
                v0 = fmaf(gate, __ldg(rel + jg), v0);
                v1 = fmaf(gate, __ldg(rel + jg + 1), v1);
              }
              if (mask_here) {
                if (jg >= it.len) v0 = -INFINITY;
                if (jg + 1 >= it.len) v1 = -INFINITY;
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
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        // ---- fold PV_j into the register accumulator ----
        mbar_wait(bar_o, n_kv & 1);
        __syncwarp();
        tc_fence_after();
        if (warp_live) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t raw[32];
            tmem_ld_32x32(tmem + lane_addr + TM_O + c * 32, raw);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) o[c * 32 + k] = fmaf(o[c * 32 + k], alpha, __uint_as_float(raw[k]));
          }
        }
        tc_fence_before();
      }
      // ---- normalise and store this row (rows at or beyond the clip length are never read downstream) ----
      if (i < it.len) {
        const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
        uint4* dst = reinterpret_cast<uint4*>(a.out + ((long long)row0 + i) * a.D + it.h * HD);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
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
  static int num_sms = 148;
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
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_set = true;
  }
  const long long items = (long long)ceil_div(a.slot, QT) * a.H * a.B;
  if (items > 2000000000LL) {
    err = "attention: too many work items";
    return -1;
  }
  const int grid = (int)(items < 2LL * num_sms ? items : 2LL * num_sms);
  if (a.gate != nullptr)
    attention_tc_kernel<true><<<grid, 192, ATT_SMEM, st>>>(tm, a, (int)items);
  else
    attention_tc_kernel<false><<<grid, 192, ATT_SMEM, st>>>(tm, a, (int)items);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("attention_tc launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

}  // namespace ssr
