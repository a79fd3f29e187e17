// Flash-style multi-head self-attention on the 5th-gen tensor cores (sm_100a), head_dim 64.
//
// Persistent, warp-specialised kernel; a work item is one (clip, head, 128-query tile); each item loops over
// 64-key blocks. Everything is double buffered so that the softmax warps never wait for a tensor-core round trip
// in steady state:
//     warp 4      TMA producer: Q per item, K_j / V_j tiles through a 4-stage ring that runs ahead across items,
//                 all straight out of the fused qkv activation matrix (no head split / transpose pass)
//     warp 5      single-thread tcgen05.mma issuer. S_{n+1} = Q K^T is issued BEFORE PV_n, so the next block's scores
//                 are computed while the softmax warps work on the current one:
//                     S_n  -> TMEM S[n&1] (64 columns) ;  O += P_n V_n -> TMEM O (64 columns, one per item)
//                 (driver warps carry the highest warp ids: the sub-partition arbiter favours them)
//     warps 0-3   softmax: thread = query row (TMEM lane). tcgen05.ld S, + gated relative-position bias, key mask,
//                 fp32 sum, P (bf16) -> shared memory P[n&1] in the UMMA 128B-swizzled K-major layout.
//                 O accumulates in TMEM over the whole item against the block-0 reference maximum (exact row maximum
//                 of the first block); no per-element maximum is tracked afterwards. A stale reference only shifts
//                 the exponent (bf16 and fp32 share its range); if a block's partial sum leaves the safe range the
//                 warp rescales its rows of O in TMEM in place and redoes the block (rare). The softmax warps wait for
//                 a tensor-core round trip once per item (PV of the last block). Measured on B200: Whisper shape
//                 (B=64, T=1500, H=20) 1.91 ms -> 1.29 ms per layer on peaked scores, 1.50 ms on flat scores (denser
//                 P: the kernel is power-limited, 1.78 of 1.965 GHz under ncu); WavLM shape (B=256, T=149, H=16, gated
//                 bias) 183 -> 168 us versus a per-block two-pass online softmax with a register accumulator.
//   Two CTAs are co-resident per SM (112 KB smem, 256 TMEM columns each). Padding is not computed: the last key
//   block uses an MMA N / K extent rounded to 16 live keys, softmax touches only the 32-column chunks that hold live
//   keys, and warps whose 32 query rows are all beyond the clip's length only keep the barrier protocol going.
//
// Reference arithmetic: see attention.cu (same math; that mma.sync kernel is kept as a cross-check).
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace ssr {

using namespace ptx;

namespace {

constexpr int QT = 128;   // queries per item
constexpr int KBLK = 64;  // keys per block
constexpr int HD = 64;
constexpr int KV_STAGES = 4;
constexpr int Q_BYTES = 128 * 64 * 2;   // [128 x 64] bf16
constexpr int KV_BYTES = 64 * 64 * 2;   // [64 x 64] bf16 (K or V of one block)
constexpr int P_BYTES = 128 * 64 * 2;   // [128 q x 64 keys] bf16
constexpr int SM_Q = 0;
constexpr int SM_KV = Q_BYTES;                               // KV_STAGES x (K, V)
constexpr int SM_P = SM_KV + KV_STAGES * 2 * KV_BYTES;       // 2 buffers
constexpr int SM_BAR = SM_P + 2 * P_BYTES;
constexpr int ATT_SMEM = SM_BAR + 256;
constexpr int TMEM_COLS = 256;
constexpr int TM_S = 0, TM_O = 128;  // S[2] at columns 0 / 64, O at columns 128..191 (allocation is a power of two)
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


// One 32-key chunk of one query row: scores (+ gated relative-position bias, + key mask) -> running block max and,
// if WRITE_P, probabilities exp(s - ref) accumulated into l_blk and stored as bf16 into the row's P tile columns
// col0 .. col0+3 (16-byte units, XOR-swizzled by row % 8 = the UMMA / TMA 128B swizzle).
// WRITE_BACK keeps the biased / masked scores in `raw`, so that a second pass over the same registers needs neither
// the bias table nor the mask again.
template <bool HAS_BIAS, bool MASK, bool WRITE_P, bool TRACK_MAX = true, bool WRITE_BACK = false>
__device__ __forceinline__ void chunk(uint32_t (&raw)[32], int jg0, int len, float gate, const float* rel, float mu2,
                                      float& m_blk, float& l_blk, uint8_t* prow, uint32_t col0) {
  uint32_t packed[16];
#pragma unroll
  for (int k = 0; k < 32; k += 2) {
    float v0 = __uint_as_float(raw[k]), v1 = __uint_as_float(raw[k + 1]);
    if (HAS_BIAS) {
      v0 = fmaf(gate, __ldg(rel + jg0 + k), v0);
      v1 = fmaf(gate, __ldg(rel + jg0 + k + 1), v1);
    }
    if (MASK) {
      if (jg0 + k >= len) v0 = -INFINITY;
      if (jg0 + k + 1 >= len) v1 = -INFINITY;
    }
    if (WRITE_BACK) {
      raw[k] = __float_as_uint(v0);
      raw[k + 1] = __float_as_uint(v1);
    }
    if (TRACK_MAX) m_blk = fmaxf(m_blk, fmaxf(v0, v1));
    if (WRITE_P) {
      const float p0 = ex2_approx(fmaf(v0, LOG2E, -mu2));
      const float p1 = ex2_approx(fmaf(v1, LOG2E, -mu2));
      l_blk += p0 + p1;
      __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
      packed[k >> 1] = *reinterpret_cast<uint32_t*>(&pk);
    }
  }
  if (WRITE_P) {
    const uint32_t sw = (uint32_t)((reinterpret_cast<uintptr_t>(prow) >> 7) & 7);  // row % 8 (rows are 128 B apart)
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(prow + ((col0 + q) ^ sw) * 16) =
          make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
  }
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(192, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmkv,
                    const AttentionArgs a, const int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;               // [KV_STAGES]
  uint64_t* kv_empty = bars + 2 + KV_STAGES;  // [KV_STAGES]
  uint64_t* bar_s = bars + 2 + 2 * KV_STAGES;  // [2]
  uint64_t* bar_p = bar_s + 2;                 // [2]
  uint64_t* bar_o = bar_p + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 2);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    prefetch_tmap(&tmq);
    prefetch_tmap(&tmkv);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_p[i], 4);
      mbar_init(&bar_o[i], 1);
    }
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
    uint32_t n_item = 0, n = 0;
    Item nxt = decode_item(a, blockIdx.x);
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
      const Item it = nxt;
      if (idx + (int)gridDim.x < n_items) nxt = decode_item(a, idx + gridDim.x);  // prefetch the length load
      if (!it.valid) continue;
      const int row0 = it.b * a.slot;
      mbar_wait(q_empty, (n_item & 1) ^ 1);
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_2d(smem + SM_Q, &tmq, q_full, it.h * HD, row0 + it.q0);
      for (int j = 0; j < it.nkb; ++j, ++n) {
        const int s = n % KV_STAGES;
        mbar_wait(&kv_empty[s], ((n / KV_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * KV_BYTES);
        tma_load_2d(smem + SM_KV + s * 2 * KV_BYTES, &tmkv, &kv_full[s], a.D + it.h * HD, row0 + j * KBLK);
        tma_load_2d(smem + SM_KV + s * 2 * KV_BYTES + KV_BYTES, &tmkv, &kv_full[s], 2 * a.D + it.h * HD,
                    row0 + j * KBLK);
      }
      ++n_item;
    }
  } else if (threadIdx.x == 160) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64) | (1u << 16);  // B (= V) is MN-major
    const uint64_t dq = umma_desc_sw128(smem_u32(smem + SM_Q));
    uint32_t n_item = 0, n = 0;
    bool have_prev = false, prev_first = false;
    int prev_n16 = 0;
    // PV of block n-1 is issued after S of block n and accumulates into O over the whole item (the softmax warps
    // rescale O in place on the rare occasion the softmax reference changes).
    auto issue_pv_prev = [&]() {
      const uint32_t pn = n - 1;
      const int ps = pn % KV_STAGES;
      mbar_wait(&bar_p[pn & 1], (pn >> 1) & 1);  // P_{n-1} is in shared memory, O is ready to take PV_{n-1}
      tc_fence_after();
      const uint64_t dp = umma_desc_sw128(smem_u32(smem + SM_P + (pn & 1) * P_BYTES));
      const uint64_t dv = umma_desc_sw128_mn(smem_u32(smem + SM_KV + ps * 2 * KV_BYTES + KV_BYTES));
      const uint32_t d_o = tmem + TM_O;
      const bool keep = !prev_first;
      for (int k = 0; k < prev_n16; ++k)  // 16 keys per MMA: P advances 32 B, V two 8-row groups = 2048 B
        umma_bf16(d_o, dp + 2 * k, dv + (uint64_t)(k * 2048 >> 4), idesc_pv, (keep || k != 0) ? 1u : 0u);
      umma_commit(&kv_empty[ps]);
      umma_commit(&bar_o[pn & 1]);
    };
    Item nxt = decode_item(a, blockIdx.x);
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
      const Item it = nxt;
      if (idx + (int)gridDim.x < n_items) nxt = decode_item(a, idx + gridDim.x);
      if (!it.valid) continue;
      mbar_wait(q_full, n_item & 1);
      for (int j = 0; j < it.nkb; ++j) {
        const int s = n % KV_STAGES;
        const int nlive = min(KBLK, it.len - j * KBLK);
        const int n16 = (nlive + 15) >> 4;  // live keys in units of 16
        mbar_wait(&kv_full[s], (n / KV_STAGES) & 1);
        tc_fence_after();
        const uint64_t dk = umma_desc_sw128(smem_u32(smem + SM_KV + s * 2 * KV_BYTES));
        const uint32_t idesc_s = umma_idesc_bf16(128, n16 * 16);
        // S[n&1] was last read by the softmax of block n-2, which finished before bar_p of block n-2 (waited below)
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + TM_S + (n & 1) * 64, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        if (j == it.nkb - 1) umma_commit(q_empty);  // Q tile may be overwritten once these MMAs retire
        umma_commit(&bar_s[n & 1]);
        if (have_prev) issue_pv_prev();
        have_prev = true;
        prev_n16 = n16;
        prev_first = (j == 0);
        ++n;
      }
      ++n_item;
    }
    if (have_prev) issue_pv_prev();
  } else if (warp < 4) {
    // ============================ softmax / output warps ============================
    const uint32_t quad = warp;  // TMEM lane quadrant accessible to this warp
    const int il = quad * 32 + lane;
    const uint32_t lane_addr = (quad * 32u) << 16;
    uint32_t n = 0;
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
      const int i = it.q0 + il;                                // query index inside the clip
      const bool warp_live = it.q0 + (int)quad * 32 < it.len;   // at least one live query row in this warp
      const float* rel = nullptr;
      if (HAS_BIAS) rel = a.relbias + (long long)it.h * a.rel_stride + a.rel_center - i;  // rel[j] = table[h][j - i]
      {
        // ---------------- O accumulates in TMEM over the whole item ----------
        // Block 0 takes the exact row maximum as the softmax reference. Later blocks exponentiate against that
        // (possibly stale) reference without tracking a maximum at all: any shift of the reference is mathematically
        // exact, and bf16 / fp32 share the exponent range, so a probability far above 1 is harmless. Only if a block's
        // partial sum leaves the safe range (some p > ~2^64, or non-finite) is the reference raised: the warp waits
        // for the PV in flight, rescales its 32 rows of O in TMEM and its running sum, and redoes the block.
        constexpr float L_SAFE = 1.8446744e19f * 64.0f;  // 2^70
        float m_ref = -INFINITY, l_run = 0.f;
        for (int j = 0; j < it.nkb; ++j, ++n) {
          const int k0 = j * KBLK;
          const int nlive = min(KBLK, it.len - k0);
          const int nch = (nlive + 31) >> 5;
          const bool need_mask = (nlive & 31) != 0;
          const uint32_t ts = tmem + lane_addr + TM_S + (n & 1) * 64;
          uint8_t* prow = smem + SM_P + (n & 1) * P_BYTES + il * 128;
          mbar_wait(&bar_s[n & 1], (n >> 1) & 1);
          __syncwarp();
          tc_fence_after();
          if (warp_live) {
            float dummy = 0.f, l_blk = 0.f;
            uint32_t r0[32], r1[32];  // both chunks in flight before the single wait
            tmem_ld_32x32(ts, r0);
            if (nch == 2) tmem_ld_32x32(ts + 32, r1);
            tmem_wait_ld();
            float mu2;
            if (j == 0) {
              // exact row maximum of the first block; the biased / masked scores stay in the registers
              float m_blk = -INFINITY;
              if (nch == 2) {
                chunk<HAS_BIAS, false, false, true, HAS_BIAS>(r0, k0, it.len, gate, rel, 0.f, m_blk, l_blk, nullptr, 0);
                chunk<HAS_BIAS, true, false, true, true>(r1, k0 + 32, it.len, gate, rel, 0.f, m_blk, l_blk, nullptr, 0);
              } else {
                chunk<HAS_BIAS, true, false, true, true>(r0, k0, it.len, gate, rel, 0.f, m_blk, l_blk, nullptr, 0);
              }
              m_ref = m_blk;
              mu2 = ((m_ref == -INFINITY) ? 0.f : m_ref) * LOG2E;
              chunk<false, false, true, false>(r0, k0, it.len, 0.f, nullptr, mu2, dummy, l_blk, prow, 0);
              if (nch == 2) chunk<false, false, true, false>(r1, k0 + 32, it.len, 0.f, nullptr, mu2, dummy, l_blk, prow, 4);
            } else {
              mu2 = ((m_ref == -INFINITY) ? 0.f : m_ref) * LOG2E;
              if (nch == 2) {
                chunk<HAS_BIAS, false, true, false>(r0, k0, it.len, gate, rel, mu2, dummy, l_blk, prow, 0);
                if (need_mask)
                  chunk<HAS_BIAS, true, true, false>(r1, k0 + 32, it.len, gate, rel, mu2, dummy, l_blk, prow, 4);
                else
                  chunk<HAS_BIAS, false, true, false>(r1, k0 + 32, it.len, gate, rel, mu2, dummy, l_blk, prow, 4);
              } else {
                if (need_mask)
                  chunk<HAS_BIAS, true, true, false>(r0, k0, it.len, gate, rel, mu2, dummy, l_blk, prow, 0);
                else
                  chunk<HAS_BIAS, false, true, false>(r0, k0, it.len, gate, rel, mu2, dummy, l_blk, prow, 0);
              }
            }
            const bool unsafe = !(l_blk < L_SAFE);  // also true for NaN / inf
            if (j > 0 && __any_sync(0xffffffffu, unsafe)) {
              // rare: raise the reference to this block's exact maximum (rows that do not need it keep theirs)
              float m_blk = -INFINITY, l_dummy = 0.f;
              chunk<HAS_BIAS, true, false>(r0, k0, it.len, gate, rel, 0.f, m_blk, l_dummy, nullptr, 0);
              if (nch == 2) chunk<HAS_BIAS, true, false>(r1, k0 + 32, it.len, gate, rel, 0.f, m_blk, l_dummy, nullptr, 0);
              const float m_new = unsafe ? fmaxf(m_ref, m_blk) : m_ref;
              const float alpha = (m_new == m_ref) ? 1.f : ex2_approx((m_ref - m_new) * LOG2E);
              m_ref = m_new;
              mu2 = ((m_ref == -INFINITY) ? 0.f : m_ref) * LOG2E;
              l_run *= alpha;
              // O currently holds PV of blocks < n of this item; PV_{n-1} may still be in flight
              mbar_wait(&bar_o[(n - 1) & 1], ((n - 1) >> 1) & 1);
              __syncwarp();
              tc_fence_after();
              {
                uint32_t a0[32], a1[32];
                tmem_ld_32x32(tmem + lane_addr + TM_O, a0);
                tmem_ld_32x32(tmem + lane_addr + TM_O + 32, a1);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                  a0[k] = __float_as_uint(__uint_as_float(a0[k]) * alpha);
                  a1[k] = __float_as_uint(__uint_as_float(a1[k]) * alpha);
                }
                tmem_st_32x32(tmem + lane_addr + TM_O, a0);
                tmem_st_32x32(tmem + lane_addr + TM_O + 32, a1);
                tmem_wait_st();
              }
              l_blk = 0.f;
              chunk<HAS_BIAS, true, true, false>(r0, k0, it.len, gate, rel, mu2, dummy, l_blk, prow, 0);
              if (nch == 2) chunk<HAS_BIAS, true, true, false>(r1, k0 + 32, it.len, gate, rel, mu2, dummy, l_blk, prow, 4);
            }
            l_run += l_blk;
            fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_p[n & 1]);
        }
        // the one true round-trip wait per item: PV of the last block
        mbar_wait(&bar_o[(n - 1) & 1], ((n - 1) >> 1) & 1);
        __syncwarp();
        tc_fence_after();
        if (warp_live) {
          uint32_t a0[32], a1[32];
          tmem_ld_32x32(tmem + lane_addr + TM_O, a0);
          tmem_ld_32x32(tmem + lane_addr + TM_O + 32, a1);
          tmem_wait_ld();
          if (i < it.len) {
            const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
            uint4* dst = reinterpret_cast<uint4*>(a.out + ((long long)row0 + i) * a.D + it.h * HD);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t* src = q < 4 ? a0 : a1;
                const int c = (q & 3) * 8 + 2 * e;
                __nv_bfloat162 pk =
                    __floats2bfloat162_rn(__uint_as_float(src[c]) * inv, __uint_as_float(src[c + 1]) * inv);
                w[e] = *reinterpret_cast<uint32_t*>(&pk);
              }
              dst[q] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
        tc_fence_before();  // the TMEM reads above are ordered before the bar_p arrival that lets the next item's PV overwrite O
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
  CUtensorMap tmq, tmkv;
  if (make_tmap_2d(&tmq, a.qkv, 3ULL * a.D, (unsigned long long)a.B * a.slot, 3ULL * a.D, QT, err)) return -1;
  if (make_tmap_2d(&tmkv, a.qkv, 3ULL * a.D, (unsigned long long)a.B * a.slot, 3ULL * a.D, KBLK, err)) return -1;
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
    attention_tc_kernel<true><<<grid, 192, ATT_SMEM, st>>>(tmq, tmkv, a, (int)items);
  else
    attention_tc_kernel<false><<<grid, 192, ATT_SMEM, st>>>(tmq, tmkv, a, (int)items);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("attention_tc launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

}  // namespace ssr
