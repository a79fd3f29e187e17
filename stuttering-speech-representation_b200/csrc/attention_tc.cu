// Flash-style multi-head self-attention on the 5th-gen tensor cores (sm_100a), head_dim 64.
//
// Persistent, warp-specialised kernel; a work item is one (clip, head, 128-query tile); each item loops over
// 64-key blocks. Everything is double buffered so that the softmax warps never wait for a tensor-core round trip
// in steady state:
//     warp 4      TMA producer: Q per item, K_j / V_j tiles through a 4-stage ring (3 with the gated bias) that runs ahead,
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
//   Two CTAs are co-resident per SM (<= 113 KB smem, 256 TMEM columns each). Padding is not computed: the last key
//   block uses an MMA N / K extent rounded to 16 live keys, softmax touches only the 32-column chunks that hold live
//   keys, and warps whose 32 query rows are all beyond the clip's length only keep the barrier protocol going.
//
//   What a per-warp-role timeline of one CTA (clock64 at every barrier; build with -DSSR_ATT_TRACE, tools/attn_trace.py)
//   showed for the short WavLM items (3 key blocks), and what the kernel does about it:
//     * the gated bias, fetched per element with __ldg, doubled the time of a block (two L1 wavefronts per load, all
//       warps bursting at once) -> each warp stages the 95 table entries its 32 rows need per block in shared memory;
//     * the last PV of an item was queued behind the next item's Q load (single Q buffer: that load can only be
//       requested once this item's last S has retired, and takes ~2 us under load) -> it is issued at item end;
//     * clip length and gate prefetched into registers were spilled at once, i.e. waited for -> they travel
//       global -> shared by cp.async and are read one item later; the item walk needs no division;
//     * output rows are transposed through the idle P buffer so that a store instruction writes 4 whole 128-byte rows.
//   WavLM shape 167 -> 158 us per layer, Whisper shape 1318 -> 1262 us. What is left is a chain of fixed latencies per
//   block (mbarrier wait, tcgen05.ld, proxy fence, arrive: ~900 of ~2400 cycles) that only more resident CTAs per SM
//   could hide; 8 softmax warps per CTA (two per TMEM lane quadrant) and S running three blocks ahead of PV were both
//   built and measured: slower / no change.
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
constexpr int Q_BYTES = 128 * 64 * 2;   // [128 x 64] bf16
constexpr int KV_BYTES = 64 * 64 * 2;   // [64 x 64] bf16 (K or V of one block)
constexpr int P_BYTES = 128 * 64 * 2;   // [128 q x 64 keys] bf16
constexpr int SM_Q = 0;
constexpr int SM_KV = Q_BYTES;  // KV_STAGES x (K, V)
constexpr int WIN = 96;         // bias window per softmax warp and key block: 64 keys + 31 rows of skew (+1 pad)
// The gated-bias kernel trades one K/V stage for the per-warp relative-position windows (2 CTAs per SM either way).
template <bool HAS_BIAS>
struct Lay {
  static constexpr int KV_STAGES = HAS_BIAS ? 3 : 4;
  static constexpr int SM_P = SM_KV + KV_STAGES * 2 * KV_BYTES;  // 2 buffers
  static constexpr int SM_BAR = SM_P + 2 * P_BYTES;
  static constexpr int SM_WIN = SM_BAR + 256;
  static constexpr int SM_LEN = SM_BAR + 192;  // 6 x [2] int inside the barrier block: prefetched clip lengths
  static constexpr int SM_GATE = SM_WIN + (HAS_BIAS ? 4 * WIN * 4 : 0);  // [2][128] fp32: prefetched gates
  static constexpr int SMEM = SM_GATE + (HAS_BIAS ? 2 * 128 * 4 : 0);
  static_assert(2 * (SMEM + 1024) <= 228 * 1024, "two CTAs per SM");
};
constexpr int TMEM_COLS = 256;
constexpr int TM_S = 0, TM_O = 128;  // S[2] at columns 0 / 64, O at columns 128..191 (allocation is a power of two)
constexpr float LOG2E = 1.4426950408889634f;

#ifdef SSR_ATT_TRACE
__device__ long long g_att_trace[3][1024];
#define TR(role, ctr) do { if (blockIdx.x == 0 && (ctr) < 1024) g_att_trace[role][(ctr)++] = clock64(); } while (0)
#else
#define TR(role, ctr) do { } while (0)
#endif

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial, relative error 7.5e-5 — thirty times
// below the bf16 rounding P gets anyway): the MUFU unit issues 16 ex2 per clock per SM, which at head_dim 64 is twice
// the tensor-core time of a block, so a quarter of the exponentials of every row are computed here instead.
//   t = x + 1.5 * 2^23 rounds x to the nearest integer n in the low mantissa bits; r = x - n in [-0.5, 0.5];
//   2^x = p(r) * 2^n, the scaling done by adding n to the exponent field.
// x is clamped to [-125, 126]: below, the result is ~2^-125 (rounds to nothing next to the other terms, and -inf of a
// masked key lands here too); above, it is >= 2^125, which trips the L_SAFE check exactly like an overflowing MUFU.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fminf(fmaxf(x, -125.0f), 126.0f);
  const float t = x + 12582912.0f;
  const float r = x - (t - 12582912.0f);
  float p = fmaf(r, 0.05517164617776871f, 0.2426111251115799f);
  p = fmaf(p, r, 0.6932609677314758f);
  p = fmaf(p, r, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
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
// Decoding is split so that the clip-length fetch of item n+1 is issued one whole item before its first use
// (Cursor::load at the top of item n, finish_item at the top of item n+1): no role ever stalls on it.
struct ItemPre {
  int b, h, q0;
};
struct Step {
  int dq, db, dh;  // gridDim.x in the mixed radix of the item index (host-computed: lives in the constant bank)
  int paired;      // two query tiles per clip (129..256 frames): see Cursor
  int grouped;     // > 0: number of query tiles per clip, walked as the FASTEST digit of the item index: see Cursor
  int reverse;     // clips are visited from the last one: the qkv rows the producing GEMM wrote last are still in L2
};
// Walks the item list of one CTA (slot s = blockIdx.x + k * gridDim.x, k = 0, 1, ...) without a division per item.
//   default : s = (qt * B + b) * H + h  (query-tile-major: every CTA sees the same mix of full and partial tiles)
//   paired  : two query tiles per clip. s >> 1 = b * H + h and qt = (s & 1) ^ (k & 1) (gridDim.x is even): the two
//             tiles of a (clip, head) are worked on AT THE SAME TIME by neighbouring CTAs, so K and V come out of
//             HBM once and the second read hits L2 (query-tile-major order reads them twice, far apart: 471 MB
//             instead of 314 MB per WavLM-Large layer at B = 256), and every CTA still alternates between full and
//             tail tiles from round to round.
//   grouped : three or more query tiles per clip (Whisper: 12). s = (b * H + h) * NQT + qt: the NQT tiles of a
//             (clip, head) are worked on at the same time by NQT neighbouring CTAs, so its K and V leave HBM once and
//             the other NQT - 1 reads hit L2. In query-tile-major order the 296 resident CTAs sweep the whole batch
//             (737 MB of qkv at B = 64, six times L2) once per query tile and K / V are read from HBM NQT times:
//             ncu measured 6.0 GB per launch against 0.98 GB of algorithmic traffic, 78 % of HBM bandwidth.
struct Cursor {
  int qt, b, h;  // current item
  __device__ __forceinline__ void init(const AttentionArgs& a, const Step& st) {
    if (st.paired) {
      const int p = (int)blockIdx.x >> 1;
      b = p / a.H;
      h = p - b * a.H;
      qt = (int)blockIdx.x & 1;
      return;
    }
    if (st.grouped) {
      const int p = (int)blockIdx.x / st.grouped;
      qt = (int)blockIdx.x - p * st.grouped;
      b = p / a.H;
      h = p - b * a.H;
      return;
    }
    const int bh = a.B * a.H;
    qt = (int)blockIdx.x / bh;
    const int rem = (int)blockIdx.x - qt * bh;
    b = rem / a.H;
    h = rem - b * a.H;
  }
  __device__ __forceinline__ void advance(const AttentionArgs& a, const Step& st) {
    if (st.grouped) {  // digits (fastest first): qt in [0, NQT), h in [0, H), b
      qt += st.dq;
      if (qt >= st.grouped) {
        qt -= st.grouped;
        ++h;
      }
      h += st.dh;
      if (h >= a.H) {
        h -= a.H;
        ++b;
      }
      b += st.db;
      return;
    }
    h += st.dh;
    b += st.db;
    qt += st.dq;
    if (h >= a.H) {
      h -= a.H;
      ++b;
    }
    if (st.paired) {
      qt ^= 1;
    } else if (b >= a.B) {
      b -= a.B;
      ++qt;
    }
  }
  // The clip length of the item goes global -> shared memory asynchronously (consumed one item later): a value
  // prefetched into a register would be spilled by the compiler at once, i.e. waited for.
  __device__ __forceinline__ ItemPre load(const AttentionArgs& a, const Step& st, int* len_slot, bool copy = true) const {
    ItemPre p;
    p.b = st.reverse ? a.B - 1 - b : b;
    p.h = h;
    p.q0 = qt * QT;
    if (copy)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(len_slot)), "l"(a.lens + p.b) : "memory");
    return p;
  }
};
// Host side of the item walk: gridDim.x decomposed in the mixed radix of the chosen order (tests/test_host_cpu.py
// compiles Step, Cursor and this function for the host and checks that every item is visited exactly once).
static Step make_step(const AttentionArgs& a, int grid, int paired, int grouped, int reverse) {
  const int nqt = (a.slot + QT - 1) / QT;
  Step step;
  step.dq = grid / (a.B * a.H);
  step.db = (grid % (a.B * a.H)) / a.H;
  step.dh = grid % a.H;
  step.paired = 0;
  step.grouped = 0;
  step.reverse = reverse;
  if (paired && nqt == 2 && (grid & 1) == 0) {
    step.paired = 1;
    step.dq = 0;
    step.db = (grid / 2) / a.H;
    step.dh = (grid / 2) % a.H;
  } else if (grouped && nqt >= 3) {
    step.grouped = nqt;
    step.dq = grid % nqt;
    step.dh = (grid / nqt) % a.H;
    step.db = grid / (nqt * a.H);
  }
  return step;
}
__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ Item finish_item(const AttentionArgs& a, const ItemPre& p, int len_raw) {
  Item it;
  it.b = p.b;
  it.h = p.h;
  it.q0 = p.q0;
  it.len = min(len_raw, a.slot);
  it.valid = it.q0 < it.len;  // tiles past the clip's live frames are never consumed (engine.cu: slot layout)
  it.nkb = (it.len + KBLK - 1) / KBLK;
  return it;
}


// One 32-key chunk of one query row: scores (+ gated relative-position bias, + key mask) -> running block max and,
// if WRITE_P, probabilities exp(s - ref) accumulated into l_blk and stored as bf16 into the row's P tile columns
// col0 .. col0+3 (16-byte units, XOR-swizzled by row % 8 = the UMMA / TMA 128B swizzle).
// WRITE_BACK keeps the biased / masked scores in `raw`, so that a second pass over the same registers needs neither
// the bias table nor the mask again.
// Packed fp32 pairs (sm_100 FFMA2 / FADD2: one issue slot for two lanes of work). The softmax warps are bound by
// instruction issue and dependent-issue latency (two warps per scheduler), not by any one pipe, so every instruction
// taken out of the per-element sequence shortens the block.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <bool HAS_BIAS, bool MASK, bool WRITE_P, bool TRACK_MAX = true, bool WRITE_BACK = false, bool POLY = false,
          bool PACK2 = false>
__device__ __forceinline__ void chunk(uint32_t (&raw)[32], int jg0, int len, float gate, const float* rel, float mu2,
                                      float& m_blk, float& l_blk, uint8_t* prow, uint32_t col0) {
  uint32_t packed[16];
  uint64_t l2 = 0;  // (0.f, 0.f): PACK2 keeps the row sum as a pair of partial sums
  const uint64_t scale2 = pack2(LOG2E, LOG2E), shift2 = pack2(-mu2, -mu2), gate2 = pack2(gate, gate);
#pragma unroll
  for (int k = 0; k < 32; k += 2) {
    float v0 = __uint_as_float(raw[k]), v1 = __uint_as_float(raw[k + 1]);
    if (HAS_BIAS) {
      // rel: this row's view of the warp's shared-memory window (see WIN); its alignment depends on the lane
      if (PACK2) {
        unpack2(fma2(gate2, pack2(rel[k], rel[k + 1]), pack2(v0, v1)), v0, v1);
      } else {
        v0 = fmaf(gate, rel[k], v0);
        v1 = fmaf(gate, rel[k + 1], v1);
      }
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
      // POLY: elements 0, 1 of every 8 (a quarter of the row) take the polynomial instead of the MUFU unit
      const bool poly = POLY && (k & 6) == 0;
      float x0, x1;
      if (PACK2) {
        unpack2(fma2(pack2(v0, v1), scale2, shift2), x0, x1);
      } else {
        x0 = fmaf(v0, LOG2E, -mu2);
        x1 = fmaf(v1, LOG2E, -mu2);
      }
      const float p0 = poly ? ex2_poly(x0) : ex2_approx(x0);
      const float p1 = poly ? ex2_poly(x1) : ex2_approx(x1);
      if (PACK2)
        l2 = add2(l2, pack2(p0, p1));
      else
        l_blk += p0 + p1;
      __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
      packed[k >> 1] = *reinterpret_cast<uint32_t*>(&pk);
    }
  }
  if (WRITE_P) {
    if (PACK2) {
      float la, lb;
      unpack2(l2, la, lb);
      l_blk += la + lb;
    }
    const uint32_t sw = (uint32_t)((reinterpret_cast<uintptr_t>(prow) >> 7) & 7);  // row % 8 (rows are 128 B apart)
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(prow + ((col0 + q) ^ sw) * 16) =
          make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
  }
}

// VAR bit 0: packed fp32 pair arithmetic in the softmax (FFMA2 / FADD2); bit 1: a quarter of the exponentials on the
// FMA pipe (ex2_poly). Measured on B200 (tools/attn_probe.py): packed pairs are worth 1-3 % (WavLM
// shape 146 -> 142 us per layer, Whisper shape 1.30 -> 1.28 ms); the polynomial share makes both shapes SLOWER
// (Whisper 1.30 -> 1.37 ms): in their exponential phase the softmax warps do queue on the MUFU unit (ncu source page),
// but nine FMA / ALU slots per polynomial exponential cost as much issue time as the eight MUFU cycles they free.
template <bool HAS_BIAS, int VAR>
__global__ void __launch_bounds__(192, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmkv,
                    const AttentionArgs a, const int n_items, const Step step) {
  constexpr bool PACK2 = (VAR & 1) != 0, POLY = (VAR & 2) != 0;
  constexpr int KV_STAGES = Lay<HAS_BIAS>::KV_STAGES;
  constexpr int SM_P = Lay<HAS_BIAS>::SM_P;
  constexpr int SM_BAR = Lay<HAS_BIAS>::SM_BAR;
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
  if (HAS_BIAS && threadIdx.x < 128) {
    // rows beyond the clip slot never get a gate copied: they must read a finite value, not stale shared memory
    float* g0 = reinterpret_cast<float*>(smem + Lay<HAS_BIAS>::SM_GATE);
    g0[threadIdx.x] = 0.f;
    g0[128 + threadIdx.x] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  griddep_wait();  // programmatic dependent launch: qkv / gate / lens come from earlier kernels
  griddep_launch();

  if (threadIdx.x == 128) {
    // ============================ TMA producer ============================
    uint32_t n_item = 0, n = 0;
    int trc = 0; (void)trc;
    Cursor cur;
    cur.init(a, step);
    int* lsm = reinterpret_cast<int*>(smem + Lay<HAS_BIAS>::SM_LEN) + 8;  // [2] private to this thread
    uint32_t lbuf = 0;
    ItemPre nxt = cur.load(a, step, lsm);
    cp_async_commit();
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, lbuf ^= 1) {
      cp_async_wait_all();
      const Item it = finish_item(a, nxt, lsm[lbuf]);
      if (idx + (int)gridDim.x < n_items) {  // prefetch the next item's length
        cur.advance(a, step);
        nxt = cur.load(a, step, lsm + (lbuf ^ 1));
        cp_async_commit();
      }
      if (!it.valid) continue;
      const int row0 = it.b * a.slot;
      mbar_wait(q_empty, (n_item & 1) ^ 1);
      TR(0, trc);
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_2d(smem + SM_Q, &tmq, q_full, it.h * HD, row0 + it.q0);
      for (int j = 0; j < it.nkb; ++j, ++n) {
        const int s = n % KV_STAGES;
        mbar_wait(&kv_empty[s], ((n / KV_STAGES) & 1) ^ 1);
        TR(0, trc);
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
    int trc = 0; (void)trc;
    // PV of block n-1 is issued after S of block n and accumulates into O over the whole item (the softmax warps
    // rescale O in place on the rare occasion the softmax reference changes).
    auto issue_pv_prev = [&]() {
      const uint32_t pn = n - 1;
      const int ps = pn % KV_STAGES;
      mbar_wait(&bar_p[pn & 1], (pn >> 1) & 1);  // P_{n-1} is in shared memory, O is ready to take PV_{n-1}
      TR(1, trc);
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
    Cursor cur;
    cur.init(a, step);
    int* lsm = reinterpret_cast<int*>(smem + Lay<HAS_BIAS>::SM_LEN) + 10;  // [2] private to this thread
    uint32_t lbuf = 0;
    ItemPre nxt = cur.load(a, step, lsm);
    cp_async_commit();
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, lbuf ^= 1) {
      cp_async_wait_all();
      const Item it = finish_item(a, nxt, lsm[lbuf]);
      if (idx + (int)gridDim.x < n_items) {  // prefetch the next item's length
        cur.advance(a, step);
        nxt = cur.load(a, step, lsm + (lbuf ^ 1));
        cp_async_commit();
      }
      if (!it.valid) continue;
      mbar_wait(q_full, n_item & 1);
      TR(1, trc);
      for (int j = 0; j < it.nkb; ++j) {
        const int s = n % KV_STAGES;
        const int nlive = min(KBLK, it.len - j * KBLK);
        const int n16 = (nlive + 15) >> 4;  // live keys in units of 16
        mbar_wait(&kv_full[s], (n / KV_STAGES) & 1);
        TR(1, trc);
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
      // The last PV of the item is not held back behind the next item's Q load (which can only be requested once
      // this item's last S has retired): O reaches the softmax warps one TMA latency earlier.
      issue_pv_prev();
      have_prev = false;
      ++n_item;
    }
  } else if (warp < 4) {
    // ============================ softmax / output warps ============================
    const uint32_t quad = warp;  // TMEM lane quadrant accessible to this warp
    const int il = quad * 32 + lane;
    const uint32_t lane_addr = (quad * 32u) << 16;
    uint32_t n = 0;
    int trc = 0; (void)trc;
    Cursor cur;
    cur.init(a, step);
    int* lsm = reinterpret_cast<int*>(smem + Lay<HAS_BIAS>::SM_LEN) + quad * 2;  // [2] per warp, copied by lane 0
    ItemPre nxt = cur.load(a, step, lsm, lane == 0);
    // The row's gate of the NEXT item travels global -> shared memory by cp.async (no register is live across the
    // item, so nothing makes the warp wait for the load) and is read one item later; each thread reads its own word.
    float* gsm = reinterpret_cast<float*>(smem + Lay<HAS_BIAS>::SM_GATE) + il;
    uint32_t gbuf = 0;
    auto prefetch_gate = [&](const ItemPre& p, uint32_t buf) {
      if (HAS_BIAS && p.q0 + il < a.slot)
        cp_async_f32(gsm + buf * 128, a.gate + ((long long)p.b * a.slot + p.q0 + il) * a.H + p.h);
      cp_async_commit();  // closes the group that also holds the length copy
    };
    prefetch_gate(nxt, 0);
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, gbuf ^= 1) {
      if (threadIdx.x == 0) TR(2, trc);
      cp_async_wait_all();
      __syncwarp();  // lane 0's length copy is visible to the warp
      const Item it = finish_item(a, nxt, lsm[gbuf]);
      float gate = 0.f;
      if (HAS_BIAS) gate = gsm[gbuf * 128];
      // prefetch the next item's length and this row's gate: consumed one iteration later
      if (idx + (int)gridDim.x < n_items) {
        cur.advance(a, step);
        nxt = cur.load(a, step, lsm + (gbuf ^ 1), lane == 0);
        prefetch_gate(nxt, gbuf ^ 1);
      }
      if (!it.valid) continue;
      if (threadIdx.x == 0) TR(2, trc);
      const int row0 = it.b * a.slot;
      const bool warp_live = it.q0 + (int)quad * 32 < it.len;   // at least one live query row in this warp
      // Gated bias: row i needs table[h][j - i] for the 64 keys j of a block. The 32 rows of a warp together touch
      // only 95 consecutive table entries per block, so the warp stages that window in shared memory (3 coalesced
      // loads per lane, issued before the wait for S) and every row reads its 64 values at a lane-dependent skew:
      // win[x] = table[h][64 j_blk + x - 31 - i_lane0]  =>  row (lane l), key k of the block: win[k + 31 - l].
      // (Per-element global loads cost two L1 wavefronts each and made the bias blocks twice as long as plain ones.)
      float* win = reinterpret_cast<float*>(smem + Lay<HAS_BIAS>::SM_WIN) + quad * WIN;
      const float* rel = win + 31 - lane;
      const float* table = nullptr;
      int win_base = 0;
      if (HAS_BIAS) {
        table = a.relbias + (long long)it.h * a.rel_stride;
        win_base = a.rel_center - 31 - (it.q0 + (int)quad * 32);
      }
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
          float w0 = 0.f, w1 = 0.f, w2 = 0.f;
          if (HAS_BIAS && warp_live) {
            // entries outside the table belong to padded rows / masked keys only: clamp the index, never use the value
            const int t0 = win_base + k0 + (int)lane, hi = a.rel_stride - 1;
            w0 = __ldg(table + min(max(t0, 0), hi));
            w1 = __ldg(table + min(max(t0 + 32, 0), hi));
            w2 = __ldg(table + min(max(t0 + 64, 0), hi));
          }
          mbar_wait(&bar_s[n & 1], (n >> 1) & 1);
          if (threadIdx.x == 0) TR(2, trc);
          __syncwarp();  // also: every lane is done reading the previous block's window
          tc_fence_after();
          if (warp_live) {
            if (HAS_BIAS) {
              win[lane] = w0;
              win[lane + 32] = w1;
              win[lane + 64] = w2;
              __syncwarp();
            }
            float dummy = 0.f, l_blk = 0.f;
            uint32_t r0[32], r1[32];  // both chunks in flight before the single wait
            tmem_ld_32x32(ts, r0);
            if (nch == 2) tmem_ld_32x32(ts + 32, r1);
            tmem_wait_ld();
            if (threadIdx.x == 0) TR(2, trc);
            float mu2;
            if (j == 0) {
              // exact row maximum of the first block; the biased / masked scores stay in the registers
              float m_blk = -INFINITY;
              if (nch == 2) {
                chunk<HAS_BIAS, false, false, true, HAS_BIAS, false, PACK2>(r0, k0, it.len, gate, rel, 0.f, m_blk, l_blk, nullptr, 0);
                chunk<HAS_BIAS, true, false, true, true, false, PACK2>(r1, k0 + 32, it.len, gate, rel + 32, 0.f, m_blk, l_blk, nullptr, 0);
              } else {
                chunk<HAS_BIAS, true, false, true, true, false, PACK2>(r0, k0, it.len, gate, rel, 0.f, m_blk, l_blk, nullptr, 0);
              }
              m_ref = m_blk;
              mu2 = ((m_ref == -INFINITY) ? 0.f : m_ref) * LOG2E;
              chunk<false, false, true, false, false, POLY, PACK2>(r0, k0, it.len, 0.f, nullptr, mu2, dummy, l_blk, prow, 0);
              if (nch == 2)
                chunk<false, false, true, false, false, POLY, PACK2>(r1, k0 + 32, it.len, 0.f, nullptr, mu2, dummy, l_blk, prow, 4);
            } else {
              mu2 = ((m_ref == -INFINITY) ? 0.f : m_ref) * LOG2E;
              if (nch == 2) {
                chunk<HAS_BIAS, false, true, false, false, POLY, PACK2>(r0, k0, it.len, gate, rel, mu2, dummy, l_blk, prow, 0);
                if (need_mask)
                  chunk<HAS_BIAS, true, true, false, false, POLY, PACK2>(r1, k0 + 32, it.len, gate, rel + 32, mu2, dummy, l_blk, prow, 4);
                else
                  chunk<HAS_BIAS, false, true, false, false, POLY, PACK2>(r1, k0 + 32, it.len, gate, rel + 32, mu2, dummy, l_blk, prow, 4);
              } else {
                if (need_mask)
                  chunk<HAS_BIAS, true, true, false, false, POLY, PACK2>(r0, k0, it.len, gate, rel, mu2, dummy, l_blk, prow, 0);
                else
                  chunk<HAS_BIAS, false, true, false, false, POLY, PACK2>(r0, k0, it.len, gate, rel, mu2, dummy, l_blk, prow, 0);
              }
            }
            const bool unsafe = !(l_blk < L_SAFE);  // also true for NaN / inf
            if (j > 0 && __any_sync(0xffffffffu, unsafe)) {
              // rare: raise the reference to this block's exact maximum (rows that do not need it keep theirs)
              float m_blk = -INFINITY, l_dummy = 0.f;
              chunk<HAS_BIAS, true, false>(r0, k0, it.len, gate, rel, 0.f, m_blk, l_dummy, nullptr, 0);
              if (nch == 2) chunk<HAS_BIAS, true, false>(r1, k0 + 32, it.len, gate, rel + 32, 0.f, m_blk, l_dummy, nullptr, 0);
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
              if (nch == 2) chunk<HAS_BIAS, true, true, false>(r1, k0 + 32, it.len, gate, rel + 32, mu2, dummy, l_blk, prow, 4);
            }
            l_run += l_blk;
            if (threadIdx.x == 0) TR(2, trc);
            fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            if (threadIdx.x == 0) TR(2, trc);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_p[n & 1]);
          if (threadIdx.x == 0) TR(2, trc);
        }
        // the one true round-trip wait per item: PV of the last block
        mbar_wait(&bar_o[(n - 1) & 1], ((n - 1) >> 1) & 1);
        if (threadIdx.x == 0) TR(2, trc);
        __syncwarp();
        tc_fence_after();
        if (warp_live) {
          uint32_t a0[32], a1[32];
          tmem_ld_32x32(tmem + lane_addr + TM_O, a0);
          tmem_ld_32x32(tmem + lane_addr + TM_O + 32, a1);
          tmem_wait_ld();
          if (threadIdx.x == 0) TR(2, trc);
          // A row of the output is 128 contiguous bytes: transpose through this warp's 32 rows of the P buffer the
          // last PV has just finished reading (same XOR swizzle => conflict-free), so that each store instruction
          // writes 4 whole rows instead of 32 scattered 16-byte pieces.
          const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
          uint8_t* stage = smem + SM_P + ((n - 1) & 1) * P_BYTES + quad * 32 * 128;
          {
            uint8_t* mine = stage + lane * 128;
            const uint32_t sw = lane & 7;
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
              *reinterpret_cast<uint4*>(mine + ((q ^ sw) * 16)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
          __syncwarp();
          if (threadIdx.x == 0) TR(2, trc);
          {
            const uint32_t rsub = lane >> 3, cch = lane & 7;
            bf16* dst = a.out + ((long long)row0 + it.q0 + quad * 32) * a.D + it.h * HD + cch * 8;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const uint32_t r = t * 4 + rsub;
              const uint4 v = *reinterpret_cast<const uint4*>(stage + r * 128 + ((cch ^ (r & 7)) * 16));
              if (it.q0 + (int)(quad * 32 + r) < it.len) *reinterpret_cast<uint4*>(dst + (long long)r * a.D) = v;
            }
          }
          __syncwarp();  // the staging rows are rewritten by this warp's next P tile
        }
        if (threadIdx.x == 0) TR(2, trc);
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

// Kernel variant (see attention_tc_kernel: bit 0 = packed pair arithmetic, bit 1 = polynomial exp2 share);
// ssr_tuning_set.
int g_attention_variant = 1;
// 1: two-tile clips use the paired item order (see Cursor); 0: query-tile-major order everywhere.
int g_attention_paired = 1;
// 1: clips are visited from the last one (the qkv rows the QKV GEMM wrote last are still in L2).
int g_attention_reverse = 1;
// 1: clips of three or more query tiles use the grouped item order (see Cursor); 0: query-tile-major order.
int g_attention_grouped = 1;

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
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const AttentionArgs, const int, const Step);
  static const KernelFn kern[2][4] = {
      {attention_tc_kernel<false, 0>, attention_tc_kernel<false, 1>, attention_tc_kernel<false, 2>,
       attention_tc_kernel<false, 3>},
      {attention_tc_kernel<true, 0>, attention_tc_kernel<true, 1>, attention_tc_kernel<true, 2>,
       attention_tc_kernel<true, 3>}};
  if (!attr_set) {
    for (int bsel = 0; bsel < 2; ++bsel)
      for (int v = 0; v < 4; ++v) {
        cudaError_t c1 = cudaFuncSetAttribute(kern[bsel][v], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              bsel ? Lay<true>::SMEM : Lay<false>::SMEM);
        if (c1 != cudaSuccess) {
          err = std::string("cudaFuncSetAttribute(attention_tc_kernel): ") + cudaGetErrorString(c1);
          return -1;
        }
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
  const Step step = make_step(a, grid, g_attention_paired, g_attention_grouped, g_attention_reverse);
  const int bsel = a.gate != nullptr ? 1 : 0;
  const int var = g_attention_variant & 3;
  launch_pdl(kern[bsel][var], dim3(grid), dim3(192), bsel ? Lay<true>::SMEM : Lay<false>::SMEM, st, tmq, tmkv, a,
             (int)items, step);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    err = std::string("attention_tc launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 0;
}

}  // namespace ssr

#ifdef SSR_ATT_TRACE
extern "C" int ssr_att_trace_fetch(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, ssr::g_att_trace, sizeof(long long) * 3 * 1024);
}
#endif
