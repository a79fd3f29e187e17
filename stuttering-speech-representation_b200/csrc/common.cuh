// Shared host/device declarations for the ssr_b200 kernels (internal; the public C ABI is include/ssr_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>

namespace ssr {

typedef __nv_bfloat16 bf16;

enum Act { ACT_NONE = 0, ACT_GELU = 1 };

// Epilogue applied to every GEMM output element  acc[r, c]  (r = flat A row, c = output column):
//   v = act(acc + bias[c]) + resid[...]
// and the row is routed through a slot remap, which is how padded / strided per-clip layouts are compacted:
//   b = r / in_slot, t = r % in_slot;  row is live iff t < (lens ? lens[b] : valid)
//   out_row = b * out_slot + t + out_off
// resid is indexed by out_row (resid_by_t == 0) or by t (resid_by_t == 1, e.g. a positional table).
struct EpiParams {
  const float* bias;
  const float* resid;
  float* out_f32;
  bf16* out_bf16;
  const int* lens;
  int ldr, ldo32, ldo16;
  int act;
  int in_slot, out_slot, valid, out_off;  // in_slot == 0: identity mapping, all rows < M live
  int resid_by_t;
  // fused time mean-pool: per 32-row group and segment (0: clip of the group's first row, 1: the next clip)
  // column sums are written to pool_part[(group * 2 + seg) * N + c]; a finalize kernel reduces them in fixed order.
  float* pool_part;
  int pool_slot;  // rows per clip in OUTPUT row space (live rows only are summed)
};

struct GemmParams {
  int M, N, K;
  int a_mode;  // 0: A[r, k] plain (tensor map {K, M});  1: positional conv, A[r, tap*64 + c] = X[r + tap, ntile*64 + c]
  int num_m_tiles, num_n_tiles, num_kb;
  EpiParams epi;
};

struct GemmOp {
  const bf16* A;
  long long lda;  // elements between consecutive A rows (may be < K: overlapping conv windows)
  long long a_rows;  // rows addressable in A (tensor-map bound; rows >= a_rows read as zero)
  const bf16* W;  // [N, K] row-major
  int M, N, K;
  int a_mode;
  int a_cols;  // a_mode 1: width (channels) of X
  EpiParams epi;
};

// WavLM positional convolution on the zero-padded, group-padded signal X [B * pslot (+ tail), 1024] (bf16) with the
// block-diagonal weight W [1024, 128 taps * 64] (bf16). Output row r (flat over B * pslot) is routed by `epi`.
struct PosConvOp {
  const bf16* X;
  long long x_rows;   // addressable rows of X (rows beyond read as zero)
  const bf16* W;
  int B, pslot;       // rows per clip in X
  int rows_per_clip;  // output rows needed per clip (<= pslot - 128)
  EpiParams epi;
};

// ---- launchers (each returns cudaError_t / sets message in err) ----
int launch_gemm(const GemmOp& op, cudaStream_t stream, bool simt, int num_sms, std::string& err);
int launch_posconv(const PosConvOp& op, cudaStream_t stream, int num_sms, std::string& err);

// 2-D bf16 tensor map: dim0 (contiguous) x dim1 rows of pitch `pitch_elems`; box = 64 x box_rows; 128-byte swizzle.
int make_tmap_2d(CUtensorMap* m, const void* base, unsigned long long dim0, unsigned long long dim1,
                 unsigned long long pitch_elems, unsigned box_rows, std::string& err);

#ifdef __CUDACC__
// erf-GELU with erfc(|z|) from Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the bf16 / fp32-residual
// noise floor of the consumers) : 2 MUFU ops + ~10 FMA instead of libdevice erff's two-branch polynomial.
//   gelu(x) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt(2))
// Constants are pre-folded (1/sqrt(2) into p and into the exponent scale, 1/2 into the polynomial) so that the whole
// function is 13 instructions: FFMA, MUFU.RCP, 4 FFMA, 4 FMUL, MUFU.EX2, FMNMX, FFMA.
__device__ __forceinline__ float gelu_fast(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.0f)));
  float pl = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  pl = fmaf(pl, t, 0.5f * 1.421413741f);
  pl = fmaf(pl, t, 0.5f * -0.284496736f);
  pl = fmaf(pl, t, 0.5f * 0.254829592f);
  float ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"((ax * -0.5f * 1.4426950408889634f) * ax));  // exp(-x^2/2)
  const float w = (pl * t) * ax;                // 0.5 |x| erfc(z) / exp(-z^2)
  return fmaf(-w, ex, fmaxf(x, 0.0f));
}

#endif

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace ssr
