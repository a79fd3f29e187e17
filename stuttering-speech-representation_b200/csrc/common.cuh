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
// erf-GELU:  gelu(x) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt(2)),  0.5 |x| erfc(z) = Q(t) exp(-z^2),  t = 1 / (1 + c z).
// The Abramowitz-Stegun 7.1.25/26 family writes erfc(z) = t P(t) exp(-z^2); since |x| = sqrt(2) (1 - t) / (c t), the
// factor 0.5 |x| t folds into the polynomial: Q(t) = (1 - t) R(t), a cubic, fitted directly (minimax in the absolute
// error of gelu over |x| <= 9, c chosen by scan) under R(1) = 1 / (sqrt(2) c), which makes value and slope at x = 0
// exact: gelu(0) = 0 bit for bit (the fp32 Horner sum of the coefficients is exactly 0 — an all-zero clip must stay
// all-zero through LayerNorm stacks), |abs err| <= 9.2e-6 in fp32 arithmetic (50 times below the tanh form's
// 4.7e-4), relative error 1.7e-4 at |x| = 1e-3. Ten instructions: FFMA, MUFU.RCP, 3 FFMA, 2
// FMUL, MUFU.EX2, FMNMX, FFMA (the 5-term 7.1.26 form took 13; on the power-limited FFN1 epilogue every FP32
// instruction per output element costs about 0.5 ms per WavLM-Large step).
__device__ __forceinline__ float gelu_fast(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3379970414071697f, ax, 1.0f)));
  float q = fmaf(-1.033210277557373f, t, 1.0818655490875244f);
  q = fmaf(q, t, -0.5434032082557678f);
  q = fmaf(q, t, 0.49474793672561646f);
  float ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"((ax * (-0.5f * 1.4426950408889634f)) * ax));  // exp(-x^2 / 2)
  return fmaf(-q, ex, fmaxf(x, 0.0f));
}

#endif

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// Process-wide switch (ssr_tuning_set "pdl"): 1 = the per-layer kernels (LayerNorm, CTA-pair GEMM, attention) are
// launched with programmatic stream serialization, so that each one's prologue (barrier init, TMEM allocation,
// tensor-map prefetch, CTA scheduling) overlaps the tail of its predecessor instead of following it.
// Default 0: measured on B200 the step gets SLOWER with it (WavLM-Large 33.5 -> 34.3 ms, Whisper-large 162 -> 165 ms,
// two alternating runs each): successor CTAs that become resident early sit in griddepcontrol.wait holding warp
// slots, registers and TMEM that the running kernel's own CTAs then lack.
extern int g_pdl;

// Launch `kernel` (which calls ptx::griddep_wait before its first dependent global access) with the programmatic
// dependent launch attribute. Compile-time __cluster_dims__ of the kernel are honoured by cudaLaunchKernelEx.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace ssr
