// Downstream classifier head on the pooled embeddings, trained data-parallel (BASELINE configs[3]).
//
// NEW component: the reference trains only sklearn estimators inside `Pipeline([StandardScaler, classifier])` with
// `class_weight='balanced'` (/root/reference/model_training_1.py:576-589, :630-680). What is kept from it: the
// StandardScaler semantics (population variance, zero-variance features scale 1), the balanced class weights
// (n / (n_classes * count_c), sklearn compute_class_weight) and the evaluation metrics. The classifier itself is a
// one-hidden-layer MLP (D -> H -> C, ReLU) under class-weighted softmax cross-entropy and Adam.
//
// Everything here is stateless: the caller owns parameters, gradients, optimizer moments and scratch as flat device
// buffers (so that the gradient buffer can be handed to NCCL as is). Flat parameter layout, float32:
//     W1 [H, D] | b1 [H] | W2 [C, H] | b2 [C]                       P = H*D + H + C*H + C
// Gradient buffer: P + 2 floats; the two extra slots carry  sum_i w_i * loss_i  and  sum_i w_i  of the local rows, so
// ONE all-reduce(sum) of the whole buffer yields the global unnormalised gradient and its normaliser; `ssr_head_adam`
// divides by the reduced weight sum on the device (no host round trip inside a step).
//
// The head is small (a few GFLOP per step); the kernels are fp32 CUDA-core GEMMs with fixed reduction order: the
// 1-GPU and N-GPU runs differ only by the all-reduce's summation order, and a run is bit-reproducible.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ssr_b200.h"

namespace ssr {

constexpr int TM = 64, TN = 64, TK = 16;

struct Operand {
  const float* p;
  long long ld;        // stride between consecutive "major" indices
  const int* gather;   // optional row gather (applies to the sample dimension of X)
  const float* mean;   // optional fused StandardScaler: (x - mean[f]) * inv_std[f], f = feature index
  const float* inv_std;
};

// A(i, k): A_K ? p[row(i)*ld + k] : p[row(k)*ld + i]     (gather applies to the index that multiplies ld)
// B(k, j): B_K ? p[row(j)*ld + k] : p[row(k)*ld + j]
// SCALE_A / SCALE_B: the scaler's feature index is the contiguous index of that operand.
template <bool A_K, bool B_K, bool SCALE_A, bool SCALE_B, bool BIAS_RELU>
__global__ void __launch_bounds__(256) head_sgemm_kernel(Operand A, Operand B, float* __restrict__ Cm, long long ldc,
                                                         int M, int N, int K, const float* __restrict__ bias) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    // ---- load A tile (TM x TK) and B tile (TK x TN): 1024 elements each, 4 per thread
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      {
        int ii, kk;
        if (A_K) { kk = idx % TK; ii = idx / TK; } else { ii = idx % TM; kk = idx / TM; }
        const int gi = i0 + ii, gk = k0 + kk;
        float v = 0.f;
        if (gi < M && gk < K) {
          const int major = A_K ? gi : gk, minor = A_K ? gk : gi;
          const long long r = A.gather ? A.gather[major] : major;
          v = A.p[r * A.ld + minor];
          if (SCALE_A) v = (v - A.mean[minor]) * A.inv_std[minor];
        }
        As[kk][ii] = v;
      }
      {
        int jj, kk;
        if (B_K) { kk = idx % TK; jj = idx / TK; } else { jj = idx % TN; kk = idx / TN; }
        const int gj = j0 + jj, gk = k0 + kk;
        float v = 0.f;
        if (gj < N && gk < K) {
          const int major = B_K ? gj : gk, minor = B_K ? gk : gj;
          const long long r = B.gather ? B.gather[major] : major;
          v = B.p[r * B.ld + minor];
          if (SCALE_B) v = (v - B.mean[minor]) * B.inv_std[minor];
        }
        Bs[kk][jj] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int gi = i0 + ty * 4 + r;
    if (gi >= M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gj = j0 + tx * 4 + c;
      if (gj >= N) continue;
      float v = acc[r][c];
      if (BIAS_RELU) v = fmaxf(v + bias[gj], 0.f);
      Cm[(long long)gi * ldc + gj] = v;
    }
  }
}

// Output layer, loss and the back-propagated hidden gradient, one warp per row.
//   logits = A1[row] . W2^T + b2 ; p = softmax(logits)
//   training (y != null): R[row] = (w*loss, w), dZ2[row, c] = w * (p_c - [c == y]),
//                         dZ1[row, h] = (sum_c dZ2[row, c] * W2[c, h]) * [A1[row, h] > 0]
//   inference (y == null): pred[row] = argmax, proba[row, :] = p
constexpr int MAX_C = 32;
__global__ void __launch_bounds__(256) head_out_kernel(const float* __restrict__ A1, int n, int H, int C,
                                                       const float* __restrict__ W2, const float* __restrict__ b2,
                                                       const int* __restrict__ y, const int* __restrict__ gather,
                                                       const float* __restrict__ class_w, float* __restrict__ dZ2,
                                                       float* __restrict__ dZ1, float* __restrict__ R,
                                                       int* __restrict__ pred, float* __restrict__ proba) {
  extern __shared__ float w2s[];  // [C][H]
  for (int i = threadIdx.x; i < C * H; i += blockDim.x) w2s[i] = W2[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row = blockIdx.x * 8 + warp; row < n; row += gridDim.x * 8) {
    const float* a = A1 + (long long)row * H;
    float logit = 0.f;  // lane c keeps class c
    for (int c = 0; c < C; ++c) {
      float part = 0.f;
      for (int h = lane; h < H; h += 32) part = fmaf(a[h], w2s[c * H + h], part);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      if (lane == c) logit = part + b2[c];
    }
    float mx = lane < C ? logit : -INFINITY;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    const float ex = lane < C ? expf(logit - mx) : 0.f;
    float den = ex;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) den += __shfl_xor_sync(0xffffffffu, den, off);
    const float p = ex / den;
    if (y == nullptr) {
      if (proba != nullptr && lane < C) proba[(long long)row * C + lane] = p;
      // argmax, lowest index on ties (numpy argmax)
      float best = lane < C ? logit : -INFINITY;
      int bi = lane;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0 && pred != nullptr) pred[row] = bi;
      continue;
    }
    const int label = y[gather ? gather[row] : row];
    const float w = class_w ? class_w[label] : 1.f;
    const float g = lane < C ? w * (p - (lane == label ? 1.f : 0.f)) : 0.f;
    if (lane < C) dZ2[(long long)row * C + lane] = g;
    const float logp = (logit - mx) - logf(den);
    const float my_loss = __shfl_sync(0xffffffffu, logp, label);
    if (lane == 0) {
      R[2 * (long long)row] = -w * my_loss;
      R[2 * (long long)row + 1] = w;
    }
    for (int h0 = 0; h0 < H; h0 += 32) {  // uniform trip count: the shuffles below need the whole warp
      const int h = h0 + lane;
      const int hc = h < H ? h : H - 1;
      float d = 0.f;
      for (int c = 0; c < C; ++c) d = fmaf(__shfl_sync(0xffffffffu, g, c), w2s[c * H + hc], d);
      if (h < H) dZ1[(long long)row * H + h] = a[h] > 0.f ? d : 0.f;
    }
  }
}

// dst[c] = sum over rows of src[r*ld + c], fixed order: block = 32 columns x 8 row lanes, rows strided by 8, then a
// serial 8-term sum. T_ACC = double for the scaler statistics.
template <typename T_ACC, int MODE>  // MODE 0: x ; 1: (x - mean[c])^2
__global__ void __launch_bounds__(256) head_colsum_kernel(const float* __restrict__ src, long long n, int cols,
                                                          long long ld, const int* __restrict__ gather,
                                                          const double* __restrict__ mean, T_ACC* __restrict__ dst) {
  __shared__ T_ACC part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  T_ACC acc = 0;
  if (c < cols) {
    const T_ACC mu = MODE == 1 ? (T_ACC)mean[c] : (T_ACC)0;
    for (long long r = ty; r < n; r += 8) {
      const long long rr = gather ? gather[r] : r;
      const T_ACC v = (T_ACC)src[rr * ld + c];
      acc += MODE == 1 ? (v - mu) * (v - mu) : v;
    }
  }
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    T_ACC s = part[0][tx];
#pragma unroll
    for (int i = 1; i < 8; ++i) s += part[i][tx];
    dst[c] = s;
  }
}

__global__ void __launch_bounds__(256) head_adam_kernel(float* __restrict__ params, const float* __restrict__ G,
                                                        float* __restrict__ m, float* __restrict__ v, long long P,
                                                        float lr, float beta1, float beta2, float eps, float wd,
                                                        float bc1, float bc2) {
  const float inv_norm = 1.0f / G[P + 1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
    float g = G[i] * inv_norm + wd * params[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * g;
    const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
    m[i] = mi;
    v[i] = vi;
    params[i] -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
  }
}

static int fail(char* err, int err_len, const char* msg) {
  if (err && err_len > 0) snprintf(err, (size_t)err_len, "%s", msg);
  return -1;
}

static int check(cudaError_t ce, const char* what, char* err, int err_len) {
  if (ce == cudaSuccess) return 0;
  char msg[200];
  snprintf(msg, sizeof msg, "%s: %s", what, cudaGetErrorString(ce));
  return fail(err, err_len, msg);
}

static bool dims_ok(int D, int H, int C) { return D > 0 && H > 0 && C > 1 && C <= MAX_C && (long long)C * H * 4 <= 160 * 1024; }

struct Split {
  const float *W1, *b1, *W2, *b2;
};
static Split split(const float* p, int D, int H, int C) {
  Split s;
  s.W1 = p;
  s.b1 = s.W1 + (long long)H * D;
  s.W2 = s.b1 + H;
  s.b2 = s.W2 + (long long)C * H;
  return s;
}

static int hidden_forward(const float* X, const int* gather, long long n, int D, int H, const float* mean,
                          const float* inv_std, const Split& w, float* A1, cudaStream_t st) {
  Operand a{X, D, gather, mean, inv_std}, b{w.W1, D, nullptr, nullptr, nullptr};
  dim3 grid((unsigned)((H + TN - 1) / TN), (unsigned)((n + TM - 1) / TM));
  if (mean != nullptr)
    head_sgemm_kernel<true, true, true, false, true><<<grid, 256, 0, st>>>(a, b, A1, H, (int)n, H, D, w.b1);
  else
    head_sgemm_kernel<true, true, false, false, true><<<grid, 256, 0, st>>>(a, b, A1, H, (int)n, H, D, w.b1);
  return 0;
}

}  // namespace ssr

using namespace ssr;

extern "C" int64_t ssr_head_param_count(int32_t D, int32_t H, int32_t C) {
  if (!dims_ok(D, H, C)) return -1;
  return (int64_t)H * D + H + (int64_t)C * H + C;
}

extern "C" int64_t ssr_head_work_bytes(int64_t n, int32_t D, int32_t H, int32_t C) {
  if (!dims_ok(D, H, C) || n < 0) return -1;
  return (int64_t)sizeof(float) * n * (2LL * H + C + 2) + 256;
}

extern "C" int ssr_head_scaler_stats(const float* X_dev, int64_t n, int32_t D, int64_t ld, const double* mean_dev,
                                     double* out_dev, void* cuda_stream, char* err, int32_t err_len) {
  if (!X_dev || !out_dev || D <= 0 || n < 0) return fail(err, err_len, "ssr_head_scaler_stats: bad argument");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const unsigned blocks = (unsigned)((D + 31) / 32);
  if (mean_dev == nullptr)
    head_colsum_kernel<double, 0><<<blocks, 256, 0, st>>>(X_dev, n, D, ld, nullptr, nullptr, out_dev);
  else
    head_colsum_kernel<double, 1><<<blocks, 256, 0, st>>>(X_dev, n, D, ld, nullptr, mean_dev, out_dev);
  return check(cudaGetLastError(), "ssr_head_scaler_stats", err, err_len);
}

extern "C" int ssr_head_grad(const float* X_dev, const int32_t* y_dev, const int32_t* rows_dev, int64_t n, int32_t D,
                             int32_t H, int32_t C, const float* mean_dev, const float* inv_std_dev,
                             const float* params_dev, const float* class_w_dev, float* grad_dev, void* work_dev,
                             int64_t work_bytes, void* cuda_stream, char* err, int32_t err_len) {
  if (!dims_ok(D, H, C)) return fail(err, err_len, "ssr_head_grad: unsupported dimensions");
  if (!X_dev || !y_dev || !params_dev || !grad_dev) return fail(err, err_len, "ssr_head_grad: null argument");
  if ((mean_dev == nullptr) != (inv_std_dev == nullptr))
    return fail(err, err_len, "ssr_head_grad: mean and inv_std go together");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const long long P = (long long)H * D + H + (long long)C * H + C;
  if (n == 0) return check(cudaMemsetAsync(grad_dev, 0, sizeof(float) * (size_t)(P + 2), st), "memset", err, err_len);
  if (n > 0x7fffffff / 2) return fail(err, err_len, "ssr_head_grad: too many rows in one call");
  if (!work_dev || work_bytes < ssr_head_work_bytes(n, D, H, C))
    return fail(err, err_len, "ssr_head_grad: work buffer too small");
  float* A1 = reinterpret_cast<float*>(work_dev);
  float* dZ1 = A1 + n * H;
  float* dZ2 = dZ1 + n * H;
  float* R = dZ2 + n * C;
  const Split w = split(params_dev, D, H, C);
  float* gW1 = grad_dev;
  float* gb1 = gW1 + (long long)H * D;
  float* gW2 = gb1 + H;
  float* gb2 = gW2 + (long long)C * H;
  float* gLoss = grad_dev + P;

  hidden_forward(X_dev, rows_dev, n, D, H, mean_dev, inv_std_dev, w, A1, st);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(head_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr = true;
  }
  const unsigned oblocks = (unsigned)((n + 7) / 8 < 1184 ? (n + 7) / 8 : 1184);
  head_out_kernel<<<oblocks, 256, sizeof(float) * (size_t)C * H, st>>>(A1, (int)n, H, C, w.W2, w.b2, y_dev, rows_dev,
                                                                     class_w_dev, dZ2, dZ1, R, nullptr, nullptr);
  // dW2 [C, H] = dZ2^T [C, n] . A1 [n, H]
  {
    Operand a{dZ2, C, nullptr, nullptr, nullptr}, b{A1, H, nullptr, nullptr, nullptr};
    dim3 grid((unsigned)((H + TN - 1) / TN), (unsigned)((C + TM - 1) / TM));
    head_sgemm_kernel<false, false, false, false, false><<<grid, 256, 0, st>>>(a, b, gW2, H, C, H, (int)n, nullptr);
  }
  // dW1 [H, D] = dZ1^T [H, n] . Xs [n, D]
  {
    Operand a{dZ1, H, nullptr, nullptr, nullptr}, b{X_dev, D, rows_dev, mean_dev, inv_std_dev};
    dim3 grid((unsigned)((D + TN - 1) / TN), (unsigned)((H + TM - 1) / TM));
    if (mean_dev != nullptr)
      head_sgemm_kernel<false, false, false, true, false><<<grid, 256, 0, st>>>(a, b, gW1, D, H, D, (int)n, nullptr);
    else
      head_sgemm_kernel<false, false, false, false, false><<<grid, 256, 0, st>>>(a, b, gW1, D, H, D, (int)n, nullptr);
  }
  head_colsum_kernel<float, 0><<<(unsigned)((H + 31) / 32), 256, 0, st>>>(dZ1, n, H, H, nullptr, nullptr, gb1);
  head_colsum_kernel<float, 0><<<(unsigned)((C + 31) / 32), 256, 0, st>>>(dZ2, n, C, C, nullptr, nullptr, gb2);
  head_colsum_kernel<float, 0><<<1, 256, 0, st>>>(R, n, 2, 2, nullptr, nullptr, gLoss);
  return check(cudaGetLastError(), "ssr_head_grad", err, err_len);
}

extern "C" int ssr_head_adam(float* params_dev, const float* grad_dev, float* m_dev, float* v_dev, int64_t P, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int32_t step,
                             void* cuda_stream, char* err, int32_t err_len) {
  if (!params_dev || !grad_dev || !m_dev || !v_dev || P <= 0 || step < 1)
    return fail(err, err_len, "ssr_head_adam: bad argument");
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  const unsigned blocks = (unsigned)((P + 255) / 256 < 1184 ? (P + 255) / 256 : 1184);
  head_adam_kernel<<<blocks, 256, 0, (cudaStream_t)cuda_stream>>>(params_dev, grad_dev, m_dev, v_dev, P, lr, beta1,
                                                                  beta2, eps, weight_decay, bc1, bc2);
  return check(cudaGetLastError(), "ssr_head_adam", err, err_len);
}

extern "C" int ssr_head_predict(const float* X_dev, int64_t n, int32_t D, int32_t H, int32_t C, const float* mean_dev,
                                const float* inv_std_dev, const float* params_dev, int32_t* pred_dev,
                                float* proba_dev, void* work_dev, int64_t work_bytes, void* cuda_stream, char* err,
                                int32_t err_len) {
  if (!dims_ok(D, H, C)) return fail(err, err_len, "ssr_head_predict: unsupported dimensions");
  if (!X_dev || !params_dev || (!pred_dev && !proba_dev)) return fail(err, err_len, "ssr_head_predict: null argument");
  if (n == 0) return 0;
  if (n > 0x7fffffff / 2) return fail(err, err_len, "ssr_head_predict: too many rows in one call");
  if (!work_dev || work_bytes < (int64_t)sizeof(float) * n * H)
    return fail(err, err_len, "ssr_head_predict: work buffer too small");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  float* A1 = reinterpret_cast<float*>(work_dev);
  const Split w = split(params_dev, D, H, C);
  hidden_forward(X_dev, nullptr, n, D, H, mean_dev, inv_std_dev, w, A1, st);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(head_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr = true;
  }
  const unsigned oblocks = (unsigned)((n + 7) / 8 < 1184 ? (n + 7) / 8 : 1184);
  head_out_kernel<<<oblocks, 256, sizeof(float) * (size_t)C * H, st>>>(A1, (int)n, H, C, w.W2, w.b2, nullptr, nullptr,
                                                                     nullptr, nullptr, nullptr, nullptr, pred_dev,
                                                                     proba_dev);
  return check(cudaGetLastError(), "ssr_head_predict", err, err_len);
}
