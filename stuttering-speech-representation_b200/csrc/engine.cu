// Model drivers + the C ABI (include/ssr_b200.h).  Host-side C++: weight packing from HF state_dict tensors,
// workspace arena, the WavLM / Whisper-encoder layer loops, fused per-layer time pooling.
//
// Reference call stacks being replaced:
//   WavLMModel.forward          HF/models/wavlm/modeling_wavlm.py:1039-1095  (feature encoder :779-789, projection :100-105,
//                               encoder :388-447 / :465-522, layers :314-373)
//   WhisperEncoder.forward      HF/models/whisper/modeling_whisper.py:593-647 (layer :380-414, attention :284-357)
//   pooling                     REF/WavLM_embeddings.py:321, REF/whisper_embeddings_large.py:278
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ssr_b200.h"
#include "common.cuh"
#include "kernels.cuh"

using namespace ssr;

namespace {

std::string g_create_error;
}  // namespace
namespace ssr {
int g_pdl = 0;
}
namespace {

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t ce__ = (call);                                                                    \
    if (ce__ != cudaSuccess) {                                                                    \
      err = std::string(#call) + ": " + cudaGetErrorString(ce__);                                 \
      return -1;                                                                                  \
    }                                                                                             \
  } while (0)

inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);                                            // round to nearest even
  return (uint16_t)(u >> 16);
}

// Bumped by every workspace (re)allocation: cached device state that bakes pointers in (uploaded lengths, captured
// CUDA graphs) is only valid for the epoch it was made in. Atomic: engines of different host threads share it.
std::atomic<uint64_t> g_arena_epoch{1};

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  ~Buf() {
    if (p) cudaFree(p);
  }
  // Grow-only; new memory is zero-filled (padding regions rely on it and are never written afterwards).
  int ensure(size_t bytes, cudaStream_t st, std::string& err) {
    if (bytes <= cap) return 0;
    ++g_arena_epoch;
    if (p) {
      CK(cudaStreamSynchronize(st));
      CK(cudaFree(p));
      p = nullptr;
      cap = 0;
    }
    bytes = (bytes + 255) & ~size_t(255);
    CK(cudaMalloc(&p, bytes));
    CK(cudaMemsetAsync(p, 0, bytes, st));
    cap = bytes;
    return 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

struct DbgBuf {
  const void* p;
  int64_t bytes;
  int64_t dims[4];
  int32_t dtype;
};

struct ProfEntry {
  std::string name;
  cudaEvent_t a, b;
  double flops;
};

struct LayerW {
  bf16 *wqkv, *wo, *w1, *w2;
  float *bqkv, *bo, *b1, *b2;
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  float *gate_wa, *gate_wb, *gate_const;
  float gate_ba, gate_bb;
};

// Whisper decoder layer, single-token use: self-attention needs only v_proj / out_proj (softmax over one key).
struct DecLayerW {
  bf16 *w_sv, *w_so, *w_cq, *w_ck, *w_cv, *w_co, *w1, *w2;
  float *b_sv, *b_so, *b_cq, *b_cv, *b_co, *b1, *b2;
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
};

}  // namespace

struct ssr_engine {
  ssr_model_desc d;
  int device = 0;
  int num_sms = 148;
  std::string err;
  int64_t launches = 0;
  // options
  int opt_simt = 0, opt_fused_pool = 1, opt_snapshot_layer = -1, opt_profile = 0, opt_attn_simt = 0;
  int opt_posconv_generic = 0;  // 1: positional conv through the generic GEMM kernel (cross-check)
  int opt_conv_ln_fused = 1;    // 1: WavLM-Large conv layers 1-6 as conv + LayerNorm + GELU in one kernel (gemm_ln.cu)
  std::vector<ProfEntry> prof;
  std::string prof_json;
  // Algorithmic FLOPs of the profile entries count LIVE rows: the transformer runs on `slot` rows per clip of which
  // lens[b] are live (149 of 150 for 3 s clips). prof_rows = live rows / slot rows of the current forward, applied
  // to the per-layer GEMMs; prof_att = sum(len^2) / (B * slot^2), applied to attention.
  double prof_rows = 1.0, prof_att = 1.0;

  // ---- weights (device) ----
  std::vector<void*> owned;  // every cudaMalloc'd weight block
  // WavLM front end
  float* c0_w = nullptr;
  float* c0_wstat = nullptr;
  float *cln_g[7] = {}, *cln_b[7] = {};
  bf16* conv_w[7] = {};
  float *fp_ln_g = nullptr, *fp_ln_b = nullptr, *fp_b = nullptr;
  bf16* fp_w = nullptr;
  bf16* pos_w = nullptr;
  float* pos_b = nullptr;
  float *enc_ln_g = nullptr, *enc_ln_b = nullptr;
  std::vector<float> rel_embed_host;  // [320, H]
  Buf relbias;
  int rel_R = 0;
  // Whisper front end
  float *twiddle = nullptr, *melw = nullptr, *dft_tab = nullptr;
  int *mel_lo = nullptr, *mel_hi = nullptr, *mel_band = nullptr;
  bf16 *melw_hi = nullptr, *melw_lo = nullptr;
  int opt_logmel_dense = 0;  // 1: the round-1 dense-DFT log-mel kernel (cross-check)
  bf16 *wc1 = nullptr, *wc2 = nullptr;
  float *bc1 = nullptr, *bc2 = nullptr, *pos_emb = nullptr;
  std::vector<LayerW> layers;
  // Whisper decoder (optional: present when the caller passed decoder.* tensors)
  std::vector<DecLayerW> dec_layers;
  int dec_L = 0, dec_F = 0;
  float *dec_h0 = nullptr, *dec_ln_g = nullptr, *dec_ln_b = nullptr;
  Buf dec_h, dec_x, dec_t1, dec_q, dec_qp, dec_scores, dec_ctx, dec_cv, dec_mid;

  // ---- workspace ----
  Buf nsamp_dev, lens_dev, stats, gn_acc;
  Buf conv[7], feat_ln, feat, xp, posconv, h, tmp, xn, qkv, ctx, mid, gate, pool_part, pool_layers;
  Buf logspec, gmax, conv_in, c1, lens1500;
  Buf audio_stage, pooled_stage;
  Buf snap[8];
  std::map<std::string, DbgBuf> dbg;

  // ---- host-entry plumbing ----
  // what nsamp_dev / lens_dev currently hold (skips two pageable H2D copies when a call repeats the lengths)
  std::vector<int> up_nsamp, up_lens;
  const void *up_ptr_n = nullptr, *up_ptr_l = nullptr;
  cudaStream_t up_stream = nullptr;
  bool capturing = false;
  // Lengths travel host -> device through a small pinned ring owned by the engine: a cudaMemcpyAsync from pageable
  // memory (the caller's n_samples, a stack vector) would be staged by the runtime and synchronise the stream.
  static constexpr int kLenRing = 8;
  int* len_ring = nullptr;        // pinned, kLenRing slots of 2 * len_ring_cap ints (n_samples | frame counts)
  int len_ring_cap = 0;
  int len_ring_next = 0;
  cudaEvent_t len_ring_done[kLenRing] = {};
  // One workspace serves every entry point: a forward on stream B must not start while a forward on stream A is
  // still using it. Every entry point makes its stream wait for the previous forward's completion event.
  cudaEvent_t last_done = nullptr;
  cudaStream_t last_stream = nullptr;
  bool last_valid = false;
  // Host-entry pipeline for large batches (the synchronous *_host call is the end-to-end headline): the batch's
  // audio travels host -> device in chunks on a copy stream while conv0 of the chunks that have landed already runs, and the pooled rows of hidden_states[0 .. L-1] travel device -> host while the last layer still computes.
  struct HostPipe {
    bool active = false;        // set by run_host around one forward
    int n_chunks = 0;
    int chunk_start[9] = {};  // clip ranges [chunk_start[c], chunk_start[c + 1])
    cudaEvent_t h2d_done[8] = {};
    cudaEvent_t pool_early = nullptr;  // recorded on the compute stream once rows 0 .. L-1 of `pooled` are final
    bool pool_early_recorded = false;
    cudaEvent_t fence = nullptr;
  } pipe;
  cudaStream_t copy_stream = nullptr;
  int opt_host_pipeline = 1;
  int opt_host_chunks = 4;  // H2D chunks of the host-entry pipeline (1..8), geometrically growing sizes (measured best)
  // Small host-entry batches are launch-latency bound (about 210 kernels per WavLM-Large forward): the second
  // identical call (same model path, batch, pitch and lengths — the reference's per-clip loop over equal-length
  // clips) is captured into a CUDA graph and replayed from then on.
  int opt_graphs = 1;
  cudaStream_t host_stream = nullptr;
  cudaEvent_t host_fence = nullptr;
  struct GraphSlot {
    bool seen = false;
    int kind = -1, B = 0;
    long long ld = 0;
    std::vector<int> n;
    uint64_t epoch = 0;
    const void *audio = nullptr, *pooled = nullptr, *dec = nullptr;
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
  } graph;
  void drop_graph() {
    if (graph.exec) cudaGraphExecDestroy(graph.exec);
    graph.exec = nullptr;
    graph.seen = false;
  }

  ~ssr_engine() {
    drop_graph();
    for (cudaEvent_t ev : len_ring_done)
      if (ev) cudaEventDestroy(ev);
    if (len_ring) cudaFreeHost(len_ring);
    if (last_done) cudaEventDestroy(last_done);
    for (cudaEvent_t ev : pipe.h2d_done)
      if (ev) cudaEventDestroy(ev);
    if (pipe.pool_early) cudaEventDestroy(pipe.pool_early);
    if (pipe.fence) cudaEventDestroy(pipe.fence);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (host_fence) cudaEventDestroy(host_fence);
    if (host_stream) cudaStreamDestroy(host_stream);
    for (void* p : owned) cudaFree(p);
  }
};

namespace {

// ------------------------------------------------------------------------------------------------ weight access
struct WeightMap {
  std::map<std::string, const ssr_weight*> m;
  std::string err;
  const float* get(const std::string& name, int64_t numel) {
    auto it = m.find(name);
    if (it == m.end()) it = m.find("encoder." + name);
    if (it == m.end()) {
      if (err.empty()) err = "missing weight '" + name + "'";
      return nullptr;
    }
    if (it->second->numel != numel) {
      if (err.empty())
        err = "weight '" + name + "' has " + std::to_string(it->second->numel) + " elements, expected " +
              std::to_string(numel);
      return nullptr;
    }
    return it->second->data;
  }
};

template <typename T>
int upload(ssr_engine* e, const std::vector<T>& host, T** dev, std::string& err) {
  void* p = nullptr;
  CK(cudaMalloc(&p, host.size() * sizeof(T) + 16));
  CK(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  e->owned.push_back(p);
  *dev = reinterpret_cast<T*>(p);
  return 0;
}
int upload_f32(ssr_engine* e, const float* src, int64_t n, float** dev, std::string& err) {
  if (!src) return -1;
  std::vector<float> v(src, src + n);
  return upload(e, v, dev, err);
}
int upload_bf16(ssr_engine* e, const float* src, int64_t n, float scale, bf16** dev, std::string& err) {
  if (!src) return -1;
  std::vector<uint16_t> v((size_t)n);
  for (int64_t i = 0; i < n; ++i) v[i] = f2bf(src[i] * scale);
  uint16_t* d = nullptr;
  if (upload(e, v, &d, err)) return -1;
  *dev = reinterpret_cast<bf16*>(d);
  return 0;
}

// Fused [3D, D] q/k/v weight (q rows pre-scaled by head_dim^-0.5 = 1/8, exact in bf16) and [3D] bias.
int pack_qkv(ssr_engine* e, WeightMap& w, const std::string& pq, const std::string& pk, const std::string& pv, int D,
             bool k_has_bias, LayerW& L, std::string& err) {
  const float* wq = w.get(pq + ".weight", (int64_t)D * D);
  const float* wk = w.get(pk + ".weight", (int64_t)D * D);
  const float* wv = w.get(pv + ".weight", (int64_t)D * D);
  const float* bq = w.get(pq + ".bias", D);
  const float* bk = k_has_bias ? w.get(pk + ".bias", D) : nullptr;
  const float* bv = w.get(pv + ".bias", D);
  if (!wq || !wk || !wv || !bq || !bv || (k_has_bias && !bk)) return -1;
  std::vector<uint16_t> W((size_t)3 * D * D);
  const size_t DD = (size_t)D * D;
  for (size_t i = 0; i < DD; ++i) {
    W[i] = f2bf(wq[i] * 0.125f);
    W[DD + i] = f2bf(wk[i]);
    W[2 * DD + i] = f2bf(wv[i]);
  }
  std::vector<float> B((size_t)3 * D, 0.f);
  for (int i = 0; i < D; ++i) {
    B[i] = bq[i] * 0.125f;
    B[D + i] = bk ? bk[i] : 0.f;
    B[2 * D + i] = bv[i];
  }
  uint16_t* dW = nullptr;
  if (upload(e, W, &dW, err)) return -1;
  L.wqkv = reinterpret_cast<bf16*>(dW);
  return upload(e, B, &L.bqkv, err);
}

// Conv1d weight [Co, Ci, k] -> implicit-GEMM weight [Co, k*Ci] (tap-major, matching channels-last im2col rows).
int pack_conv(ssr_engine* e, const float* w, int Co, int Ci, int k, bf16** dev, std::string& err) {
  if (!w) return -1;
  std::vector<uint16_t> W((size_t)Co * Ci * k);
  for (int co = 0; co < Co; ++co)
    for (int ci = 0; ci < Ci; ++ci)
      for (int t = 0; t < k; ++t) W[((size_t)co * k + t) * Ci + ci] = f2bf(w[((size_t)co * Ci + ci) * k + t]);
  uint16_t* d = nullptr;
  if (upload(e, W, &d, err)) return -1;
  *dev = reinterpret_cast<bf16*>(d);
  return 0;
}

// HF WavLMAttention._relative_positions_bucket (modeling_wavlm.py:252-271), num_buckets 320, max_distance 800.
int rel_bucket(int rel) {
  const int nb = 160, max_exact = 80;
  int bucket = rel > 0 ? nb : 0;
  const int a = rel < 0 ? -rel : rel;
  if (a < max_exact) return bucket + a;
  float v = logf((float)a / (float)max_exact);
  v = v / (float)log(800.0 / 80.0);
  v = v * (float)(nb - max_exact);
  long large = (long)((float)max_exact + v);
  if (large > nb - 1) large = nb - 1;
  return bucket + (int)large;
}

void reg_dbg(ssr_engine* e, const char* name, const void* p, int dtype, int64_t d0, int64_t d1, int64_t d2, int64_t d3);

int build_relbias(ssr_engine* e, int R, cudaStream_t st, std::string& err) {
  if (R <= e->rel_R) return 0;
  int newR = 256;
  while (newR < R) newR *= 2;
  const int H = e->d.heads, W = 2 * newR - 1;
  std::vector<float> tab((size_t)H * W);
  for (int rel = -(newR - 1); rel <= newR - 1; ++rel) {
    const int bkt = rel_bucket(rel);
    for (int h = 0; h < H; ++h) tab[(size_t)h * W + rel + newR - 1] = e->rel_embed_host[(size_t)bkt * H + h];
  }
  if (e->relbias.ensure(tab.size() * 4, st, err)) return -1;
  CK(cudaMemcpyAsync(e->relbias.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  e->rel_R = newR;
  reg_dbg(e, "relbias", e->relbias.p, 0, H, W, 1, 1);
  return 0;
}

int create_wavlm(ssr_engine* e, WeightMap& w, std::string& err) {
  const ssr_model_desc& d = e->d;
  const int D = d.hidden, F = d.ffn, H = d.heads, L = d.layers;
  if (D != H * 64) {
    err = "WavLM: head_dim must be 64";
    return -1;
  }
  if (D % 128 != 0 || D / 16 > 64) {
    err = "WavLM: unsupported hidden size";
    return -1;
  }
  static const int ks[7] = {10, 3, 3, 3, 3, 2, 2};
  const bool layer_norm = d.feat_norm == SSR_FEAT_NORM_LAYER;
  if (upload_f32(e, w.get("feature_extractor.conv_layers.0.conv.weight", 5120), 5120, &e->c0_w, err)) return -1;
  {
    // conv0 + LayerNorm: per-frame channel statistics as a linear / quadratic form of the 10 window samples
    const float* w0 = w.get("feature_extractor.conv_layers.0.conv.weight", 5120);
    std::vector<float> st(112, 0.f);
    for (int k = 0; k < 10; ++k) {
      double m = 0.0;
      for (int c = 0; c < 512; ++c) m += (double)w0[c * 10 + k];
      st[k] = (float)(m / 512.0);
      for (int k2 = k; k2 < 10; ++k2) {
        double gsum = 0.0;
        for (int c = 0; c < 512; ++c) gsum += (double)w0[c * 10 + k] * (double)w0[c * 10 + k2];
        st[10 + k * 10 + k2] = (float)((k2 == k ? 1.0 : 2.0) * gsum / 512.0);
      }
    }
    if (upload(e, st, &e->c0_wstat, err)) return -1;
  }
  for (int i = 0; i < 7; ++i) {
    const std::string p = "feature_extractor.conv_layers." + std::to_string(i);
    if (i > 0 && pack_conv(e, w.get(p + ".conv.weight", 512LL * 512 * ks[i]), 512, 512, ks[i], &e->conv_w[i], err))
      return -1;
    if (layer_norm || i == 0) {
      if (upload_f32(e, w.get(p + ".layer_norm.weight", 512), 512, &e->cln_g[i], err)) return -1;
      if (upload_f32(e, w.get(p + ".layer_norm.bias", 512), 512, &e->cln_b[i], err)) return -1;
    }
  }
  if (upload_f32(e, w.get("feature_projection.layer_norm.weight", 512), 512, &e->fp_ln_g, err)) return -1;
  if (upload_f32(e, w.get("feature_projection.layer_norm.bias", 512), 512, &e->fp_ln_b, err)) return -1;
  if (upload_bf16(e, w.get("feature_projection.projection.weight", 512LL * D), 512LL * D, 1.f, &e->fp_w, err))
    return -1;
  if (upload_f32(e, w.get("feature_projection.projection.bias", D), D, &e->fp_b, err)) return -1;
  {
    // weight-norm (dim = 2): w[co, ci, k] = g[k] * v[co, ci, k] / ||v[:, :, k]||   (modeling_wavlm.py:59-77)
    const int gw = D / 16;
    const float* g = w.get("encoder.pos_conv_embed.conv.parametrizations.weight.original0", 128);
    const float* v = w.get("encoder.pos_conv_embed.conv.parametrizations.weight.original1", (int64_t)D * gw * 128);
    if (!g || !v) return -1;
    std::vector<double> nrm(128, 0.0);
    for (int64_t i = 0; i < (int64_t)D * gw; ++i)
      for (int k = 0; k < 128; ++k) nrm[k] += (double)v[i * 128 + k] * (double)v[i * 128 + k];
    std::vector<float> sc(128);
    for (int k = 0; k < 128; ++k) sc[k] = (float)((double)g[k] / sqrt(nrm[k]));
    std::vector<uint16_t> W((size_t)1024 * 8192, 0);
    for (int grp = 0; grp < 16; ++grp)
      for (int co = 0; co < gw; ++co)
        for (int ci = 0; ci < gw; ++ci)
          for (int k = 0; k < 128; ++k)
            W[((size_t)(grp * 64 + co)) * 8192 + k * 64 + ci] =
                f2bf(v[(((size_t)(grp * gw + co)) * gw + ci) * 128 + k] * sc[k]);
    uint16_t* dW = nullptr;
    if (upload(e, W, &dW, err)) return -1;
    e->pos_w = reinterpret_cast<bf16*>(dW);
    if (upload_f32(e, w.get("encoder.pos_conv_embed.conv.bias", D), D, &e->pos_b, err)) return -1;
  }
  if (upload_f32(e, w.get("encoder.layer_norm.weight", D), D, &e->enc_ln_g, err)) return -1;
  if (upload_f32(e, w.get("encoder.layer_norm.bias", D), D, &e->enc_ln_b, err)) return -1;
  {
    const float* re = w.get("encoder.layers.0.attention.rel_attn_embed.weight", 320LL * H);
    if (!re) return -1;
    e->rel_embed_host.assign(re, re + 320LL * H);
  }
  e->layers.resize(L);
  for (int l = 0; l < L; ++l) {
    LayerW& Lw = e->layers[l];
    const std::string p = "encoder.layers." + std::to_string(l);
    if (pack_qkv(e, w, p + ".attention.q_proj", p + ".attention.k_proj", p + ".attention.v_proj", D, true, Lw, err))
      return -1;
    if (upload_bf16(e, w.get(p + ".attention.out_proj.weight", (int64_t)D * D), (int64_t)D * D, 1.f, &Lw.wo, err))
      return -1;
    if (upload_f32(e, w.get(p + ".attention.out_proj.bias", D), D, &Lw.bo, err)) return -1;
    if (upload_f32(e, w.get(p + ".layer_norm.weight", D), D, &Lw.ln1_g, err)) return -1;
    if (upload_f32(e, w.get(p + ".layer_norm.bias", D), D, &Lw.ln1_b, err)) return -1;
    if (upload_bf16(e, w.get(p + ".feed_forward.intermediate_dense.weight", (int64_t)F * D), (int64_t)F * D, 1.f,
                    &Lw.w1, err))
      return -1;
    if (upload_f32(e, w.get(p + ".feed_forward.intermediate_dense.bias", F), F, &Lw.b1, err)) return -1;
    if (upload_bf16(e, w.get(p + ".feed_forward.output_dense.weight", (int64_t)D * F), (int64_t)D * F, 1.f, &Lw.w2,
                    err))
      return -1;
    if (upload_f32(e, w.get(p + ".feed_forward.output_dense.bias", D), D, &Lw.b2, err)) return -1;
    if (upload_f32(e, w.get(p + ".final_layer_norm.weight", D), D, &Lw.ln2_g, err)) return -1;
    if (upload_f32(e, w.get(p + ".final_layer_norm.bias", D), D, &Lw.ln2_b, err)) return -1;
    const float* gw8 = w.get(p + ".attention.gru_rel_pos_linear.weight", 8 * 64);
    const float* gb8 = w.get(p + ".attention.gru_rel_pos_linear.bias", 8);
    const float* gc = w.get(p + ".attention.gru_rel_pos_const", H);
    if (!gw8 || !gb8 || !gc) return -1;
    std::vector<float> wa(64, 0.f), wb(64, 0.f);
    for (int j = 0; j < 64; ++j) {
      wa[j] = (gw8[0 * 64 + j] + gw8[1 * 64 + j]) + (gw8[2 * 64 + j] + gw8[3 * 64 + j]);
      wb[j] = (gw8[4 * 64 + j] + gw8[5 * 64 + j]) + (gw8[6 * 64 + j] + gw8[7 * 64 + j]);
    }
    Lw.gate_ba = (gb8[0] + gb8[1]) + (gb8[2] + gb8[3]);
    Lw.gate_bb = (gb8[4] + gb8[5]) + (gb8[6] + gb8[7]);
    if (upload(e, wa, &Lw.gate_wa, err)) return -1;
    if (upload(e, wb, &Lw.gate_wb, err)) return -1;
    if (upload_f32(e, gc, H, &Lw.gate_const, err)) return -1;
  }
  return 0;
}

int create_whisper(ssr_engine* e, WeightMap& w, std::string& err) {
  const ssr_model_desc& d = e->d;
  const int D = d.hidden, F = d.ffn, H = d.heads, L = d.layers, NM = d.n_mels;
  if (D != H * 64 || D % 128 != 0 || D > 1280) {
    err = "Whisper: head_dim must be 64 and d_model one of 256, 384, 512, 768, 1024, 1280";
    return -1;
  }
  if (NM != 80) {
    err = "Whisper: only 80 mel bins are supported (whisper-large v1/v2, small, base, ...)";
    return -1;
  }
  // hann-windowed DFT table: column 2f = w[k] cos(2 pi f k / 400), 2f+1 = -w[k] sin(...), periodic hann.
  {
    std::vector<float> tw((size_t)400 * 448, 0.f);
    const double PI = 3.14159265358979323846;
    for (int k = 0; k < 400; ++k) {
      const double win = 0.5 - 0.5 * cos(2.0 * PI * k / 400.0);
      for (int f = 0; f < 201; ++f) {
        const int ph = (int)(((long long)f * k) % 400);
        const double ang = 2.0 * PI * ph / 400.0;
        tw[(size_t)k * 448 + 2 * f] = (float)(win * cos(ang));
        tw[(size_t)k * 448 + 2 * f + 1] = (float)(-win * sin(ang));
      }
    }
    if (upload(e, tw, &e->twiddle, err)) return -1;
    const float* mf = w.get("mel_filters", 201LL * NM);  // [201, 80] as WhisperFeatureExtractor.mel_filters
    if (!mf) return -1;
    std::vector<float> mw((size_t)NM * 201);
    std::vector<int> lo(NM), hi(NM);
    for (int m = 0; m < NM; ++m) {
      lo[m] = 201;
      hi[m] = -1;
      for (int f = 0; f < 201; ++f) {
        const float v = mf[(size_t)f * NM + m];
        mw[(size_t)m * 201 + f] = v;
        if (v != 0.f) {
          if (f < lo[m]) lo[m] = f;
          hi[m] = f;
        }
      }
      if (hi[m] < 0) {
        lo[m] = 0;
        hi[m] = -1;
      }
    }
    if (upload(e, mw, &e->melw, err)) return -1;
    if (upload(e, lo, &e->mel_lo, err)) return -1;
    if (upload(e, hi, &e->mel_hi, err)) return -1;
    // folded-DFT tables (frontend.cu, logmel_folded_kernel): [4][104][104] = {cos, sin} of even k = 2j, then of odd
    // k = 2j + 1; entry [j][f] = hann[k] * cos|sin(2 pi f k / 400) for k <= 200 (sin: 0 < k < 200), f <= 100; else 0.
    {
      std::vector<float> tab((size_t)4 * 104 * 104, 0.f);
      for (int par = 0; par < 2; ++par)
        for (int j = 0; j < 104; ++j) {
          const int k = 2 * j + par;
          if (k > 200) continue;
          const double win = 0.5 - 0.5 * cos(2.0 * PI * k / 400.0);
          for (int f = 0; f <= 100; ++f) {
            const int ph = (int)(((long long)f * k) % 400);
            const double ang = 2.0 * PI * ph / 400.0;
            tab[((size_t)(2 * par) * 104 + j) * 104 + f] = (float)(win * cos(ang));
            if (k > 0 && k < 200) tab[((size_t)(2 * par + 1) * 104 + j) * 104 + f] = (float)(win * sin(ang));
          }
        }
      if (upload(e, tab, &e->dft_tab, err)) return -1;
      // filterbank split into bf16 hi + lo, [mel][208]; per 8-mel tile the range of non-zero 16-bin steps
      std::vector<uint16_t> wh((size_t)NM * 208, 0), wl((size_t)NM * 208, 0);
      std::vector<int> band(2 * (NM / 8));
      for (int nt = 0; nt < NM / 8; ++nt) {
        int blo = 201, bhi = -1;
        for (int m = nt * 8; m < nt * 8 + 8; ++m) {
          for (int f = 0; f < 201; ++f) {
            const float v = mf[(size_t)f * NM + m];
            const uint16_t h = f2bf(v);
            uint32_t hu = (uint32_t)h << 16;
            float hf;
            memcpy(&hf, &hu, 4);
            wh[(size_t)m * 208 + f] = h;
            wl[(size_t)m * 208 + f] = f2bf(v - hf);
          }
          if (hi[m] >= lo[m]) {
            if (lo[m] < blo) blo = lo[m];
            if (hi[m] > bhi) bhi = hi[m];
          }
        }
        band[2 * nt] = bhi < 0 ? 0 : blo / 16;
        band[2 * nt + 1] = bhi < 0 ? 0 : bhi / 16 + 1;
      }
      uint16_t *dh = nullptr, *dl = nullptr;
      if (upload(e, wh, &dh, err)) return -1;
      if (upload(e, wl, &dl, err)) return -1;
      e->melw_hi = reinterpret_cast<bf16*>(dh);
      e->melw_lo = reinterpret_cast<bf16*>(dl);
      if (upload(e, band, &e->mel_band, err)) return -1;
    }
  }
  if (pack_conv(e, w.get("conv1.weight", (int64_t)D * NM * 3), D, NM, 3, &e->wc1, err)) return -1;
  if (upload_f32(e, w.get("conv1.bias", D), D, &e->bc1, err)) return -1;
  if (pack_conv(e, w.get("conv2.weight", (int64_t)D * D * 3), D, D, 3, &e->wc2, err)) return -1;
  if (upload_f32(e, w.get("conv2.bias", D), D, &e->bc2, err)) return -1;
  if (upload_f32(e, w.get("embed_positions.weight", 1500LL * D), 1500LL * D, &e->pos_emb, err)) return -1;
  if (upload_f32(e, w.get("layer_norm.weight", D), D, &e->enc_ln_g, err)) return -1;
  if (upload_f32(e, w.get("layer_norm.bias", D), D, &e->enc_ln_b, err)) return -1;
  e->layers.resize(L);
  for (int l = 0; l < L; ++l) {
    LayerW& Lw = e->layers[l];
    memset(&Lw, 0, sizeof(Lw));
    const std::string p = "layers." + std::to_string(l);
    if (pack_qkv(e, w, p + ".self_attn.q_proj", p + ".self_attn.k_proj", p + ".self_attn.v_proj", D, false, Lw, err))
      return -1;
    if (upload_bf16(e, w.get(p + ".self_attn.out_proj.weight", (int64_t)D * D), (int64_t)D * D, 1.f, &Lw.wo, err))
      return -1;
    if (upload_f32(e, w.get(p + ".self_attn.out_proj.bias", D), D, &Lw.bo, err)) return -1;
    if (upload_f32(e, w.get(p + ".self_attn_layer_norm.weight", D), D, &Lw.ln1_g, err)) return -1;
    if (upload_f32(e, w.get(p + ".self_attn_layer_norm.bias", D), D, &Lw.ln1_b, err)) return -1;
    if (upload_bf16(e, w.get(p + ".fc1.weight", (int64_t)F * D), (int64_t)F * D, 1.f, &Lw.w1, err)) return -1;
    if (upload_f32(e, w.get(p + ".fc1.bias", F), F, &Lw.b1, err)) return -1;
    if (upload_bf16(e, w.get(p + ".fc2.weight", (int64_t)D * F), (int64_t)D * F, 1.f, &Lw.w2, err)) return -1;
    if (upload_f32(e, w.get(p + ".fc2.bias", D), D, &Lw.b2, err)) return -1;
    if (upload_f32(e, w.get(p + ".final_layer_norm.weight", D), D, &Lw.ln2_g, err)) return -1;
    if (upload_f32(e, w.get(p + ".final_layer_norm.bias", D), D, &Lw.ln2_b, err)) return -1;
  }
  // ---- optional decoder (single-token probe): desc.reserved = {decoder_layers, decoder_ffn_dim} ----
  const int Ld = d.reserved[0], Fd = d.reserved[1];
  if (Ld > 0) {
    if (Fd <= 0 || Fd % 256 != 0 || H > 20) {
      err = "Whisper decoder: unsupported decoder_ffn_dim / head count";
      return -1;
    }
    // token 0 at position 0: hidden_states[0] = embed_tokens.weight[0] + embed_positions.weight[0]
    const float* et = w.get("decoder.embed_tokens.weight[0]", D);
    const float* ep = w.get("decoder.embed_positions.weight[0]", D);
    if (!et || !ep) return -1;
    std::vector<float> h0(D);
    for (int i = 0; i < D; ++i) h0[i] = et[i] + ep[i];
    if (upload(e, h0, &e->dec_h0, err)) return -1;
    if (upload_f32(e, w.get("decoder.layer_norm.weight", D), D, &e->dec_ln_g, err)) return -1;
    if (upload_f32(e, w.get("decoder.layer_norm.bias", D), D, &e->dec_ln_b, err)) return -1;
    e->dec_layers.resize(Ld);
    const int64_t DD = (int64_t)D * D;
    for (int l = 0; l < Ld; ++l) {
      DecLayerW& W = e->dec_layers[l];
      memset(&W, 0, sizeof(W));
      const std::string p = "decoder.layers." + std::to_string(l);
      if (upload_bf16(e, w.get(p + ".self_attn.v_proj.weight", DD), DD, 1.f, &W.w_sv, err)) return -1;
      if (upload_f32(e, w.get(p + ".self_attn.v_proj.bias", D), D, &W.b_sv, err)) return -1;
      if (upload_bf16(e, w.get(p + ".self_attn.out_proj.weight", DD), DD, 1.f, &W.w_so, err)) return -1;
      if (upload_f32(e, w.get(p + ".self_attn.out_proj.bias", D), D, &W.b_so, err)) return -1;
      if (upload_f32(e, w.get(p + ".self_attn_layer_norm.weight", D), D, &W.ln1_g, err)) return -1;
      if (upload_f32(e, w.get(p + ".self_attn_layer_norm.bias", D), D, &W.ln1_b, err)) return -1;
      // cross-attention: q pre-scaled by head_dim^-0.5 (exact: 1/8), k has no bias
      if (upload_bf16(e, w.get(p + ".encoder_attn.q_proj.weight", DD), DD, 0.125f, &W.w_cq, err)) return -1;
      {
        const float* bq = w.get(p + ".encoder_attn.q_proj.bias", D);
        if (!bq) return -1;
        std::vector<float> b(bq, bq + D);
        for (float& x : b) x *= 0.125f;
        if (upload(e, b, &W.b_cq, err)) return -1;
      }
      if (upload_bf16(e, w.get(p + ".encoder_attn.k_proj.weight", DD), DD, 1.f, &W.w_ck, err)) return -1;
      if (upload_bf16(e, w.get(p + ".encoder_attn.v_proj.weight", DD), DD, 1.f, &W.w_cv, err)) return -1;
      if (upload_f32(e, w.get(p + ".encoder_attn.v_proj.bias", D), D, &W.b_cv, err)) return -1;
      if (upload_bf16(e, w.get(p + ".encoder_attn.out_proj.weight", DD), DD, 1.f, &W.w_co, err)) return -1;
      if (upload_f32(e, w.get(p + ".encoder_attn.out_proj.bias", D), D, &W.b_co, err)) return -1;
      if (upload_f32(e, w.get(p + ".encoder_attn_layer_norm.weight", D), D, &W.ln2_g, err)) return -1;
      if (upload_f32(e, w.get(p + ".encoder_attn_layer_norm.bias", D), D, &W.ln2_b, err)) return -1;
      if (upload_bf16(e, w.get(p + ".fc1.weight", (int64_t)Fd * D), (int64_t)Fd * D, 1.f, &W.w1, err)) return -1;
      if (upload_f32(e, w.get(p + ".fc1.bias", Fd), Fd, &W.b1, err)) return -1;
      if (upload_bf16(e, w.get(p + ".fc2.weight", (int64_t)D * Fd), (int64_t)D * Fd, 1.f, &W.w2, err)) return -1;
      if (upload_f32(e, w.get(p + ".fc2.bias", D), D, &W.b2, err)) return -1;
      if (upload_f32(e, w.get(p + ".final_layer_norm.weight", D), D, &W.ln3_g, err)) return -1;
      if (upload_f32(e, w.get(p + ".final_layer_norm.bias", D), D, &W.ln3_b, err)) return -1;
    }
    e->dec_L = Ld;
    e->dec_F = Fd;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ step helpers
EpiParams epi_plain(const float* bias, int act, const float* resid, int ldr, float* o32, int ld32, bf16* o16,
                    int ld16) {
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = bias;
  ep.act = act;
  ep.resid = resid;
  ep.ldr = ldr;
  ep.out_f32 = o32;
  ep.ldo32 = ld32;
  ep.out_bf16 = o16;
  ep.ldo16 = ld16;
  return ep;
}

// Per-launch CUDA-event timing (option "profile"): events go on the launching stream around each kernel group.
struct ProfScope {
  ssr_engine* e;
  cudaStream_t st;
  int idx = -1;
  ProfScope(ssr_engine* e_, cudaStream_t st_, const char* name, double flops = 0.0) : e(e_), st(st_) {
    if (!e->opt_profile) return;
    ProfEntry pe;
    pe.name = name;
    pe.flops = flops;
    cudaEventCreate(&pe.a);
    cudaEventCreate(&pe.b);
    cudaEventRecord(pe.a, st);
    e->prof.push_back(pe);
    idx = (int)e->prof.size() - 1;
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(e->prof[idx].b, st);
  }
};

int run_gemm(ssr_engine* e, const GemmOp& op, cudaStream_t st, const char* name = "gemm") {
  e->launches += e->opt_simt ? (op.epi.pool_part ? 2 : 1) : 1;
  const bool layer_gemm = name[0] == 'g' && (!strcmp(name, "gemm_qkv") || !strcmp(name, "gemm_out") ||
                                              !strcmp(name, "gemm_ffn1") || !strcmp(name, "gemm_ffn2") ||
                                              !strcmp(name, "gemm_proj"));
  ProfScope ps(e, st, name, 2.0 * (double)op.M * (double)op.N * (double)op.K * (layer_gemm ? e->prof_rows : 1.0));
  return launch_gemm(op, st, e->opt_simt != 0, e->num_sms, e->err);
}

int run_ln(ssr_engine* e, const LayerNormArgs& a, cudaStream_t st) {
  e->launches++;
  ProfScope ps(e, st, a.gelu ? "layernorm_gelu" : "layernorm");
  return launch_layernorm(a, st, e->err);
}
int run_attn(ssr_engine* e, const AttentionArgs& a, cudaStream_t st) {
  e->launches++;
  // QK^T and PV: 2 * 2 * T^2 * 64 per head, T = the clip's live frames
  ProfScope ps(e, st, "attention", 4.0 * (double)a.B * a.H * (double)a.slot * a.slot * 64.0 * e->prof_att);
  return e->opt_attn_simt ? launch_attention(a, st, e->err) : launch_attention_tc(a, st, e->err);
}

GemmOp linear_op(const bf16* A, int M, int K, const bf16* W, int N, const EpiParams& ep) {
  GemmOp op;
  op.A = A;
  op.lda = K;
  op.a_rows = M;
  op.W = W;
  op.M = M;
  op.N = N;
  op.K = K;
  op.a_mode = 0;
  op.a_cols = 0;
  op.epi = ep;
  return op;
}

void reg_dbg(ssr_engine* e, const char* name, const void* p, int dtype, int64_t d0, int64_t d1, int64_t d2 = 1,
             int64_t d3 = 1) {
  DbgBuf b;
  b.p = p;
  b.dtype = dtype;
  b.dims[0] = d0;
  b.dims[1] = d1;
  b.dims[2] = d2;
  b.dims[3] = d3;
  b.bytes = d0 * d1 * d2 * d3 * (dtype == 0 ? 4 : 2);
  e->dbg[name] = b;
}

int snapshot(ssr_engine* e, int slot_idx, const char* name, const void* src, int dtype, int64_t rows, int64_t cols,
             cudaStream_t st) {
  std::string& err = e->err;
  const size_t bytes = (size_t)rows * cols * (dtype == 0 ? 4 : 2);
  if (e->snap[slot_idx].ensure(bytes, st, err)) return -1;
  CK(cudaMemcpyAsync(e->snap[slot_idx].p, src, bytes, cudaMemcpyDeviceToDevice, st));
  reg_dbg(e, name, e->snap[slot_idx].p, dtype, rows, cols);
  return 0;
}

int wavlm_frames(int n) {
  static const int ks[7] = {10, 3, 3, 3, 3, 2, 2}, ss[7] = {5, 2, 2, 2, 2, 2, 2};
  long long len = n;
  for (int i = 0; i < 7; ++i) {
    if (len < ks[i]) return 0;
    len = (len - ks[i]) / ss[i] + 1;
  }
  return (int)len;
}

int pool_into(ssr_engine* e, const float* x, int B, int slot, int D, float* pooled, int layer, int L1,
              cudaStream_t st) {
  e->launches += 1;
  ProfScope ps(e, st, "pool_mean");
  return launch_pool_mean(x, B, slot, D, e->lens_dev.as<int>(), pooled + (long long)layer * D, (long long)L1 * D, st,
                          e->err);
}

// One transformer layer stack shared by WavLM (both LayerNorm placements) and Whisper.
//   pre_ln  (WavLM-Large "stable", Whisper): h += Attn(LN1(h)); h += FFN(LN2(h))
//   post_ln (WavLM-Base+):                   h = LN1(h + Attn(h)); h = LN2(h + FFN(h))
int run_layers(ssr_engine* e, int B, int slot, bool pre_ln, bool wavlm, float* pooled, cudaStream_t st) {
  std::string& err = e->err;
  const int D = e->d.hidden, F = e->d.ffn, H = e->d.heads, L = e->d.layers, L1 = L + 1;
  const int M = B * slot;
  const int* lens = e->lens_dev.as<int>();
  float* h = e->h.as<float>();
  float* tmp = e->tmp.as<float>();
  bf16* xn = e->xn.as<bf16>();
  bf16* qkv = e->qkv.as<bf16>();
  bf16* ctx = e->ctx.as<bf16>();
  bf16* mid = e->mid.as<bf16>();
  float* gate = wavlm ? e->gate.as<float>() : nullptr;
  float* part = e->pool_part.as<float>();
  const bool fused = e->opt_fused_pool && slot >= 32;
  // pre-LN stacks: every layer's output projection leaves its pooling partials in its own slice, and ONE finalize
  // launch after the loop reduces all of them (23 / 31 tiny launches otherwise)
  const size_t part_stride = (size_t)ceil_div(M, 32) * 2 * D;
  float* part_layers = nullptr;
  if (pre_ln && fused && L > 1) {
    if (e->pool_layers.ensure((size_t)(L - 1) * part_stride * 4, st, err)) return -1;
    part_layers = e->pool_layers.as<float>();
  }

  for (int l = 0; l < L; ++l) {
    const LayerW& W = e->layers[l];
    const bool snap = (e->opt_snapshot_layer == l);
    if (pre_ln) {
      LayerNormArgs a;
      memset(&a, 0, sizeof(a));
      a.in_f32 = h;
      a.rows = M;
      a.D = D;
      a.ld_in = D;
      a.gamma = W.ln1_g;
      a.beta = W.ln1_b;
      a.eps = 1e-5f;
      a.out_bf16 = xn;
      a.ld_out16 = D;
      if (wavlm) {
        a.gate_out = gate;
        a.gate_wa = W.gate_wa;
        a.gate_wb = W.gate_wb;
        a.gate_ba = W.gate_ba;
        a.gate_bb = W.gate_bb;
        a.gate_const = W.gate_const;
        a.n_heads = H;
      }
      if (run_ln(e, a, st)) return -1;
    }
    // (post-LN: xn / gate for this layer were produced by the previous LayerNorm)
    if (snap && snapshot(e, 0, "L.attn_in", xn, 1, M, D, st)) return -1;
    if (snap && wavlm && snapshot(e, 1, "L.gate", gate, 0, M, H, st)) return -1;
    if (run_gemm(e, linear_op(xn, M, D, W.wqkv, 3 * D, epi_plain(W.bqkv, ACT_NONE, nullptr, 0, nullptr, 0, qkv, 3 * D)),
                 st, "gemm_qkv"))
      return -1;
    if (snap && snapshot(e, 2, "L.qkv", qkv, 1, M, 3 * D, st)) return -1;
    {
      AttentionArgs a;
      memset(&a, 0, sizeof(a));
      a.qkv = qkv;
      a.out = ctx;
      a.B = B;
      a.slot = slot;
      a.H = H;
      a.D = D;
      a.lens = lens;
      if (wavlm) {
        a.gate = gate;
        a.relbias = e->relbias.as<float>();
        a.rel_stride = 2 * e->rel_R - 1;
        a.rel_center = e->rel_R - 1;
      }
      if (run_attn(e, a, st)) return -1;
    }
    if (snap && snapshot(e, 3, "L.ctx", ctx, 1, M, D, st)) return -1;
    if (pre_ln) {
      if (run_gemm(e, linear_op(ctx, M, D, W.wo, D, epi_plain(W.bo, ACT_NONE, h, D, h, D, nullptr, 0)), st, "gemm_out")) return -1;
      if (snap && snapshot(e, 4, "L.h_attn", h, 0, M, D, st)) return -1;
      LayerNormArgs a;
      memset(&a, 0, sizeof(a));
      a.in_f32 = h;
      a.rows = M;
      a.D = D;
      a.ld_in = D;
      a.gamma = W.ln2_g;
      a.beta = W.ln2_b;
      a.eps = 1e-5f;
      a.out_bf16 = xn;
      a.ld_out16 = D;
      if (run_ln(e, a, st)) return -1;
      if (run_gemm(e, linear_op(xn, M, D, W.w1, F, epi_plain(W.b1, ACT_GELU, nullptr, 0, nullptr, 0, mid, F)), st, "gemm_ffn1"))
        return -1;
      if (snap && snapshot(e, 5, "L.mid", mid, 1, M, F, st)) return -1;
      EpiParams ep = epi_plain(W.b2, ACT_NONE, h, D, h, D, nullptr, 0);
      const bool pool_here = (l < L - 1);  // hidden_states[l+1] = this layer's output (the last one is LN'd first)
      if (pool_here && fused) {
        ep.pool_part = part_layers + (size_t)l * part_stride;
        ep.pool_slot = slot;
        ep.lens = lens;
      }
      if (run_gemm(e, linear_op(mid, M, F, W.w2, D, ep), st, "gemm_ffn2")) return -1;
      if (snap && snapshot(e, 6, "L.h_out", h, 0, M, D, st)) return -1;
      if (pool_here && !fused && pool_into(e, h, B, slot, D, pooled, l + 1, L1, st)) return -1;
      if (fused && l == L - 2) {  // the last pooled-here layer is done: hidden_states[1 .. L-1] in one launch
        e->launches++;
        ProfScope ps(e, st, "pool_finalize");
        if (launch_pool_finalize(part_layers, B, slot, D, lens, pooled + D, (long long)L1 * D, st, err, L - 1,
                                 (long long)part_stride, (long long)D))
          return -1;
        if (e->pipe.active && e->pipe.pool_early) {  // rows 0 .. L-1 of `pooled` are final: their D2H may start
          CK(cudaEventRecord(e->pipe.pool_early, st));
          e->pipe.pool_early_recorded = true;
        }
      }
    } else {
      // post-LN (WavLM Base+)
      if (run_gemm(e, linear_op(ctx, M, D, W.wo, D, epi_plain(W.bo, ACT_NONE, h, D, tmp, D, nullptr, 0)), st, "gemm_out"))
        return -1;
      LayerNormArgs a;
      memset(&a, 0, sizeof(a));
      a.in_f32 = tmp;
      a.rows = M;
      a.D = D;
      a.ld_in = D;
      a.gamma = W.ln1_g;
      a.beta = W.ln1_b;
      a.eps = 1e-5f;
      a.out_f32 = h;
      a.ld_out32 = D;
      a.out_bf16 = xn;
      a.ld_out16 = D;
      if (run_ln(e, a, st)) return -1;
      if (snap && snapshot(e, 4, "L.h_attn", h, 0, M, D, st)) return -1;
      if (run_gemm(e, linear_op(xn, M, D, W.w1, F, epi_plain(W.b1, ACT_GELU, nullptr, 0, nullptr, 0, mid, F)), st, "gemm_ffn1"))
        return -1;
      if (snap && snapshot(e, 5, "L.mid", mid, 1, M, F, st)) return -1;
      if (run_gemm(e, linear_op(mid, M, F, W.w2, D, epi_plain(W.b2, ACT_NONE, h, D, tmp, D, nullptr, 0)), st, "gemm_ffn2"))
        return -1;
      memset(&a, 0, sizeof(a));
      a.in_f32 = tmp;
      a.rows = M;
      a.D = D;
      a.ld_in = D;
      a.gamma = W.ln2_g;
      a.beta = W.ln2_b;
      a.eps = 1e-5f;
      a.out_f32 = h;
      a.ld_out32 = D;
      a.out_bf16 = xn;
      a.ld_out16 = D;
      if (wavlm && l + 1 < L) {
        const LayerW& Wn = e->layers[l + 1];
        a.gate_out = gate;
        a.gate_wa = Wn.gate_wa;
        a.gate_wb = Wn.gate_wb;
        a.gate_ba = Wn.gate_ba;
        a.gate_bb = Wn.gate_bb;
        a.gate_const = Wn.gate_const;
        a.n_heads = H;
      }
      if (run_ln(e, a, st)) return -1;
      if (snap && snapshot(e, 6, "L.h_out", h, 0, M, D, st)) return -1;
      if (pool_into(e, h, B, slot, D, pooled, l + 1, L1, st)) return -1;
    }
  }
  if (pre_ln) {
    // hidden_states[L] = final LayerNorm of the last layer's output
    LayerNormArgs a;
    memset(&a, 0, sizeof(a));
    a.in_f32 = h;
    a.rows = M;
    a.D = D;
    a.ld_in = D;
    a.gamma = e->enc_ln_g;
    a.beta = e->enc_ln_b;
    a.eps = 1e-5f;
    a.out_f32 = tmp;
    a.ld_out32 = D;
    if (!wavlm && e->dec_L > 0) {  // bf16 copy of last_hidden_state for the decoder's cross-attention
      a.out_bf16 = xn;
      a.ld_out16 = D;
    }
    if (run_ln(e, a, st)) return -1;
    if (pool_into(e, tmp, B, slot, D, pooled, L, L1, st)) return -1;
    reg_dbg(e, "last_hidden", tmp, 0, M, D);
  } else {
    reg_dbg(e, "last_hidden", h, 0, M, D);
  }
  return 0;
}

int upload_lengths(ssr_engine* e, const int32_t* n_samples, int B, const std::vector<int>& lens, cudaStream_t st) {
  std::string& err = e->err;
  if (e->nsamp_dev.ensure(sizeof(int) * B, st, err)) return -1;
  if (e->lens_dev.ensure(sizeof(int) * B, st, err)) return -1;
  // unchanged since the previous call (the usual case in a per-clip loop): the device copies are still right
  // (same buffers, same stream only: the earlier copy is then ordered before this call's kernels)
  if (e->up_ptr_n == e->nsamp_dev.p && e->up_ptr_l == e->lens_dev.p && e->up_stream == st &&
      (int)e->up_nsamp.size() == B && e->up_lens == lens &&
      memcmp(e->up_nsamp.data(), n_samples, sizeof(int) * B) == 0)
    return 0;
  if (e->capturing) {  // a captured graph contains no length upload: it relies on what the buffers hold
    err = "length upload needed during graph capture";
    return -1;
  }
  // The device length buffers are about to change: a captured graph was recorded against their old contents.
  e->drop_graph();
  e->up_nsamp.clear();  // stays empty if a copy below fails
  if (B > e->len_ring_cap) {
    // grow the pinned ring (rare: only when the batch size exceeds every earlier one); all slots must be idle
    for (cudaEvent_t ev : e->len_ring_done)
      if (ev) CK(cudaEventSynchronize(ev));
    if (e->len_ring) CK(cudaFreeHost(e->len_ring));
    e->len_ring = nullptr;
    e->len_ring_cap = 0;
    int cap = 64;
    while (cap < B) cap *= 2;
    CK(cudaMallocHost(reinterpret_cast<void**>(&e->len_ring), sizeof(int) * 2 * (size_t)cap * ssr_engine::kLenRing));
    e->len_ring_cap = cap;
    for (cudaEvent_t& ev : e->len_ring_done)
      if (!ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  const int k = e->len_ring_next;
  e->len_ring_next = (k + 1) % ssr_engine::kLenRing;
  CK(cudaEventSynchronize(e->len_ring_done[k]));  // the copy issued kLenRing uploads ago; returns at once if unused
  int* slot_n = e->len_ring + (size_t)k * 2 * e->len_ring_cap;
  int* slot_l = slot_n + e->len_ring_cap;
  memcpy(slot_n, n_samples, sizeof(int) * B);
  memcpy(slot_l, lens.data(), sizeof(int) * B);
  CK(cudaMemcpyAsync(e->nsamp_dev.p, slot_n, sizeof(int) * B, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(e->lens_dev.p, slot_l, sizeof(int) * B, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(e->len_ring_done[k], st));
  e->up_nsamp.assign(n_samples, n_samples + B);
  e->up_lens = lens;
  e->up_ptr_n = e->nsamp_dev.p;
  e->up_ptr_l = e->lens_dev.p;
  e->up_stream = st;
  return 0;
}

// Entry / exit of every forward: cross-stream ordering on the shared workspace (see ssr_engine::last_done).
int forward_enter(ssr_engine* e, cudaStream_t st) {
  std::string& err = e->err;
  if (e->last_valid && e->last_stream != st) CK(cudaStreamWaitEvent(st, e->last_done, 0));
  return 0;
}
int forward_leave(ssr_engine* e, cudaStream_t st) {
  std::string& err = e->err;
  if (!e->last_done) CK(cudaEventCreateWithFlags(&e->last_done, cudaEventDisableTiming));
  CK(cudaEventRecord(e->last_done, st));
  e->last_stream = st;
  e->last_valid = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------ WavLM forward
int wavlm_forward(ssr_engine* e, const float* audio, int64_t audio_ld, const int32_t* n_samples, int B,
                  float* pooled, cudaStream_t st) {
  std::string& err = e->err;
  const ssr_model_desc& d = e->d;
  const int D = d.hidden, H = d.heads, L1 = d.layers + 1;
  static const int ks[7] = {10, 3, 3, 3, 3, 2, 2}, ss[7] = {5, 2, 2, 2, 2, 2, 2};
  if (B <= 0) return 0;
  int max_n = 0;
  std::vector<int> lens(B);
  for (int b = 0; b < B; ++b) {
    if (n_samples[b] < 0 || n_samples[b] > audio_ld) {
      err = "n_samples[" + std::to_string(b) + "] out of range";
      return -1;
    }
    lens[b] = wavlm_frames(n_samples[b]);
    if (lens[b] <= 0) {
      err = "clip " + std::to_string(b) + " is too short for the WavLM feature encoder (needs >= 400 samples)";
      return -1;
    }
    if (n_samples[b] > max_n) max_n = n_samples[b];
  }
  // Slotted flat layout: every clip owns S0 = roundup(max_n, 320) samples, hence S0/5, S0/10, ... S0/320 rows in the
  // conv stack. Row t of clip b of layer i is flat row b*slot_i + t, so each strided conv is ONE 2-D GEMM view.
  const long long S0 = ((long long)max_n + 319) / 320 * 320;
  int slots[7];
  {
    long long s = S0;
    for (int i = 0; i < 7; ++i) {
      s /= ss[i];
      slots[i] = (int)s;
    }
  }
  const int slot = slots[6];
  const long long Mll = (long long)B * slot;
  if ((long long)B * slots[0] > 2000000000LL) {
    err = "batch too large for 32-bit row indexing; split the batch";
    return -1;
  }
  const int M = (int)Mll;
  {
    double live = 0.0, live2 = 0.0;
    for (int b = 0; b < B; ++b) {
      live += lens[b];
      live2 += (double)lens[b] * lens[b];
    }
    e->prof_rows = live / ((double)B * slot);
    e->prof_att = live2 / ((double)B * slot * slot);
  }
  if (upload_lengths(e, n_samples, B, lens, st)) return -1;
  if (build_relbias(e, slot, st, err)) return -1;

  // workspace
  for (int i = 0; i < 7; ++i)
    if (e->conv[i].ensure(((size_t)B * slots[i] + 8) * 512 * 2, st, err)) return -1;
  if (e->stats.ensure(sizeof(float) * 2 * B, st, err)) return -1;
  if (d.feat_norm == SSR_FEAT_NORM_GROUP && e->gn_acc.ensure(sizeof(double) * 1024 * B, st, err)) return -1;
  const int pslot = slot + 128;
  if (e->feat_ln.ensure((size_t)M * 512 * 2, st, err)) return -1;
  if (e->feat.ensure((size_t)M * D * 4, st, err)) return -1;
  if (e->xp.ensure(((size_t)B * pslot + 256) * 1024 * 2, st, err)) return -1;
  if (e->posconv.ensure((size_t)B * pslot * 1024 * 4, st, err)) return -1;
  if (e->h.ensure((size_t)M * D * 4, st, err)) return -1;
  if (e->tmp.ensure((size_t)M * D * 4, st, err)) return -1;
  if (e->xn.ensure((size_t)M * D * 2, st, err)) return -1;
  if (e->qkv.ensure((size_t)M * 3 * D * 2, st, err)) return -1;
  if (e->ctx.ensure((size_t)M * D * 2, st, err)) return -1;
  if (e->mid.ensure((size_t)M * d.ffn * 2, st, err)) return -1;
  if (e->gate.ensure((size_t)M * H * 4, st, err)) return -1;
  if (e->pool_part.ensure((size_t)ceil_div(M, 32) * 2 * D * 4, st, err)) return -1;

  // 1. waveform normalisation + conv0 (+ norm + GELU), 2. conv layers 1..6 as implicit GEMMs over the channels-last
  // signal. With the host pipeline active (run_host, LayerNorm variant) conv0 runs chunk by chunk as the chunks'
  // host-to-device copies land; the conv GEMMs run once over the whole batch.
  const bool ln = d.feat_norm == SSR_FEAT_NORM_LAYER;
  const bool fused_ln = ln && e->opt_conv_ln_fused && !e->opt_simt;
  const bool piped = e->pipe.active && fused_ln && e->pipe.n_chunks > 1;
  auto conv0_range = [&](int c0, int c1) -> int {
    Conv0Args a;
    memset(&a, 0, sizeof(a));
    a.audio = audio + (long long)c0 * audio_ld;
    a.audio_ld = audio_ld;
    a.n_samples = e->nsamp_dev.as<int>() + c0;
    a.B = c1 - c0;
    a.do_normalize = d.do_normalize;
    a.stats = e->stats.as<float>() + 2 * c0;
    a.w = e->c0_w;
    a.wstat = e->c0_wstat;
    a.gamma = e->cln_g[0];
    a.beta = e->cln_b[0];
    a.mode = ln ? 0 : 1;
    a.gn_acc = e->gn_acc.as<double>() + (ln ? 0 : (long long)1024 * c0);
    a.out = e->conv[0].as<bf16>() + (long long)c0 * slots[0] * 512;
    a.slot0 = slots[0];
    e->launches += a.mode == 0 ? 2 : 4;
    ProfScope ps(e, st, "wavlm_conv0", 2.0 * 10 * 512 * (double)(c1 - c0) * slots[0]);
    return launch_wavlm_conv0(a, st, err);
  };
  auto conv_range = [&](int i, int c0, int c1) -> int {
    const int Mi = (c1 - c0) * slots[i];
    GemmOp op;
    op.A = e->conv[i - 1].as<bf16>() + (long long)c0 * slots[i - 1] * 512;
    op.lda = (long long)ss[i] * 512;
    op.a_rows = Mi;
    op.W = e->conv_w[i];
    op.M = Mi;
    op.N = 512;
    op.K = ks[i] * 512;
    op.a_mode = 0;
    op.a_cols = 0;
    bf16* out = e->conv[i].as<bf16>() + (long long)c0 * slots[i] * 512;
    if (fused_ln) {
      e->launches++;
      ProfScope ps(e, st, "gemm_conv_ln", 2.0 * (double)Mi * 512.0 * (double)op.K);
      return launch_gemm_ln(op.A, op.lda, op.a_rows, op.W, Mi, op.K, e->cln_g[i], e->cln_b[i], 1e-5f, out, st,
                            e->num_sms, err);
    }
    op.epi = epi_plain(nullptr, ln ? ACT_NONE : ACT_GELU, nullptr, 0, nullptr, 0, out, 512);
    if (run_gemm(e, op, st, "gemm_conv")) return -1;
    if (ln) {
      LayerNormArgs a;
      memset(&a, 0, sizeof(a));
      a.in_bf16 = out;
      a.rows = Mi;
      a.D = 512;
      a.ld_in = 512;
      a.gamma = e->cln_g[i];
      a.beta = e->cln_b[i];
      a.eps = 1e-5f;
      a.gelu = 1;
      a.out_bf16 = out;
      a.ld_out16 = 512;
      if (run_ln(e, a, st)) return -1;
    }
    return 0;
  };
  int first_full_layer = 1;
  if (piped) {
    for (int c = 0; c < e->pipe.n_chunks; ++c) {
      const int c0 = e->pipe.chunk_start[c], c1 = e->pipe.chunk_start[c + 1];
      if (c0 >= c1) continue;
      CK(cudaStreamWaitEvent(st, e->pipe.h2d_done[c], 0));
      if (conv0_range(c0, c1)) return -1;
    }
    // (conv1 per chunk as well was measured: its 4-CTA-cluster tiles quantise worse on a quarter batch than the
    // copy time it would hide; conv0 alone, 0.35 ms per chunk, already covers the next chunk's 0.22 ms copy)
  } else {
    if (e->pipe.active)  // pipeline requested but not applicable to this model variant: wait for every chunk
      for (int c = 0; c < e->pipe.n_chunks; ++c) CK(cudaStreamWaitEvent(st, e->pipe.h2d_done[c], 0));
    if (conv0_range(0, B)) return -1;
  }
  reg_dbg(e, "conv0", e->conv[0].p, 1, B, slots[0], 512);
  for (int i = first_full_layer; i < 7; ++i)
    if (conv_range(i, 0, B)) return -1;
  for (int i = 1; i < 7; ++i) {
    char nm[16];
    snprintf(nm, sizeof nm, "conv%d", i);
    reg_dbg(e, nm, e->conv[i].p, 1, B, slots[i], 512);
  }
  // 3. feature projection: LayerNorm(512) -> Linear(512 -> D)
  {
    LayerNormArgs a;
    memset(&a, 0, sizeof(a));
    a.in_bf16 = e->conv[6].as<bf16>();
    a.rows = M;
    a.D = 512;
    a.ld_in = 512;
    a.gamma = e->fp_ln_g;
    a.beta = e->fp_ln_b;
    a.eps = 1e-5f;
    a.out_bf16 = e->feat_ln.as<bf16>();
    a.ld_out16 = 512;
    if (run_ln(e, a, st)) return -1;
    if (run_gemm(e,
                 linear_op(e->feat_ln.as<bf16>(), M, 512, e->fp_w, D,
                           epi_plain(e->fp_b, ACT_NONE, nullptr, 0, e->feat.as<float>(), D, nullptr, 0)),
                 st, "gemm_proj"))
      return -1;
    reg_dbg(e, "feat", e->feat.p, 0, B, slot, D);
  }
  // 4. positional conv embedding (grouped conv k=128 as 16 block-diagonal GEMMs with K = 128 taps x 64 channels)
  {
    e->launches++;
    {
      ProfScope ps(e, st, "posconv_pack");
      if (launch_posconv_pack(e->feat.as<float>(), B, slot, D, e->lens_dev.as<int>(), e->xp.as<bf16>(), pslot, st,
                              err))
        return -1;
    }
    const bool stable = d.stable_ln != 0;
    float* dst = stable ? e->h.as<float>() : e->tmp.as<float>();
    // With 64-channel groups (D = 1024) the conv epilogue finishes the stage in place:
    //   h[b, t, :] = feat[b, t, :] + gelu(conv + bias)      for t < len_b
    const bool fused_finish = (D == 1024) && !e->opt_simt && !e->opt_posconv_generic;
    if (!e->opt_simt && !e->opt_posconv_generic) {
      PosConvOp op;
      op.X = e->xp.as<bf16>();
      op.x_rows = (long long)B * pslot + 128;
      op.W = e->pos_w;
      op.B = B;
      op.pslot = pslot;
      op.rows_per_clip = slot;
      if (fused_finish) {
        op.epi = epi_plain(e->pos_b, ACT_GELU, e->feat.as<float>(), D, dst, D, nullptr, 0);
        op.epi.in_slot = pslot;
        op.epi.out_slot = slot;
        op.epi.valid = slot;
        op.epi.lens = e->lens_dev.as<int>();
      } else {
        op.epi = epi_plain(nullptr, ACT_NONE, nullptr, 0, e->posconv.as<float>(), 1024, nullptr, 0);
      }
      e->launches++;
      ProfScope ps(e, st, "gemm_posconv", 2.0 * (double)B * slot * 1024.0 * 8192.0);
      if (launch_posconv(op, st, e->num_sms, err)) return -1;
    } else {
      GemmOp op;
      op.A = e->xp.as<bf16>();
      op.lda = 1024;
      op.a_rows = (long long)B * pslot + 128;
      op.W = e->pos_w;
      op.M = B * pslot;
      op.N = 1024;
      op.K = 8192;
      op.a_mode = 1;
      op.a_cols = 1024;
      op.epi = epi_plain(nullptr, ACT_NONE, nullptr, 0, e->posconv.as<float>(), 1024, nullptr, 0);
      if (run_gemm(e, op, st, "gemm_posconv")) return -1;
    }
    if (!fused_finish) {
      e->launches++;
      ProfScope ps(e, st, "posconv_finish");
      if (launch_posconv_finish(e->posconv.as<float>(), pslot, e->pos_b, e->feat.as<float>(), B, slot, D,
                                e->lens_dev.as<int>(), dst, st, err))
        return -1;
    }
    if (!stable) {
      // Base+: encoder.layer_norm right after the positional add; its output is hidden_states[0]
      LayerNormArgs a;
      memset(&a, 0, sizeof(a));
      a.in_f32 = dst;
      a.rows = M;
      a.D = D;
      a.ld_in = D;
      a.gamma = e->enc_ln_g;
      a.beta = e->enc_ln_b;
      a.eps = 1e-5f;
      a.out_f32 = e->h.as<float>();
      a.ld_out32 = D;
      a.out_bf16 = e->xn.as<bf16>();
      a.ld_out16 = D;
      const LayerW& W0 = e->layers[0];
      a.gate_out = e->gate.as<float>();
      a.gate_wa = W0.gate_wa;
      a.gate_wb = W0.gate_wb;
      a.gate_ba = W0.gate_ba;
      a.gate_bb = W0.gate_bb;
      a.gate_const = W0.gate_const;
      a.n_heads = H;
      if (run_ln(e, a, st)) return -1;
    }
    if (pool_into(e, e->h.as<float>(), B, slot, D, pooled, 0, L1, st)) return -1;
    if (e->opt_snapshot_layer >= 0 && snapshot(e, 7, "hs0", e->h.p, 0, M, D, st)) return -1;
  }
  // 5. transformer
  return run_layers(e, B, slot, d.stable_ln != 0, true, pooled, st);
}

// ------------------------------------------------------------------------------------------------ Whisper forward
int whisper_logmel(ssr_engine* e, const float* audio, int64_t audio_ld, const int32_t* n_samples, int B,
                   float* mel_out, bool want_conv_in, cudaStream_t st) {
  std::string& err = e->err;
  int max_n = 0;
  std::vector<int> lens(B, 1500);
  for (int b = 0; b < B; ++b) {
    if (n_samples[b] < 0 || n_samples[b] > audio_ld) {
      err = "n_samples[" + std::to_string(b) + "] out of range";
      return -1;
    }
    if (n_samples[b] > max_n) max_n = n_samples[b];
  }
  if (upload_lengths(e, n_samples, B, lens, st)) return -1;
  if (e->logspec.ensure((size_t)B * 3000 * 80 * 4, st, err)) return -1;
  if (e->gmax.ensure((size_t)B * 4, st, err)) return -1;
  if (want_conv_in && e->conv_in.ensure(((size_t)B * 3002 + 8) * 80 * 2, st, err)) return -1;
  LogMelArgs a;
  memset(&a, 0, sizeof(a));
  a.audio = audio;
  a.audio_ld = audio_ld;
  a.n_samples = e->nsamp_dev.as<int>();
  a.B = B;
  a.max_samples = max_n;
  a.twiddle = e->twiddle;
  a.melw = e->melw;
  a.mel_lo = e->mel_lo;
  a.mel_hi = e->mel_hi;
  a.dft_tab = e->dft_tab;
  a.melw_hi = e->melw_hi;
  a.melw_lo = e->melw_lo;
  a.mel_band = e->mel_band;
  a.dense = e->opt_logmel_dense;
  a.logspec = e->logspec.as<float>();
  a.gmax = e->gmax.as<unsigned int>();
  a.mel_out = mel_out;
  a.conv_in = want_conv_in ? e->conv_in.as<bf16>() : nullptr;
  e->launches += 3;
  ProfScope ps(e, st, "logmel");
  return launch_logmel(a, st, err);
}

// Decoder single-token probe over the encoder's last_hidden_state (bf16 copy in e->xn, written by run_layers):
// dec_out[b, i, :] = decoder hidden_states[i] for the start token, i = 0..Ld.  See decoder.cu for the algebra.
int whisper_decoder_token(ssr_engine* e, int B, float* dec_out, cudaStream_t st) {
  std::string& err = e->err;
  const int D = e->d.hidden, H = e->d.heads, Ld = e->dec_L, Fd = e->dec_F, T = 1500;
  const long long ldo = (long long)(Ld + 1) * D;
  if (e->dec_h.ensure((size_t)B * D * 4, st, err)) return -1;
  if (e->dec_x.ensure((size_t)B * D * 2, st, err)) return -1;
  if (e->dec_t1.ensure((size_t)B * D * 2, st, err)) return -1;
  if (e->dec_q.ensure((size_t)B * D * 4, st, err)) return -1;
  if (e->dec_qp.ensure((size_t)B * H * D * 4, st, err)) return -1;
  // row splits per clip never exceed max(1, #SMs / B): (B + 256) slots cover every batch size
  if (e->dec_scores.ensure((size_t)(B + 256) * H * 2 * 4, st, err)) return -1;
  if (e->dec_ctx.ensure((size_t)(B + 256) * H * D * 4, st, err)) return -1;
  if (e->dec_cv.ensure((size_t)B * D * 2, st, err)) return -1;
  if (e->dec_mid.ensure((size_t)B * Fd * 2, st, err)) return -1;
  float* h = e->dec_h.as<float>();
  bf16* x = e->dec_x.as<bf16>();
  bf16* t1 = e->dec_t1.as<bf16>();
  bf16* cv = e->dec_cv.as<bf16>();
  bf16* mid = e->dec_mid.as<bf16>();
  float* q = e->dec_q.as<float>();

  auto ln = [&](const float* g, const float* b, float* o32, long long ld32, bf16* o16) {
    LayerNormArgs a;
    memset(&a, 0, sizeof(a));
    a.in_f32 = h;
    a.rows = B;
    a.D = D;
    a.ld_in = D;
    a.gamma = g;
    a.beta = b;
    a.eps = 1e-5f;
    a.out_f32 = o32;
    a.ld_out32 = ld32;
    a.out_bf16 = o16;
    a.ld_out16 = D;
    return run_ln(e, a, st);
  };
  auto save_state = [&](int i) -> int {
    CK(cudaMemcpy2DAsync(dec_out + (long long)i * D, (size_t)ldo * 4, h, (size_t)D * 4, (size_t)D * 4, (size_t)B,
                         cudaMemcpyDeviceToDevice, st));
    return 0;
  };

  // token-level Linear: a handful of rows stream the weights through the GEMV kernel, larger batches use the
  // tensor-core GEMM
  const bool small = dec_gemv_applicable(B, D) && dec_gemv_applicable(B, Fd) && !e->opt_simt;
  auto lin = [&](const bf16* A, int K, const bf16* Wt, int N, const EpiParams& ep) -> int {
    if (!small) return run_gemm(e, linear_op(A, B, K, Wt, N, ep), st, "gemm_dec");
    e->launches++;
    ProfScope ps(e, st, "gemv_dec", 2.0 * (double)B * (double)N * (double)K);
    return launch_dec_gemv(A, B, K, Wt, N, ep, st, err);
  };

  e->launches++;
  if (launch_bcast_rows(e->dec_h0, h, B, D, D, st, err)) return -1;
  if (save_state(0)) return -1;
  for (int l = 0; l < Ld; ++l) {
    const DecLayerW& W = e->dec_layers[l];
    // self-attention over a single token == out_proj(v_proj(LN(h)))
    if (ln(W.ln1_g, W.ln1_b, nullptr, 0, x)) return -1;
    if (lin(x, D, W.w_sv, D, epi_plain(W.b_sv, ACT_NONE, nullptr, 0, nullptr, 0, t1, D))) return -1;
    if (lin(t1, D, W.w_so, D, epi_plain(W.b_so, ACT_NONE, h, D, h, D, nullptr, 0))) return -1;
    // cross-attention over the 1500 encoder states
    if (ln(W.ln2_g, W.ln2_b, nullptr, 0, x)) return -1;
    if (lin(x, D, W.w_cq, D, epi_plain(W.b_cq, ACT_NONE, nullptr, 0, q, D, nullptr, 0))) return -1;
    {
      DecCrossArgs a;
      a.q = q;
      a.wk = W.w_ck;
      a.wv = W.w_cv;
      a.bv = W.b_cv;
      a.enc = e->xn.as<bf16>();
      a.qp = e->dec_qp.as<float>();
      a.ml = e->dec_scores.as<float>();
      a.ctx_part = e->dec_ctx.as<float>();
      a.out = cv;
      a.B = B;
      a.T = T;
      a.D = D;
      a.H = H;
      e->launches += 3;
      ProfScope ps(e, st, "dec_cross_attention", 4.0 * (double)B * H * T * D);
      if (launch_dec_cross_attention(a, st, err)) return -1;
    }
    if (lin(cv, D, W.w_co, D, epi_plain(W.b_co, ACT_NONE, h, D, h, D, nullptr, 0))) return -1;
    // feed-forward
    if (ln(W.ln3_g, W.ln3_b, nullptr, 0, x)) return -1;
    if (lin(x, D, W.w1, Fd, epi_plain(W.b1, ACT_GELU, nullptr, 0, nullptr, 0, mid, Fd))) return -1;
    if (lin(mid, Fd, W.w2, D, epi_plain(W.b2, ACT_NONE, h, D, h, D, nullptr, 0))) return -1;
    if (l < Ld - 1 && save_state(l + 1)) return -1;
  }
  // hidden_states[Ld] = decoder.layer_norm(h), written straight into its output slot
  return ln(e->dec_ln_g, e->dec_ln_b, dec_out + (long long)Ld * D, ldo, nullptr);
}

int whisper_forward(ssr_engine* e, const float* audio, int64_t audio_ld, const int32_t* n_samples, int B,
                    float* pooled, cudaStream_t st, float* dec_out = nullptr) {
  std::string& err = e->err;
  const ssr_model_desc& d = e->d;
  const int D = d.hidden, L1 = d.layers + 1;
  if (B <= 0) return 0;
  if (dec_out != nullptr && e->dec_L == 0) {
    err = "this Whisper engine was created without decoder weights";
    return -1;
  }
  if ((long long)B * 3002 > 2000000000LL / 1) {
    err = "batch too large";
    return -1;
  }
  if (whisper_logmel(e, audio, audio_ld, n_samples, B, nullptr, true, st)) return -1;
  e->prof_rows = e->prof_att = 1.0;
  const int M = B * 1500;
  if (e->c1.ensure(((size_t)B * 3002 + 8) * D * 2, st, err)) return -1;
  if (e->h.ensure((size_t)M * D * 4, st, err)) return -1;
  if (e->tmp.ensure((size_t)M * D * 4, st, err)) return -1;
  if (e->xn.ensure((size_t)M * D * 2, st, err)) return -1;
  if (e->qkv.ensure((size_t)M * 3 * D * 2, st, err)) return -1;
  if (e->ctx.ensure((size_t)M * D * 2, st, err)) return -1;
  if (e->mid.ensure((size_t)M * d.ffn * 2, st, err)) return -1;
  if (e->pool_part.ensure((size_t)ceil_div(B * 1501, 32) * 2 * D * 4, st, err)) return -1;
  reg_dbg(e, "conv_in", e->conv_in.p, 1, B, 3002, 80);
  // conv1: k=3, pad=1 over the zero-padded channels-last mel; GELU; written into conv2's padded input
  {
    GemmOp op;
    op.A = e->conv_in.as<bf16>();
    op.lda = 80;
    op.a_rows = (long long)B * 3002;
    op.W = e->wc1;
    op.M = B * 3002;
    op.N = D;
    op.K = 240;
    op.a_mode = 0;
    op.a_cols = 0;
    op.epi = epi_plain(e->bc1, ACT_GELU, nullptr, 0, nullptr, 0, e->c1.as<bf16>(), D);
    op.epi.in_slot = 3002;
    op.epi.out_slot = 3002;
    op.epi.valid = 3000;
    op.epi.out_off = 1;
    if (run_gemm(e, op, st, "gemm_conv1")) return -1;
    reg_dbg(e, "c1", e->c1.p, 1, B, 3002, D);
  }
  // conv2: k=3, stride 2, pad=1; GELU; + embed_positions; this is hidden_states[0]
  {
    GemmOp op;
    op.A = e->c1.as<bf16>();
    op.lda = 2LL * D;
    op.a_rows = (long long)B * 1501;
    op.W = e->wc2;
    op.M = B * 1501;
    op.N = D;
    op.K = 3 * D;
    op.a_mode = 0;
    op.a_cols = 0;
    op.epi = epi_plain(e->bc2, ACT_GELU, e->pos_emb, D, e->h.as<float>(), D, nullptr, 0);
    op.epi.in_slot = 1501;
    op.epi.out_slot = 1500;
    op.epi.valid = 1500;
    op.epi.out_off = 0;
    op.epi.resid_by_t = 1;
    const bool fused = e->opt_fused_pool != 0;
    if (fused) op.epi.pool_part = e->pool_part.as<float>();
    if (run_gemm(e, op, st, "gemm_conv2")) return -1;
    if (fused) {
      e->launches++;
      ProfScope ps(e, st, "pool_finalize");
      if (launch_pool_finalize(e->pool_part.as<float>(), B, 1501, D, e->lens_dev.as<int>(), pooled,
                               (long long)L1 * D, st, err))
        return -1;
    } else if (pool_into(e, e->h.as<float>(), B, 1500, D, pooled, 0, L1, st)) {
      return -1;
    }
    if (e->opt_snapshot_layer >= 0 && snapshot(e, 7, "hs0", e->h.p, 0, M, D, st)) return -1;
  }
  if (run_layers(e, B, 1500, true, false, pooled, st)) return -1;
  if (dec_out != nullptr) return whisper_decoder_token(e, B, dec_out, st);
  return 0;
}

cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

void copy_err(const std::string& s, char* err, int32_t err_len) {
  if (err && err_len > 0) {
    snprintf(err, (size_t)err_len, "%s", s.c_str());
  }
}

}  // namespace

// ================================================================================================== C ABI
extern "C" {

int ssr_create(const ssr_model_desc* desc, const ssr_weight* weights, int32_t n_weights, int32_t cuda_device,
               ssr_engine** out) {
  if (!desc || !out || (!weights && n_weights > 0)) {
    g_create_error = "ssr_create: null argument";
    return -1;
  }
  *out = nullptr;
  std::string err;
  cudaError_t ce = cudaSetDevice(cuda_device);
  if (ce != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(ce);
    return -2;
  }
  cudaDeviceProp prop;
  ce = cudaGetDeviceProperties(&prop, cuda_device);
  if (ce != cudaSuccess) {
    g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(ce);
    return -2;
  }
  if (prop.major != 10) {
    g_create_error = "ssr_b200 kernels are built for sm_100a (Blackwell B200) only; found compute capability " +
                     std::to_string(prop.major) + "." + std::to_string(prop.minor);
    return -3;
  }
  std::unique_ptr<ssr_engine> e(new ssr_engine());
  e->d = *desc;
  e->device = cuda_device;
  e->num_sms = prop.multiProcessorCount;
  WeightMap wm;
  for (int i = 0; i < n_weights; ++i) {
    if (!weights[i].name || !weights[i].data) {
      g_create_error = "ssr_create: weight entry with null name or data";
      return -1;
    }
    wm.m[weights[i].name] = &weights[i];
  }
  int rc;
  if (desc->family == SSR_WAVLM)
    rc = create_wavlm(e.get(), wm, err);
  else if (desc->family == SSR_WHISPER_ENC)
    rc = create_whisper(e.get(), wm, err);
  else {
    g_create_error = "ssr_create: unknown model family";
    return -1;
  }
  if (rc) {
    g_create_error = !wm.err.empty() ? wm.err : (err.empty() ? "ssr_create failed" : err);
    return -4;
  }
  *out = e.release();
  return 0;
}

void ssr_destroy(ssr_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  delete e;
}

const char* ssr_last_error(const ssr_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int ssr_set_option(ssr_engine* e, const char* key, int32_t value) {
  if (!e || !key) return -1;
  const std::string k(key);
  if (k == "simt_gemm")
    e->opt_simt = value;
  else if (k == "fused_pool")
    e->opt_fused_pool = value;
  else if (k == "snapshot_layer")
    e->opt_snapshot_layer = value;
  else if (k == "profile")
    e->opt_profile = value;
  else if (k == "attn_simt")
    e->opt_attn_simt = value;
  else if (k == "posconv_generic")
    e->opt_posconv_generic = value;
  else if (k == "graphs")
    e->opt_graphs = value;
  else if (k == "conv_ln_fused")
    e->opt_conv_ln_fused = value;
  else if (k == "logmel_dense")
    e->opt_logmel_dense = value;
  else if (k == "host_pipeline")
    e->opt_host_pipeline = value;
  else if (k == "host_chunks")
    e->opt_host_chunks = value;
  else {
    e->err = "unknown option '" + k + "'";
    return -1;
  }
  e->drop_graph();  // a captured graph bakes the options of its capture in
  return 0;
}

int ssr_wavlm_pooled(ssr_engine* e, const float* audio_dev, int64_t audio_ld, const int32_t* n_samples, int32_t B,
                     float* pooled_dev, void* cuda_stream) {
  if (!e) return -1;
  if (e->d.family != SSR_WAVLM) {
    e->err = "engine is not a WavLM engine";
    return -1;
  }
  if (!audio_dev || !n_samples || !pooled_dev || B < 0) {
    e->err = "ssr_wavlm_pooled: null argument";
    return -1;
  }
  cudaSetDevice(e->device);
  cudaStream_t st = as_stream(cuda_stream);
  if (forward_enter(e, st)) return -1;
  const int rc = wavlm_forward(e, audio_dev, audio_ld, n_samples, B, pooled_dev, st);
  if (forward_leave(e, st)) return -1;
  return rc;
}

int ssr_whisper_enc_pooled(ssr_engine* e, const float* audio_dev, int64_t audio_ld, const int32_t* n_samples,
                           int32_t B, float* pooled_dev, void* cuda_stream) {
  if (!e) return -1;
  if (e->d.family != SSR_WHISPER_ENC) {
    e->err = "engine is not a Whisper engine";
    return -1;
  }
  if (!audio_dev || !n_samples || !pooled_dev || B < 0) {
    e->err = "ssr_whisper_enc_pooled: null argument";
    return -1;
  }
  cudaSetDevice(e->device);
  cudaStream_t st = as_stream(cuda_stream);
  if (forward_enter(e, st)) return -1;
  const int rc = whisper_forward(e, audio_dev, audio_ld, n_samples, B, pooled_dev, st);
  if (forward_leave(e, st)) return -1;
  return rc;
}

int ssr_logmel(ssr_engine* e, const float* audio_dev, int64_t audio_ld, const int32_t* n_samples, int32_t B,
               float* mel_dev, void* cuda_stream) {
  if (!e) return -1;
  if (e->d.family != SSR_WHISPER_ENC) {
    e->err = "engine is not a Whisper engine";
    return -1;
  }
  if (!audio_dev || !n_samples || !mel_dev || B < 0) {
    e->err = "ssr_logmel: null argument";
    return -1;
  }
  cudaSetDevice(e->device);
  if (B == 0) return 0;
  cudaStream_t st = as_stream(cuda_stream);
  if (forward_enter(e, st)) return -1;
  const int rc = whisper_logmel(e, audio_dev, audio_ld, n_samples, B, mel_dev, false, st);
  if (forward_leave(e, st)) return -1;
  return rc;
}

// The host entry points run on a private stream (the legacy default stream cannot be captured into a graph). It is
// ordered after everything already queued on the legacy default stream — which is where torch's default work and the
// device entry points of a default-stream caller live — and is synchronised before the entry point returns.
static int host_entry_stream(ssr_engine* e, cudaStream_t* out) {
  std::string& err = e->err;
  if (!e->host_stream) {
    CK(cudaStreamCreateWithFlags(&e->host_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->host_fence, cudaEventDisableTiming));
  }
  CK(cudaEventRecord(e->host_fence, nullptr));
  CK(cudaStreamWaitEvent(e->host_stream, e->host_fence, 0));
  *out = e->host_stream;
  return 0;
}

static int forward_dispatch(ssr_engine* e, int kind, const float* audio, int64_t ld, const int32_t* n, int B,
                            float* pooled, float* dec, cudaStream_t st) {
  return kind == 0 ? wavlm_forward(e, audio, ld, n, B, pooled, st) : whisper_forward(e, audio, ld, n, B, pooled, st, dec);
}

// Eager on the first sight of a (path, batch, pitch, lengths, buffers) signature, captured into a CUDA graph on the
// second, replayed afterwards. Anything that could change what the captured kernels should do invalidates the slot:
// a different signature, a workspace reallocation (arena epoch), ssr_set_option.
static int forward_graphed(ssr_engine* e, int kind, const float* audio, int64_t ld, const int32_t* n, int B,
                           float* pooled, float* dec, cudaStream_t st) {
  std::string& err = e->err;
  const bool eligible = e->opt_graphs && !e->opt_profile && e->opt_snapshot_layer < 0 && B > 0 && B <= 16;
  if (!eligible) return forward_dispatch(e, kind, audio, ld, n, B, pooled, dec, st);
  ssr_engine::GraphSlot& g = e->graph;
  const bool same = g.seen && g.kind == kind && g.B == B && g.ld == ld && g.epoch == g_arena_epoch &&
                    g.audio == audio && g.pooled == pooled && g.dec == dec &&
                    memcmp(g.n.data(), n, sizeof(int) * B) == 0;
  if (same && g.exec) {
    CK(cudaGraphLaunch(g.exec, st));
    e->launches += g.launches;
    return 0;
  }
  if (same) {
    cudaGraph_t graph = nullptr;
    const int64_t l0 = e->launches;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    e->capturing = true;
    const int rc = forward_dispatch(e, kind, audio, ld, n, B, pooled, dec, st);
    e->capturing = false;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc == 0 && ce == cudaSuccess && graph != nullptr && g.epoch == g_arena_epoch &&
        cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
      cudaGraphDestroy(graph);
      g.launches = e->launches - l0;
      CK(cudaGraphLaunch(g.exec, st));
      return 0;
    }
    // capture did not work out (e.g. a first-use allocation inside): forget it and run this call eagerly
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    e->drop_graph();
    e->launches = l0;
    e->opt_graphs = 0;  // do not try again on this engine
    return forward_dispatch(e, kind, audio, ld, n, B, pooled, dec, st);
  }
  e->drop_graph();
  const int rc = forward_dispatch(e, kind, audio, ld, n, B, pooled, dec, st);
  if (rc) return rc;
  g.seen = true;
  g.kind = kind;
  g.B = B;
  g.ld = ld;
  g.n.assign(n, n + B);
  g.epoch = g_arena_epoch;
  g.audio = audio;
  g.pooled = pooled;
  g.dec = dec;
  return 0;
}

static int run_host(ssr_engine* e, bool wavlm, const float* audio_host, int64_t audio_ld, const int32_t* n_samples,
                    int32_t B, float* pooled_host) {
  std::string& err = e->err;
  if (!audio_host || !n_samples || !pooled_host || B < 0) {
    err = "host entry point: null argument";
    return -1;
  }
  if (B == 0) return 0;
  cudaSetDevice(e->device);
  cudaStream_t st = nullptr;
  if (host_entry_stream(e, &st)) return -1;
  if (forward_enter(e, st)) return -1;
  const int L1 = e->d.layers + 1, D = e->d.hidden;
  const size_t in_bytes = (size_t)B * audio_ld * 4;
  const size_t out_bytes = (size_t)B * L1 * D * 4;
  if (e->audio_stage.ensure(in_bytes, st, err)) return -1;
  if (e->pooled_stage.ensure(out_bytes, st, err)) return -1;
  // Large batches: chunked H2D on a copy stream overlapping the front end, early D2H of the pooled rows (HostPipe).
  const bool pipe_ok = e->opt_host_pipeline && B >= 64 && !e->opt_profile && e->opt_snapshot_layer < 0;
  ssr_engine::HostPipe& hp = e->pipe;
  if (pipe_ok) {
    if (!e->copy_stream) {
      CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
      for (cudaEvent_t& ev : hp.h2d_done) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&hp.pool_early, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&hp.fence, cudaEventDisableTiming));
    }
    cudaStream_t cs = e->copy_stream;
    // the copy stream starts after everything the compute stream has been made to wait for (previous forward,
    // workspace growth)
    CK(cudaEventRecord(hp.fence, st));
    CK(cudaStreamWaitEvent(cs, hp.fence, 0));
    // Chunk sizes grow geometrically: only the FIRST chunk's copy is exposed (nothing to overlap it with), and conv0
    // of chunk c (~5 us per clip) has to cover the copy of chunk c + 1 (~4 us per clip), so each chunk may be ~1.4x
    // the one before it. opt_host_chunks caps the count; with 1 the whole batch is one copy.
    hp.n_chunks = wavlm ? std::max(1, std::min(8, e->opt_host_chunks)) : 1;
    {
      double w[8], tot = 0.0;
      for (int c = 0; c < hp.n_chunks; ++c) tot += (w[c] = pow(1.4, c));
      int at = 0;
      hp.chunk_start[0] = 0;
      for (int c = 0; c < hp.n_chunks; ++c) {
        int sz = (int)(B * w[c] / tot + 0.5);
        if (c == hp.n_chunks - 1 || at + sz > B) sz = B - at;
        at += sz;
        hp.chunk_start[c + 1] = at;
      }
      hp.chunk_start[hp.n_chunks] = B;
    }
    for (int c = 0; c < hp.n_chunks; ++c) {
      const int c0 = hp.chunk_start[c], c1 = hp.chunk_start[c + 1];
      if (c0 < c1)
        CK(cudaMemcpyAsync(e->audio_stage.as<float>() + (size_t)c0 * audio_ld, audio_host + (size_t)c0 * audio_ld,
                           (size_t)(c1 - c0) * audio_ld * 4, cudaMemcpyHostToDevice, cs));
      CK(cudaEventRecord(hp.h2d_done[c], cs));
    }
    hp.active = true;
    hp.pool_early_recorded = false;
    int rc;
    if (wavlm) {
      rc = wavlm_forward(e, e->audio_stage.as<float>(), audio_ld, n_samples, B, e->pooled_stage.as<float>(), st);
    } else {
      CK(cudaStreamWaitEvent(st, hp.h2d_done[0], 0));
      rc = whisper_forward(e, e->audio_stage.as<float>(), audio_ld, n_samples, B, e->pooled_stage.as<float>(), st);
    }
    hp.active = false;
    if (rc) {
      cudaStreamSynchronize(cs);
      return rc;
    }
    const size_t pitch = (size_t)L1 * D * 4;
    if (hp.pool_early_recorded) {
      CK(cudaStreamWaitEvent(cs, hp.pool_early, 0));
      CK(cudaMemcpy2DAsync(pooled_host, pitch, e->pooled_stage.p, pitch, (size_t)(L1 - 1) * D * 4, (size_t)B,
                           cudaMemcpyDeviceToHost, cs));
      CK(cudaMemcpy2DAsync(pooled_host + (size_t)(L1 - 1) * D, pitch, e->pooled_stage.as<float>() + (size_t)(L1 - 1) * D,
                           pitch, (size_t)D * 4, (size_t)B, cudaMemcpyDeviceToHost, st));
      CK(cudaEventRecord(hp.fence, cs));
      CK(cudaStreamWaitEvent(st, hp.fence, 0));
    } else {
      CK(cudaMemcpyAsync(pooled_host, e->pooled_stage.p, out_bytes, cudaMemcpyDeviceToHost, st));
    }
    if (forward_leave(e, st)) return -1;
    CK(cudaStreamSynchronize(st));
    return 0;
  }
  CK(cudaMemcpyAsync(e->audio_stage.p, audio_host, in_bytes, cudaMemcpyHostToDevice, st));
  int rc = forward_graphed(e, wavlm ? 0 : 1, e->audio_stage.as<float>(), audio_ld, n_samples, B,
                           e->pooled_stage.as<float>(), nullptr, st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(pooled_host, e->pooled_stage.p, out_bytes, cudaMemcpyDeviceToHost, st));
  if (forward_leave(e, st)) return -1;
  CK(cudaStreamSynchronize(st));
  return 0;
}

int ssr_wavlm_pooled_host(ssr_engine* e, const float* audio_host, int64_t audio_ld, const int32_t* n_samples,
                          int32_t B, float* pooled_host) {
  if (!e) return -1;
  if (e->d.family != SSR_WAVLM) {
    e->err = "engine is not a WavLM engine";
    return -1;
  }
  return run_host(e, true, audio_host, audio_ld, n_samples, B, pooled_host);
}

int ssr_whisper_enc_pooled_host(ssr_engine* e, const float* audio_host, int64_t audio_ld, const int32_t* n_samples,
                                int32_t B, float* pooled_host) {
  if (!e) return -1;
  if (e->d.family != SSR_WHISPER_ENC) {
    e->err = "engine is not a Whisper engine";
    return -1;
  }
  return run_host(e, false, audio_host, audio_ld, n_samples, B, pooled_host);
}

int32_t ssr_decoder_layers(const ssr_engine* e) { return e ? e->dec_L : -1; }

int ssr_whisper_full(ssr_engine* e, const float* audio_dev, int64_t audio_ld, const int32_t* n_samples, int32_t B,
                     float* pooled_dev, float* dec_dev, void* cuda_stream) {
  if (!e) return -1;
  if (e->d.family != SSR_WHISPER_ENC) {
    e->err = "engine is not a Whisper engine";
    return -1;
  }
  if (!audio_dev || !n_samples || !pooled_dev || !dec_dev || B < 0) {
    e->err = "ssr_whisper_full: null argument";
    return -1;
  }
  cudaSetDevice(e->device);
  cudaStream_t st = as_stream(cuda_stream);
  if (forward_enter(e, st)) return -1;
  const int rc = whisper_forward(e, audio_dev, audio_ld, n_samples, B, pooled_dev, st, dec_dev);
  if (forward_leave(e, st)) return -1;
  return rc;
}

int ssr_whisper_full_host(ssr_engine* e, const float* audio_host, int64_t audio_ld, const int32_t* n_samples,
                          int32_t B, float* pooled_host, float* dec_host) {
  if (!e) return -1;
  std::string& err = e->err;
  if (e->d.family != SSR_WHISPER_ENC) {
    err = "engine is not a Whisper engine";
    return -1;
  }
  if (!audio_host || !n_samples || !pooled_host || !dec_host || B < 0) {
    err = "ssr_whisper_full_host: null argument";
    return -1;
  }
  if (B == 0) return 0;
  cudaSetDevice(e->device);
  cudaStream_t st = nullptr;
  if (host_entry_stream(e, &st)) return -1;
  if (forward_enter(e, st)) return -1;
  const size_t in_bytes = (size_t)B * audio_ld * 4;
  const size_t enc_bytes = (size_t)B * (e->d.layers + 1) * e->d.hidden * 4;
  const size_t dec_bytes = (size_t)B * (e->dec_L + 1) * e->d.hidden * 4;
  if (e->audio_stage.ensure(in_bytes, st, err)) return -1;
  if (e->pooled_stage.ensure(enc_bytes + dec_bytes, st, err)) return -1;
  float* enc_dev = e->pooled_stage.as<float>();
  float* dec_dev = reinterpret_cast<float*>(reinterpret_cast<char*>(e->pooled_stage.p) + enc_bytes);
  CK(cudaMemcpyAsync(e->audio_stage.p, audio_host, in_bytes, cudaMemcpyHostToDevice, st));
  if (e->dec_L == 0) {
    err = "this Whisper engine was created without decoder weights";
    return -1;
  }
  if (forward_graphed(e, 1, e->audio_stage.as<float>(), audio_ld, n_samples, B, enc_dev, dec_dev, st)) return -1;
  CK(cudaMemcpyAsync(pooled_host, enc_dev, enc_bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(dec_host, dec_dev, dec_bytes, cudaMemcpyDeviceToHost, st));
  if (forward_leave(e, st)) return -1;
  CK(cudaStreamSynchronize(st));
  return 0;
}

int32_t ssr_num_frames(const ssr_engine* e, int32_t n_samples) {
  if (!e) return -1;
  return e->d.family == SSR_WAVLM ? wavlm_frames(n_samples) : 1500;
}

int64_t ssr_launch_count(const ssr_engine* e) { return e ? e->launches : -1; }

int32_t ssr_wavlm_rel_bucket(int32_t rel) { return rel_bucket(rel); }

int ssr_tuning_set(const char* key, int32_t value) {
  if (!key) return -1;
  const std::string k(key);
  if (k == "attention_variant") {
    g_attention_variant = value & 3;
    return 0;
  }
  if (k == "pdl") {
    g_pdl = value != 0;
    return 0;
  }
  if (k == "attention_reverse") {
    g_attention_reverse = value != 0;
    return 0;
  }
  if (k == "ln_reverse") {
    g_ln_reverse = value != 0;
    return 0;
  }
  if (k == "attention_paired") {
    g_attention_paired = value != 0;
    return 0;
  }
  if (k == "attention_grouped") {
    g_attention_grouped = value != 0;
    return 0;
  }
  return -1;
}

int ssr_gemm_bf16(int32_t cuda_device, const void* A, int64_t lda, int64_t a_rows, const void* W, int32_t M,
                  int32_t N, int32_t K, const float* bias, int32_t act, const float* resid, float* out_f32,
                  void* out_bf16, int32_t simt, void* cuda_stream, char* errbuf, int32_t err_len) {
  std::string err;
  cudaSetDevice(cuda_device);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cuda_device);
  GemmOp op;
  op.A = reinterpret_cast<const bf16*>(A);
  op.lda = lda;
  op.a_rows = a_rows;
  op.W = reinterpret_cast<const bf16*>(W);
  op.M = M;
  op.N = N;
  op.K = K;
  op.a_mode = 0;
  op.a_cols = 0;
  op.epi = epi_plain(bias, act, resid, N, out_f32, N, reinterpret_cast<bf16*>(out_bf16), N);
  int rc = launch_gemm(op, as_stream(cuda_stream), simt != 0, sms, err);
  if (rc) copy_err(err, errbuf, err_len);
  return rc;
}

int ssr_gemm_bf16_pool(int32_t cuda_device, const void* A, int64_t lda, const void* W, int32_t M, int32_t N, int32_t K,
                       const float* bias, int32_t act, const float* resid, float* out_f32, int32_t slot,
                       const int32_t* lens_dev, int32_t B, float* part_dev, float* pooled_dev, int64_t pooled_ld,
                       int32_t simt, void* cuda_stream, char* errbuf, int32_t err_len) {
  std::string err;
  cudaSetDevice(cuda_device);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cuda_device);
  GemmOp op;
  op.A = reinterpret_cast<const bf16*>(A);
  op.lda = lda;
  op.a_rows = M;
  op.W = reinterpret_cast<const bf16*>(W);
  op.M = M;
  op.N = N;
  op.K = K;
  op.a_mode = 0;
  op.a_cols = 0;
  op.epi = epi_plain(bias, act, resid, N, out_f32, N, nullptr, 0);
  op.epi.pool_part = part_dev;
  op.epi.pool_slot = slot;
  op.epi.lens = lens_dev;
  int rc = launch_gemm(op, as_stream(cuda_stream), simt != 0, sms, err);
  if (!rc) rc = launch_pool_finalize(part_dev, B, slot, N, lens_dev, pooled_dev, pooled_ld, as_stream(cuda_stream), err);
  if (rc) copy_err(err, errbuf, err_len);
  return rc;
}

int ssr_layernorm(const float* in_f32, const void* in_bf16, int64_t rows, int32_t D, const float* gamma,
                  const float* beta, int32_t gelu, float* out_f32, void* out_bf16, void* cuda_stream, char* errbuf,
                  int32_t err_len) {
  std::string err;
  LayerNormArgs a;
  memset(&a, 0, sizeof(a));
  a.in_f32 = in_f32;
  a.in_bf16 = reinterpret_cast<const bf16*>(in_bf16);
  a.rows = rows;
  a.D = D;
  a.ld_in = D;
  a.gamma = gamma;
  a.beta = beta;
  a.eps = 1e-5f;
  a.gelu = gelu;
  a.out_f32 = out_f32;
  a.ld_out32 = D;
  a.out_bf16 = reinterpret_cast<bf16*>(out_bf16);
  a.ld_out16 = D;
  int rc = launch_layernorm(a, as_stream(cuda_stream), err);
  if (rc) copy_err(err, errbuf, err_len);
  return rc;
}

int ssr_attention(const void* qkv_bf16, void* out_bf16, int32_t B, int32_t slot, int32_t H, const int32_t* lens_dev,
                  const float* gate, const float* relbias, int32_t rel_stride, int32_t rel_center, int32_t impl,
                  void* cuda_stream, char* errbuf, int32_t err_len) {
  std::string err;
  AttentionArgs a;
  memset(&a, 0, sizeof(a));
  a.qkv = reinterpret_cast<const bf16*>(qkv_bf16);
  a.out = reinterpret_cast<bf16*>(out_bf16);
  a.B = B;
  a.slot = slot;
  a.H = H;
  a.D = H * 64;
  a.lens = lens_dev;
  a.gate = gate;
  a.relbias = relbias;
  a.rel_stride = rel_stride;
  a.rel_center = rel_center;
  int rc = impl == 0 ? launch_attention_tc(a, as_stream(cuda_stream), err) : launch_attention(a, as_stream(cuda_stream), err);
  if (rc) copy_err(err, errbuf, err_len);
  return rc;
}

int ssr_pool_mean(const float* x, int32_t B, int32_t slot, int32_t D, const int32_t* lens_dev, float* pooled,
                  int64_t pooled_ld, void* cuda_stream, char* errbuf, int32_t err_len) {
  std::string err;
  int rc = launch_pool_mean(x, B, slot, D, lens_dev, pooled, pooled_ld, as_stream(cuda_stream), err);
  if (rc) copy_err(err, errbuf, err_len);
  return rc;
}

const char* ssr_profile_fetch(ssr_engine* e) {
  if (!e) return nullptr;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  struct Agg {
    double ms = 0, flops = 0;
    long n = 0;
  };
  std::map<std::string, Agg> agg;
  for (ProfEntry& pe : e->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) {
      Agg& g = agg[pe.name];
      g.ms += ms;
      g.flops += pe.flops;
      g.n += 1;
    }
    cudaEventDestroy(pe.a);
    cudaEventDestroy(pe.b);
  }
  e->prof.clear();
  std::string js = "{";
  bool first = true;
  for (auto& kv : agg) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f, \"flops\": %.6e}", first ? "" : ", ",
             kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops);
    js += buf;
    first = false;
  }
  js += "}";
  e->prof_json = js;
  return e->prof_json.c_str();
}

int64_t ssr_debug_fetch(ssr_engine* e, const char* name, void* dst_host, int64_t dst_bytes, int64_t* dims,
                        int32_t* dtype) {
  if (!e || !name) return -1;
  auto it = e->dbg.find(name);
  if (it == e->dbg.end()) {
    e->err = std::string("no debug buffer named '") + name + "'";
    return -1;
  }
  const DbgBuf& b = it->second;
  if (dims) memcpy(dims, b.dims, sizeof(b.dims));
  if (dtype) *dtype = b.dtype;
  if (!dst_host) return b.bytes;
  if (dst_bytes < b.bytes) {
    e->err = "debug fetch: destination too small";
    return -2;
  }
  cudaSetDevice(e->device);
  cudaError_t ce = cudaDeviceSynchronize();
  if (ce == cudaSuccess) ce = cudaMemcpy(dst_host, b.p, (size_t)b.bytes, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) {
    e->err = std::string("debug fetch: ") + cudaGetErrorString(ce);
    return -3;
  }
  return b.bytes;
}

}  // extern "C"
