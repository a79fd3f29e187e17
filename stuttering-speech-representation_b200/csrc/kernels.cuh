// Launcher declarations for the non-GEMM kernels (internal).
#pragma once
#include "common.cuh"

namespace ssr {

struct LayerNormArgs {
  const float* in_f32;  // exactly one of in_f32 / in_bf16
  const bf16* in_bf16;
  long long rows;
  int D;
  long long ld_in;
  const float* gamma;
  const float* beta;
  float eps;
  int gelu;
  float* out_f32;
  long long ld_out32;
  bf16* out_bf16;
  long long ld_out16;
  int reverse;  // 1: CTAs walk the rows from the end (set by launch_layernorm from g_ln_reverse)
  // optional WavLM gate (needs head_dim 64)
  float* gate_out;  // [rows, n_heads]
  const float* gate_wa;
  const float* gate_wb;
  float gate_ba, gate_bb;
  const float* gate_const;  // [n_heads]
  int n_heads;
};
int launch_layernorm(const LayerNormArgs& a, cudaStream_t st, std::string& err);

int launch_pool_mean(const float* x, int B, int slot, int D, const int* lens, float* out, long long out_stride,
                     cudaStream_t st, std::string& err);
int launch_pool_finalize(const float* part, int B, int slot, int D, const int* lens, float* out, long long out_stride,
                         cudaStream_t st, std::string& err, int n_layers = 1, long long part_layer_stride = 0,
                         long long out_layer_stride = 0);
int launch_posconv_pack(const float* x, int B, int slot, int D, const int* lens, bf16* xp, int pslot, cudaStream_t st,
                        std::string& err);
int launch_posconv_finish(const float* conv, int pslot, const float* bias, const float* x, int B, int slot, int D,
                          const int* lens, float* h, cudaStream_t st, std::string& err);

// Multi-head self-attention over fused qkv [B*slot, 3*D] (bf16; q already scaled), head_dim 64.
// Keys j >= lens[b] are masked. If gate != nullptr the WavLM gated relative-position bias is added:
//   score[i, j] += gate[b*slot + i, h] * relbias[h * rel_stride + (j - i) + rel_center]
struct AttentionArgs {
  const bf16* qkv;
  bf16* out;  // [B*slot, D]
  int B, slot, H, D;
  const int* lens;
  const float* gate;
  const float* relbias;
  int rel_stride, rel_center;
};
int launch_attention(const AttentionArgs& a, cudaStream_t st, std::string& err);      // mma.sync (debug / tiny T)
int launch_attention_tc(const AttentionArgs& a, cudaStream_t st, std::string& err);   // tcgen05 / TMEM / TMA
extern int g_attention_variant;  // attention_tc.cu kernel variant (process-wide tuning knob)
extern int g_attention_paired;   // 1: paired item order for two-tile clips
extern int g_attention_grouped;  // 1: grouped item order for clips of three or more query tiles
extern int g_attention_reverse;  // 1: clips are walked from the last one (freshest qkv rows first)
extern int g_ln_reverse;         // 1: LayerNorm CTAs walk the rows from the end

// Whisper decoder single-token cross-attention (decoder.cu). All token-level tensors have one row per clip.
struct DecCrossArgs {
  const float* q;     // [B, D] fp32 : (Wq x + bq) * head_dim^-0.5
  const bf16* wk;     // [D, D] encoder_attn.k_proj.weight (no bias)
  const bf16* wv;     // [D, D] encoder_attn.v_proj.weight
  const float* bv;    // [D]
  const bf16* enc;    // [B * T, D] encoder last_hidden_state (bf16 copy)
  float* qp;          // scratch [B, H, D]
  float* ml;          // scratch [B, S, H, 2]  running max / sum of each row split (S = dec_cross_splits)
  float* ctx_part;    // scratch [B, S, H, D]
  bf16* out;          // [B, D] attention output before out_proj (bf16: next GEMM's A operand)
  int B, T, D, H;
};
int launch_dec_cross_attention(const DecCrossArgs& a, cudaStream_t st, std::string& err);
int dec_cross_splits(int B, int T, int num_sms);  // row splits per clip (scratch sizing)
int launch_bcast_rows(const float* v, float* out, int B, int D, long long ld, cudaStream_t st, std::string& err);
// Token-level Linear for a handful of rows (M <= 16): one warp per output column, weights streamed once.
bool dec_gemv_applicable(int M, int K);
int launch_dec_gemv(const bf16* A, int M, int K, const bf16* W, int N, const EpiParams& ep, cudaStream_t st,
                    std::string& err);

// WavLM waveform statistics + first conv layer (C_in = 1, k = 10, stride 5) fused with its normalisation + GELU.
struct Conv0Args {
  const float* audio;  // [B, audio_ld]
  long long audio_ld;
  const int* n_samples;  // [B] device
  int B;
  int do_normalize;  // Wav2Vec2FeatureExtractor zero-mean / unit-variance
  float* stats;      // [B, 2] mean, rstd (scratch)
  const float* w;    // [512, 10]
  const float* wstat;  // [110]: mean_c w[c][k] (10), then mean_c w[c][k] w[c][k'] as [10][10] with the upper triangle
                       // holding doubled off-diagonals (LayerNorm mode: row statistics as a quadratic form in x)
  const float* gamma;
  const float* beta;  // [512]
  int mode;           // 0: LayerNorm over channels (Large); 1: GroupNorm over time (Base+)
  double* gn_acc;     // [B, 512, 2] scratch for mode 1
  bf16* out;          // [B, slot0, 512]
  int slot0;          // rows per clip in out
};
int launch_wavlm_conv0(const Conv0Args& a, cudaStream_t st, std::string& err);

// Conv1d (implicit GEMM, 512 output channels) + LayerNorm(512) + GELU in one kernel (gemm_ln.cu).
int launch_gemm_ln(const bf16* A, long long lda, long long a_rows, const bf16* W, int M, int K, const float* gamma,
                   const float* beta, float eps, bf16* out, cudaStream_t st, int num_sms, std::string& err);

// Whisper log-mel front end.
struct LogMelArgs {
  const float* audio;
  long long audio_ld;
  const int* n_samples;  // device [B]
  int B;
  int max_samples;       // max over the batch (host-known bound on live frames)
  const float* twiddle;  // [400, 416] fp32 : col 2f = cos, 2f+1 = -sin (f < 201), zero padded
  const float* melw;     // [80, 201] fp32 dense filterbank
  const int* mel_lo;     // [80] first non-zero bin
  const int* mel_hi;     // [80] last non-zero bin
  // folded-DFT kernel (default): window-folded tables [4][104][104] fp32 = {cos, sin} x {even k, odd k}, the
  // filterbank split into bf16 hi / lo parts [80][208] and the non-zero 16-bin k-step range of every 8-mel tile
  const float* dft_tab;
  const bf16* melw_hi;
  const bf16* melw_lo;
  const int* mel_band;   // [10][2] first / one-past-last 16-bin step
  int dense;             // 1: the round-1 dense-DFT kernel (cross-check)
  float* logspec;        // scratch [B, 3000, 80] fp32 (log10 mel, un-normalised)
  unsigned int* gmax;    // scratch [B] ordered-uint encoding of the per-clip max
  float* mel_out;        // optional [B, 80, 3000] fp32 (HF layout)
  bf16* conv_in;         // optional [B, 3002, 80] bf16 channels-last with one zero row each side
};
int launch_logmel(const LogMelArgs& a, cudaStream_t st, std::string& err);

}  // namespace ssr
