/*
 * ssr_b200 — C ABI of the B200-native embedding-extraction hot path.
 *
 * Drop-in boundary for the reference's per-clip extraction calls
 *   extract_wavlm_embeddings            /root/reference/WavLM_embeddings.py:267-341
 *   extract_embeddings_from_audio_wavlm /root/reference/model_training_1.py:235-266
 *   extract_whisper_embeddings_fixed    /root/reference/whisper_embeddings_large.py:234-299 (encoder part)
 *   extract_embeddings_from_audio_whisper /root/reference/model_training_1.py:268-316       (encoder part)
 * i.e. for   feature_extractor(audio) -> model(..., output_hidden_states=True) -> torch.mean(h, dim=1)
 * (WavLM_embeddings.py:289-323, whisper_embeddings_large.py:242-254,272-283).
 *
 * Conventions
 *   - plain C types only; no exceptions cross the boundary; return 0 = OK, negative = error
 *     (text via ssr_last_error). The Python shim turns any non-zero return into "log + return None",
 *     which is the reference's error convention (WavLM_embeddings.py:329-341).
 *   - the caller owns every buffer; the engine owns packed weights and its workspace arena.
 *   - the device-buffer entry points (ssr_wavlm_pooled, ssr_whisper_enc_pooled, ssr_whisper_full, ssr_logmel) are
 *     asynchronous on the given CUDA stream (a cudaStream_t passed as void*; NULL = legacy default stream). They
 *     do not synchronise in steady state: the host n_samples array is copied into a pinned ring owned by the engine
 *     before the call returns (the caller may reuse it at once) and travels to the device from there. The one
 *     exception is workspace growth: the first call with a larger batch / longer clips than any before reallocates
 *     (cudaFree / cudaMalloc synchronise). *_host entry points copy in, run, copy out and synchronise the stream
 *     before returning.
 *   - one engine per (process, device), one host thread at a time (the reference is single-threaded). Calls on
 *     DIFFERENT streams are ordered by the engine (its workspace is shared): each forward waits, on the device,
 *     for the previous forward of the same engine whichever stream that ran on.
 *   - pooled output layout: float32 [B, L+1, D], layer-major per clip, so pooled[b, i, :] equals
 *     torch.mean(hidden_states[i], dim=1) of the reference for clip b, for every i in 0..L
 *     (hidden_states order: HF modeling_wavlm.py:412-439,488-516; modeling_whisper.py:550-553 + output_capturing).
 */
#ifndef SSR_B200_H_
#define SSR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssr_engine ssr_engine;

enum { SSR_WAVLM = 0, SSR_WHISPER_ENC = 1 };
enum { SSR_FEAT_NORM_GROUP = 0, SSR_FEAT_NORM_LAYER = 1 };

typedef struct ssr_model_desc {
  int32_t family;       /* SSR_WAVLM | SSR_WHISPER_ENC */
  int32_t hidden;       /* D: 768 / 1024 (WavLM), 1280 (Whisper-large) */
  int32_t layers;       /* L */
  int32_t heads;        /* H, head_dim must be 64 */
  int32_t ffn;          /* F */
  int32_t feat_norm;    /* WavLM: SSR_FEAT_NORM_GROUP (Base+) | SSR_FEAT_NORM_LAYER (Large) */
  int32_t stable_ln;    /* WavLM: config.do_stable_layer_norm */
  int32_t do_normalize; /* WavLM: Wav2Vec2FeatureExtractor.do_normalize (zero-mean / unit-variance per clip) */
  int32_t n_mels;       /* Whisper: 80 */
  int32_t reserved[7];
} ssr_model_desc;

/* One tensor of the HF state_dict: fp32, contiguous, host memory, named exactly as in model.state_dict().
 * Whisper additionally takes the feature extractor's filterbank as "mel_filters" ([201, n_mels]). */
typedef struct ssr_weight {
  const char* name;
  const float* data;
  int64_t numel;
} ssr_weight;

int ssr_create(const ssr_model_desc* desc, const ssr_weight* weights, int32_t n_weights, int32_t cuda_device,
               ssr_engine** out);
void ssr_destroy(ssr_engine* e);
/* e == NULL returns the message of the last failed ssr_create on this thread's process. */
const char* ssr_last_error(const ssr_engine* e);

/* Options: "simt_gemm" (bring-up cross-check GEMM), "fused_pool" (default 1), "snapshot_layer" (-1 = off),
 * "attn_simt" (1 = mma.sync attention cross-check kernel), "posconv_generic" (1 = positional conv through the generic
 * GEMM), "graphs" (default 1: the *_host entry points replay a captured CUDA graph when a small batch (B <= 16)
 * repeats the previous call's batch size, pitch and lengths — the reference's per-clip loop over equal-length clips),
 * "host_pipeline" (default 1: the *_host entry points move a batch of >= 64 clips host -> device in chunks on a copy
 * stream while the front end already runs on the chunks that have landed, and copy the pooled rows of
 * hidden_states[0 .. L-1] back while the last layer still computes), "host_chunks" (number of those chunks, 1..8,
 * default 4, sizes growing geometrically), "logmel_dense" (1 = the dense-DFT log-mel kernel,
 * a cross-check of the default folded-DFT one), "conv_ln_fused" (default 1),
 * "profile" (1 = bracket every kernel launch with CUDA events on the launching stream; read with ssr_profile_fetch).
 * Returns 0, or -1 for an unknown key. */
int ssr_set_option(ssr_engine* e, const char* key, int32_t value);

/* ---- hot path, device buffers --------------------------------------------------------------------------------
 * audio_dev: float32 [B, audio_ld] (clip b occupies audio_dev[b*audio_ld .. + n_samples[b]); the rest is ignored)
 * n_samples: HOST int32 [B]
 * pooled_dev: float32 [B, L+1, D] */
int ssr_wavlm_pooled(ssr_engine* e, const float* audio_dev, int64_t audio_ld, const int32_t* n_samples, int32_t B,
                     float* pooled_dev, void* cuda_stream);
int ssr_whisper_enc_pooled(ssr_engine* e, const float* audio_dev, int64_t audio_ld, const int32_t* n_samples,
                           int32_t B, float* pooled_dev, void* cuda_stream);
/* Whisper log-mel stage alone: mel_dev float32 [B, n_mels, 3000] (== WhisperFeatureExtractor input_features). */
int ssr_logmel(ssr_engine* e, const float* audio_dev, int64_t audio_ld, const int32_t* n_samples, int32_t B,
               float* mel_dev, void* cuda_stream);

/* ---- hot path, host buffers (the call the drop-in shim makes; H2D and D2H happen inside) -------------------- */
int ssr_wavlm_pooled_host(ssr_engine* e, const float* audio_host, int64_t audio_ld, const int32_t* n_samples,
                          int32_t B, float* pooled_host);
int ssr_whisper_enc_pooled_host(ssr_engine* e, const float* audio_host, int64_t audio_ld, const int32_t* n_samples,
                                int32_t B, float* pooled_host);

/* ---- Whisper encoder + decoder start-token probe -------------------------------------------
 * Available when ssr_create received the decoder tensors: desc.reserved[0] = decoder_layers (Ld),
 * desc.reserved[1] = decoder_ffn_dim, and weights "decoder.layers.{l}.*", "decoder.layer_norm.*" plus the two rows
 * "decoder.embed_tokens.weight[0]" and "decoder.embed_positions.weight[0]" ([D] each: the reference feeds
 * input_ids = [[0]], REF/whisper_embeddings_large.py:257-262).
 * dec: float32 [B, Ld+1, D] with dec[b, i, :] == decoder hidden_states[i].squeeze(1) (REF :286-297); pooled as above. */
int32_t ssr_decoder_layers(const ssr_engine* e);
int ssr_whisper_full(ssr_engine* e, const float* audio_dev, int64_t audio_ld, const int32_t* n_samples, int32_t B,
                     float* pooled_dev, float* dec_dev, void* cuda_stream);
int ssr_whisper_full_host(ssr_engine* e, const float* audio_host, int64_t audio_ld, const int32_t* n_samples,
                          int32_t B, float* pooled_host, float* dec_host);

/* Number of frames the model produces for a clip of n samples (WavLM conv arithmetic, HF modeling_wavlm.py:647-653;
 * Whisper: always 1500). */
int32_t ssr_num_frames(const ssr_engine* e, int32_t n_samples);
/* WavLM relative-position bucket of rel = key_index - query_index (the one piece of integer arithmetic on the path:
 * HF WavLMAttention._relative_positions_bucket, modeling_wavlm.py:252-271, num_buckets 320, max_distance 800).
 * Host function, needs no device; the engine builds its [H, 2R-1] relative-bias table (debug tap "relbias") from it. */
int32_t ssr_wavlm_rel_bucket(int32_t rel);
/* Process-wide kernel tuning knobs (A/B measurements, tools/attn_probe.py; engines with a captured CUDA graph keep
 * the variant they captured until the graph is dropped). Keys: "attention_variant" (bit 0: packed fp32 pair arithmetic
 * in the softmax; bit 1: a quarter of the exponentials on the FMA pipe; the default, 1, is the fastest measured),
 * "attention_paired" (1, default: clips of two query tiles are walked so that both tiles of a (clip, head) run at the
 * same time on neighbouring CTAs and K / V are read from HBM once; 0: query-tile-major order), "attention_grouped" (the
 * same for clips of three or more query tiles: the query tile is the fastest digit of the item index), "pdl" (1: the per-layer kernels
 * are launched with programmatic stream serialization so that a kernel's prologue overlaps its predecessor's tail;
 * 0, default: plain stream order — measured faster), "ln_reverse" / "attention_reverse" (1, default: LayerNorm rows /
 * attention clips are visited from the end, where the producing kernel's most recent output still sits in L2). Returns 0, or -1 for an unknown key. */
int ssr_tuning_set(const char* key, int32_t value);
/* Count of this library's kernel launches since creation (bench.py's gpu_launches). */
int64_t ssr_launch_count(const ssr_engine* e);
/* With option "profile" = 1: synchronises, then returns a JSON object
 *   {"<kernel group>": {"launches": n, "ms": device-event milliseconds, "flops": algorithmic FLOPs}, ...}
 * aggregated over every launch since the previous fetch (string owned by the engine, valid until the next call). */
const char* ssr_profile_fetch(ssr_engine* e);

/* ---- stage-level entry points (kernel parity tests; all pointers are device pointers) ------------------------ */
/* C[M,N] = act(A[M,K] * W[N,K]^T + bias) + resid ; A row r starts at A + r*lda (bf16), rows >= a_rows read zero.
 * act: 0 none, 1 erf-GELU. out_f32 / out_bf16 / bias / resid may be NULL. simt != 0 selects the debug kernel. */
int ssr_gemm_bf16(int32_t cuda_device, const void* A, int64_t lda, int64_t a_rows, const void* W, int32_t M,
                  int32_t N, int32_t K, const float* bias, int32_t act, const float* resid, float* out_f32,
                  void* out_bf16, int32_t simt, void* cuda_stream, char* err, int32_t err_len);
/* As above plus fused time mean-pool: rows are clips of `slot` rows of which lens_dev[b] are live;
 * pooled[b, :] (stride pooled_ld) = mean over live rows of the fp32 result. part_dev: scratch
 * float32 [ceil(M/32) * 2 * N]. */
int ssr_gemm_bf16_pool(int32_t cuda_device, const void* A, int64_t lda, const void* W, int32_t M, int32_t N, int32_t K,
                       const float* bias, int32_t act, const float* resid, float* out_f32, int32_t slot,
                       const int32_t* lens_dev, int32_t B, float* part_dev, float* pooled_dev, int64_t pooled_ld,
                       int32_t simt, void* cuda_stream, char* err, int32_t err_len);
/* LayerNorm over the last dim (eps 1e-5), optional GELU; exactly one of in_f32 / in_bf16. */
int ssr_layernorm(const float* in_f32, const void* in_bf16, int64_t rows, int32_t D, const float* gamma,
                  const float* beta, int32_t gelu, float* out_f32, void* out_bf16, void* cuda_stream, char* err,
                  int32_t err_len);
/* Self-attention over fused qkv [B*slot, 3*D] bf16 (q pre-scaled), optional WavLM gated relative bias
 * (relbias [H, rel_stride] must cover relative positions +-roundup(slot, 128) around rel_center).
 * impl: 0 = tcgen05/TMEM kernel (the product path), 1 = mma.sync cross-check kernel. */
int ssr_attention(const void* qkv_bf16, void* out_bf16, int32_t B, int32_t slot, int32_t H, const int32_t* lens_dev,
                  const float* gate, const float* relbias, int32_t rel_stride, int32_t rel_center, int32_t impl,
                  void* cuda_stream, char* err, int32_t err_len);
/* pooled[b, :] = mean over t < lens[b] of x[b*slot + t, :]. */
int ssr_pool_mean(const float* x, int32_t B, int32_t slot, int32_t D, const int32_t* lens_dev, float* pooled,
                  int64_t pooled_ld, void* cuda_stream, char* err, int32_t err_len);

/* ---- waveform augmentation, batched on the device. Engine-independent.
 * Replaces augment_audio of /root/reference/model_training_1.py:166-213 and model_training_01.py:140-192 (kinds
 * speed / noise / volume / pitch / none; the decision which kind and which factor stays on the host, drawn exactly like the
 * reference draws it, see stuttering-speech-representation_b200/augment.py). The speed kind is torchaudio's
 * Resample(sr -> new_rate) followed by Resample(new_rate -> sr) (sinc_interp_hann, width 6, rolloff 0.99); every
 * kind ends with clamp(-1, 1) (model_training_1.py:204). The pitch kind is torchaudio's PitchShift(sr, n_steps):
 * STFT 512/128 -> phase vocoder -> inverse STFT -> float32-kernel resample -> crop / pad to the input length
 * (model_training_01.py:174-178); it needs more than 256 samples. */
enum { SSR_AUG_NONE = 0, SSR_AUG_SPEED = 1, SSR_AUG_NOISE = 2, SSR_AUG_VOLUME = 3, SSR_AUG_PITCH = 4 };

typedef struct ssr_aug_op {
  int32_t kind;     /* SSR_AUG_* */
  int32_t new_rate; /* speed: int(sample_rate * speed_factor), model_training_1.py:185; pitch: n_steps */
  float factor;     /* noise: noise_factor; volume: volume_factor */
  int32_t reserved;
  uint64_t seed;    /* noise without caller-supplied normals: key of the counter-based generator */
} ssr_aug_op;

/* torchaudio's output length of one Resample(orig_rate -> new_rate) over n samples. */
int32_t ssr_resample_length(int32_t n, int32_t orig_rate, int32_t new_rate);
/* Output length of one augmentation op over n samples (speed: the round trip; else n). */
int32_t ssr_augment_out_length(const ssr_aug_op* op, int32_t n, int32_t sample_rate);
/* Bytes of device scratch ssr_augment needs for this batch. */
int64_t ssr_augment_work_bytes(const int32_t* n_in, int32_t batch, const ssr_aug_op* ops, int32_t sample_rate);
/* audio_dev: float32 [batch, in_stride] on the device, clip b has n_in[b] (host array) samples. ops: host array.
 * noise_dev (optional): float32 [batch, noise_stride] standard normals, used as-is by the noise kind (replaying the
 * reference's torch.randn_like stream gives bit-identical output); NULL = generate on the device (Philox4x32-10
 * keyed by (op.seed, sample): give every clip its own seed). out_dev: float32 [batch, out_stride]; samples beyond n_out[b] are zero-filled.
 * n_out: host array, filled before return. Asynchronous on the stream. */
int ssr_augment(const float* audio_dev, int64_t in_stride, const int32_t* n_in, int32_t batch, const ssr_aug_op* ops,
                int32_t sample_rate, const float* noise_dev, int64_t noise_stride, void* work_dev, int64_t work_bytes,
                float* out_dev, int64_t out_stride, int32_t* n_out, void* cuda_stream, char* err, int32_t err_len);

/* ---- classifier head on the pooled embeddings, data-parallel training. Engine-independent and
 * stateless: every buffer is the caller's (device pointers unless noted). New component; what it keeps from the
 * reference is the Pipeline([StandardScaler, classifier]) / class_weight='balanced' semantics of
 * /root/reference/model_training_1.py:576-589, :658-680. Model: Linear(D->H) + ReLU + Linear(H->C), class-weighted
 * softmax cross-entropy, Adam. Flat float32 parameter layout: W1[H,D] | b1[H] | W2[C,H] | b2[C]. */
int64_t ssr_head_param_count(int32_t D, int32_t H, int32_t C); /* P, or -1 (C <= 32) */
int64_t ssr_head_work_bytes(int64_t n, int32_t D, int32_t H, int32_t C);
/* StandardScaler statistics of the local rows, float64 [D]: mean_dev == NULL -> column sums of X;
 * else column sums of (x - mean)^2. (All-reduce them across ranks, divide by the global row count.) */
int ssr_head_scaler_stats(const float* X_dev, int64_t n, int32_t D, int64_t ld, const double* mean_dev,
                          double* out_dev, void* cuda_stream, char* err, int32_t err_len);
/* Forward + backward over n local rows (rows_dev: optional int32 [n] row gather into X_dev / y_dev; mean/inv_std:
 * optional fused scaler, float32 [D]). grad_dev: float32 [P + 2], overwritten with the UNNORMALISED sums
 *   sum_i w_i * dloss_i/dparam | sum_i w_i * loss_i | sum_i w_i      (w_i = class_w_dev[y_i], or 1)
 * so that one all-reduce(sum) of the buffer gives the global gradient and its normaliser. */
int ssr_head_grad(const float* X_dev, const int32_t* y_dev, const int32_t* rows_dev, int64_t n, int32_t D, int32_t H,
                  int32_t C, const float* mean_dev, const float* inv_std_dev, const float* params_dev,
                  const float* class_w_dev, float* grad_dev, void* work_dev, int64_t work_bytes, void* cuda_stream,
                  char* err, int32_t err_len);
/* Adam step (torch.optim.Adam semantics, L2 weight decay) on g = grad[0..P) / grad[P+1]; step counts from 1. */
int ssr_head_adam(float* params_dev, const float* grad_dev, float* m_dev, float* v_dev, int64_t P, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int32_t step, void* cuda_stream, char* err,
                  int32_t err_len);
/* pred_dev int32 [n] (argmax, lowest index on ties) and/or proba_dev float32 [n, C]; work: n * H floats. */
int ssr_head_predict(const float* X_dev, int64_t n, int32_t D, int32_t H, int32_t C, const float* mean_dev,
                     const float* inv_std_dev, const float* params_dev, int32_t* pred_dev, float* proba_dev,
                     void* work_dev, int64_t work_bytes, void* cuda_stream, char* err, int32_t err_len);

/* ---- debug taps: copy a named internal buffer of the last run to host (synchronises). Returns bytes copied,
 * or a negative error. With dst == NULL returns the buffer's size in bytes. `dims` (optional, 4 entries) gets the
 * logical shape, `dtype` (optional) 0 = float32, 1 = bfloat16. */
int64_t ssr_debug_fetch(ssr_engine* e, const char* name, void* dst_host, int64_t dst_bytes, int64_t* dims,
                        int32_t* dtype);

#ifdef __cplusplus
}
#endif
#endif /* SSR_B200_H_ */
