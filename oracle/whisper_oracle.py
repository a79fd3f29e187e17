"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the reference's Whisper-encoder embedding path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

The arithmetic lives in un-vendored `transformers` 5.5.0 / `torch` 2.11 (REF/whisper_embeddings_large.py:242-254 calls
them); restated here from:

    WhisperFeatureExtractor.__call__ / _torch_extract_fbank_features
                                 HF/models/whisper/feature_extraction_whisper.py:135-164, 189-342
    mel filterbank               HF/audio_utils.py:453-544 (restated in ssr_b200.melfilters, pinned in tests)
    conv stem + positions        HF/models/whisper/modeling_whisper.py:619-625 (sinusoids :55-64)
    encoder layer / attention    HF/models/whisper/modeling_whisper.py:284-357, 380-414
    hidden-state capture         HF/models/whisper/modeling_whisper.py:550-553 + HF/utils/output_capturing.py:101-113
    pooling                      REF/whisper_embeddings_large.py:272-283 (unmasked mean over all 1500 positions)

Parity pin: none exists in the reference ("parity unpinned" there); pinned here against the live HF modules and
tests/golden/*.npz (made by tools/make_golden.py from the reference's own extract function).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erf

N_FFT, HOP, N_SAMPLES, N_FRAMES = 400, 160, 480000, 3000


def gelu(x):
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * g + b


def sinusoids(length=1500, channels=1280, max_timescale=10000.0) -> np.ndarray:
    """HF modeling_whisper.py:55-64 (float32 arithmetic like torch's default dtype)."""
    inc = np.float32(math.log(max_timescale) / (channels // 2 - 1))
    inv = np.exp(-inc * np.arange(channels // 2, dtype=np.float32)).astype(np.float32)
    t = np.arange(length, dtype=np.float32)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def log_mel(audio, mel_filters: np.ndarray, dtype=np.float64) -> np.ndarray:
    """[80, 3000] float32 == WhisperFeatureExtractor()(audio).input_features[0] (torch path, dither 0)."""
    x = np.asarray(audio, dtype=np.float32).reshape(-1)[:N_SAMPLES]
    x = np.concatenate([x, np.zeros(N_SAMPLES - x.size, np.float32)]).astype(dtype)
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")  # torch.stft(center=True, pad_mode="reflect")
    frames = np.lib.stride_tricks.sliding_window_view(xp, N_FFT)[::HOP]  # [3001, 400]
    k = np.arange(N_FFT, dtype=np.float64)
    window = (0.5 - 0.5 * np.cos(2.0 * np.pi * k / N_FFT)).astype(dtype)  # torch.hann_window(400), periodic
    spec = np.fft.rfft(frames * window, axis=1)  # [3001, 201]
    power = (spec.real ** 2 + spec.imag ** 2)[:-1].T  # drop the last frame -> [201, 3000]
    mel = mel_filters.astype(dtype).T @ power
    log_spec = np.log10(np.maximum(mel, 1e-10))
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return ((log_spec + 4.0) / 4.0).astype(np.float32)


def conv1d_cl(x, w, stride, pad, dtype):
    T, Ci = x.shape
    Co, _, k = w.shape
    xp = np.concatenate([np.zeros((pad, Ci), dtype), x.astype(dtype), np.zeros((pad, Ci), dtype)], axis=0)
    T_out = (T + 2 * pad - k) // stride + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, k, axis=0)[::stride][:T_out]
    return win.reshape(T_out, Ci * k) @ w.reshape(Co, Ci * k).astype(dtype).T


class WhisperEncoderOracle:
    def __init__(self, state_dict: dict, d_model: int, layers: int, heads: int, dtype=np.float32):
        self.sd = {k: np.asarray(v) for k, v in state_dict.items()}
        self.D, self.L, self.H = d_model, layers, heads
        self.dt = dtype

    @classmethod
    def from_hf(cls, encoder, dtype=np.float32):
        c = encoder.config
        sd = {k: v.detach().cpu().numpy() for k, v in encoder.state_dict().items()}
        return cls(sd, c.d_model, c.encoder_layers, c.encoder_attention_heads, dtype)

    def w(self, name):
        return self.sd[name].astype(self.dt)

    def stem(self, mel):
        """mel [80, 3000] -> hidden_states[0] [1500, D]."""
        x = np.asarray(mel, self.dt).T  # channels-last [3000, 80]
        x = gelu(conv1d_cl(x, self.w("conv1.weight"), 1, 1, self.dt) + self.w("conv1.bias")).astype(self.dt)
        x = gelu(conv1d_cl(x, self.w("conv2.weight"), 2, 1, self.dt) + self.w("conv2.bias")).astype(self.dt)
        return x + self.w("embed_positions.weight")

    def layer(self, h, l):
        p = f"layers.{l}"
        T, D = h.shape
        H = self.H
        x = layer_norm(h, self.w(p + ".self_attn_layer_norm.weight"), self.w(p + ".self_attn_layer_norm.bias"))
        q = (x @ self.w(p + ".self_attn.q_proj.weight").T + self.w(p + ".self_attn.q_proj.bias")) * self.dt(
            (D // H) ** -0.5)
        k = x @ self.w(p + ".self_attn.k_proj.weight").T  # no bias
        v = x @ self.w(p + ".self_attn.v_proj.weight").T + self.w(p + ".self_attn.v_proj.bias")
        q, k, v = (a.reshape(T, H, -1).transpose(1, 0, 2) for a in (q, k, v))
        s = q @ k.transpose(0, 2, 1)
        s = s - s.max(axis=-1, keepdims=True)
        pr = np.exp(s)
        pr = pr / pr.sum(axis=-1, keepdims=True)
        o = (pr @ v).transpose(1, 0, 2).reshape(T, D)
        h = h + (o @ self.w(p + ".self_attn.out_proj.weight").T + self.w(p + ".self_attn.out_proj.bias"))
        x = layer_norm(h, self.w(p + ".final_layer_norm.weight"), self.w(p + ".final_layer_norm.bias"))
        y = gelu(x @ self.w(p + ".fc1.weight").T + self.w(p + ".fc1.bias")).astype(self.dt)
        return (h + (y @ self.w(p + ".fc2.weight").T + self.w(p + ".fc2.bias"))).astype(self.dt)

    def hidden_states(self, mel):
        """L+1 arrays: [stem output, layer 0 out, ..., layer L-2 out, layer_norm(layer L-1 out)]."""
        h = self.stem(mel).astype(self.dt)
        hs = [h]
        for l in range(self.L):
            h = self.layer(h, l)
            hs.append(h)
        hs[-1] = layer_norm(h, self.w("layer_norm.weight"), self.w("layer_norm.bias"))
        return hs

    def pooled(self, audio, mel_filters) -> np.ndarray:
        mel = log_mel(audio, mel_filters)
        return np.stack([h.mean(axis=0) for h in self.hidden_states(mel)]).astype(np.float32)


class WhisperDecoderTokenOracle:
    """The reference's decoder probe (REF/whisper_embeddings_large.py:257-262, 286-297): ONE decoder step with
    input_ids = [[0]] over the encoder's last_hidden_state; outputs are hidden_states[i].squeeze(1) (not pooled).
    HF/models/whisper/modeling_whisper.py:449-506 (layer), :691-796 (decoder). With a single query token the causal
    self-attention softmax is over one key, so its output is out_proj(v_proj(x)); q/k projections do not matter."""

    def __init__(self, state_dict: dict, d_model: int, layers: int, heads: int, dtype=np.float32):
        self.sd = {k: np.asarray(v) for k, v in state_dict.items()}
        self.D, self.L, self.H = d_model, layers, heads
        self.dt = dtype

    @classmethod
    def from_hf(cls, decoder, dtype=np.float32):
        c = decoder.config
        sd = {k: v.detach().cpu().numpy() for k, v in decoder.state_dict().items()}
        return cls(sd, c.d_model, c.decoder_layers, c.decoder_attention_heads, dtype)

    def w(self, name):
        return self.sd[name].astype(self.dt)

    def hidden_states(self, enc_last):
        """enc_last: [1500, D] (encoder last_hidden_state of one clip) -> L+1 arrays [D]."""
        enc = np.asarray(enc_last, self.dt)
        D, H = self.D, self.H
        h = self.w("embed_tokens.weight")[0] + self.w("embed_positions.weight")[0]
        hs = [h]
        for l in range(self.L):
            p = f"layers.{l}"
            x = layer_norm(h, self.w(p + ".self_attn_layer_norm.weight"), self.w(p + ".self_attn_layer_norm.bias"))
            v = x @ self.w(p + ".self_attn.v_proj.weight").T + self.w(p + ".self_attn.v_proj.bias")
            h = h + (v @ self.w(p + ".self_attn.out_proj.weight").T + self.w(p + ".self_attn.out_proj.bias"))
            x = layer_norm(h, self.w(p + ".encoder_attn_layer_norm.weight"), self.w(p + ".encoder_attn_layer_norm.bias"))
            q = (x @ self.w(p + ".encoder_attn.q_proj.weight").T + self.w(p + ".encoder_attn.q_proj.bias")) * self.dt(
                (D // H) ** -0.5)
            k = enc @ self.w(p + ".encoder_attn.k_proj.weight").T  # no bias
            v = enc @ self.w(p + ".encoder_attn.v_proj.weight").T + self.w(p + ".encoder_attn.v_proj.bias")
            qh = q.reshape(H, -1)                       # [H, 64]
            kh = k.reshape(-1, H, D // H).transpose(1, 0, 2)  # [H, T, 64]
            vh = v.reshape(-1, H, D // H).transpose(1, 0, 2)
            s = np.einsum("hd,htd->ht", qh, kh)
            s = s - s.max(axis=-1, keepdims=True)
            pr = np.exp(s)
            pr = pr / pr.sum(axis=-1, keepdims=True)
            o = np.einsum("ht,htd->hd", pr, vh).reshape(D)
            h = h + (o @ self.w(p + ".encoder_attn.out_proj.weight").T + self.w(p + ".encoder_attn.out_proj.bias"))
            x = layer_norm(h, self.w(p + ".final_layer_norm.weight"), self.w(p + ".final_layer_norm.bias"))
            y = gelu(x @ self.w(p + ".fc1.weight").T + self.w(p + ".fc1.bias")).astype(self.dt)
            h = (h + (y @ self.w(p + ".fc2.weight").T + self.w(p + ".fc2.bias"))).astype(self.dt)
            hs.append(h)
        hs[-1] = layer_norm(h, self.w("layer_norm.weight"), self.w("layer_norm.bias"))
        return hs
