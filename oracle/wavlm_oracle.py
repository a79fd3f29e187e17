"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the reference's WavLM embedding-extraction path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the product package (ssr_b200) never does.

The reference (REF = /root/reference) holds no arithmetic of its own for this path: REF/WavLM_embeddings.py:289-323
calls HuggingFace `transformers` (not vendored, unpinned by the reference; 5.5.0 in this image) and `torch` 2.11.
This file restates that third-party algorithm in plain numpy, function by function, citing the HF lines it follows
(HF = site-packages/transformers):

    zero_mean_unit_var_norm      HF/models/wav2vec2/feature_extraction_wav2vec2.py:77-97
    feature encoder              HF/models/wavlm/modeling_wavlm.py:682-789
    feature projection           HF/models/wavlm/modeling_wavlm.py:93-105
    positional conv embedding    HF/models/wavlm/modeling_wavlm.py:48-90   (weight-norm: torch parametrizations)
    relative position buckets    HF/models/wavlm/modeling_wavlm.py:243-271
    gated relative-position bias HF/models/wavlm/modeling_wavlm.py:147-186
    attention                    torch F.multi_head_attention_forward (separate q/k/v weights, additive float mask)
    encoder layers / stacks      HF/models/wavlm/modeling_wavlm.py:298-522
    pooling                      REF/WavLM_embeddings.py:315-325  (torch.mean over time, per selected layer)

Parity pin: the reference has NO tests / golden vectors for this path ("parity unpinned" by the reference itself,
SURVEY.md section 4). The pin used instead is (a) the live HF modules and (b) tests/golden/*.npz, produced by
tools/make_golden.py by importing the reference's own extract_* functions from /root/reference in the build
container; tests/test_oracle_cpu.py checks this restatement against both.

`dtype` selects the arithmetic: float32 mimics the reference, float64 gives a tighter ground truth.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erf

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)


def num_frames(n_samples: int) -> int:
    """HF modeling_wavlm.py:647-653 (_get_feat_extract_output_lengths)."""
    n = int(n_samples)
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        if n < k:
            return 0
        n = (n - k) // s + 1
    return n


def gelu(x):
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * g + b


def zero_mean_unit_var_norm(x):
    """Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm (single un-padded clip): numpy mean / population var."""
    x = np.asarray(x, dtype=np.float32)
    return ((x - x.mean()) / np.sqrt(x.var() + 1e-7)).astype(np.float32)


def conv1d_cl(x, w, stride, dtype):
    """x: [T, Ci] channels-last, w: [Co, Ci, k] (torch layout), no padding -> [T_out, Co]."""
    T, Ci = x.shape
    Co, _, k = w.shape
    T_out = (T - k) // stride + 1
    win = np.lib.stride_tricks.sliding_window_view(x, k, axis=0)[::stride][:T_out]  # [T_out, Ci, k]
    return (win.reshape(T_out, Ci * k).astype(dtype) @ w.reshape(Co, Ci * k).astype(dtype).T).astype(dtype)


def rel_bucket(rel: np.ndarray) -> np.ndarray:
    """_relative_positions_bucket with num_buckets=320, max_distance=800; float32 log as in the reference."""
    rel = np.asarray(rel, dtype=np.int64)
    nb, max_exact = 160, 80
    out = (rel > 0).astype(np.int64) * nb
    a = np.abs(rel)
    with np.errstate(divide="ignore"):
        large = np.log(a.astype(np.float32) / np.float32(max_exact))
    large = large / np.float32(math.log(800 / max_exact))
    large = large * np.float32(nb - max_exact)
    large = (np.float32(max_exact) + large)
    large = np.where(np.isfinite(large), large, 0).astype(np.int64)
    large = np.minimum(large, nb - 1)
    return out + np.where(a < max_exact, a, large)


def position_bias(T: int, rel_attn_embed: np.ndarray) -> np.ndarray:
    """compute_bias: [H, T, T] with entry [h, i, j] = embed[bucket(j - i), h]."""
    i = np.arange(T)
    b = rel_bucket(i[None, :] - i[:, None])
    return np.transpose(rel_attn_embed[b], (2, 0, 1))


class WavLMOracle:
    def __init__(self, state_dict: dict, hidden: int, layers: int, heads: int, ffn: int, feat_norm: str,
                 stable_ln: bool, dtype=np.float32):
        self.sd = {k: np.asarray(v) for k, v in state_dict.items()}
        self.D, self.L, self.H, self.F = hidden, layers, heads, ffn
        self.feat_norm, self.stable = feat_norm, bool(stable_ln)
        self.dt = dtype

    @classmethod
    def from_hf(cls, model, dtype=np.float32):
        c = model.config
        sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
        return cls(sd, c.hidden_size, c.num_hidden_layers, c.num_attention_heads, c.intermediate_size,
                   c.feat_extract_norm, c.do_stable_layer_norm, dtype)

    def w(self, name):
        return self.sd[name].astype(self.dt)

    # ---- front end ----
    def feature_encoder(self, audio):
        x = np.asarray(audio, dtype=self.dt)[:, None]  # [T, 1]
        for i, (k, s) in enumerate(zip(CONV_KERNEL, CONV_STRIDE)):
            p = f"feature_extractor.conv_layers.{i}"
            x = conv1d_cl(x, self.w(p + ".conv.weight"), s, self.dt)
            if self.feat_norm == "layer":
                x = layer_norm(x, self.w(p + ".layer_norm.weight"), self.w(p + ".layer_norm.bias"))
            elif i == 0:  # GroupNorm(512 groups, 512 channels): per-channel statistics over time
                mu = x.mean(axis=0, keepdims=True)
                var = ((x - mu) ** 2).mean(axis=0, keepdims=True)
                x = (x - mu) / np.sqrt(var + 1e-5) * self.w(p + ".layer_norm.weight") + self.w(p + ".layer_norm.bias")
            x = gelu(x).astype(self.dt)
        return x  # [T', 512]

    def feature_projection(self, feats):
        x = layer_norm(feats, self.w("feature_projection.layer_norm.weight"),
                       self.w("feature_projection.layer_norm.bias"))
        return x @ self.w("feature_projection.projection.weight").T + self.w("feature_projection.projection.bias")

    def pos_conv(self, h):
        g = self.w("encoder.pos_conv_embed.conv.parametrizations.weight.original0")  # [1, 1, 128]
        v = self.w("encoder.pos_conv_embed.conv.parametrizations.weight.original1")  # [D, D/16, 128]
        w = g * v / np.sqrt((v.astype(np.float64) ** 2).sum(axis=(0, 1), keepdims=True)).astype(self.dt)
        bias = self.w("encoder.pos_conv_embed.conv.bias")
        T, D = h.shape
        gw = D // 16
        xp = np.concatenate([np.zeros((64, D), self.dt), h, np.zeros((64, D), self.dt)], axis=0)
        out = np.empty((T + 1, D), self.dt)
        for grp in range(16):
            sl = slice(grp * gw, (grp + 1) * gw)
            out[:, sl] = conv1d_cl(xp[:, sl], w[sl], 1, self.dt)
        out = out[:T] + bias  # SamePad drops the last frame (even kernel)
        return gelu(out).astype(self.dt)

    # ---- transformer ----
    def attention(self, x, l, bias):
        p = f"encoder.layers.{l}.attention"
        T, D = x.shape
        H = self.H
        xh = x.reshape(T, H, D // H)
        proj = xh @ self.w(p + ".gru_rel_pos_linear.weight").T + self.w(p + ".gru_rel_pos_linear.bias")  # [T, H, 8]
        proj = proj.reshape(T, H, 2, 4).sum(-1)
        gate = 1.0 / (1.0 + np.exp(-proj))
        ga, gb = gate[..., 0], gate[..., 1]
        const = self.w(p + ".gru_rel_pos_const").reshape(1, H)
        gate_out = ga * (gb * const - 1.0) + 2.0  # [T, H]
        gated = gate_out.T[:, :, None] * bias  # [H, T, T]
        q = (x @ self.w(p + ".q_proj.weight").T + self.w(p + ".q_proj.bias")).reshape(T, H, -1).transpose(1, 0, 2)
        k = (x @ self.w(p + ".k_proj.weight").T + self.w(p + ".k_proj.bias")).reshape(T, H, -1).transpose(1, 0, 2)
        v = (x @ self.w(p + ".v_proj.weight").T + self.w(p + ".v_proj.bias")).reshape(T, H, -1).transpose(1, 0, 2)
        s = (q * self.dt((D // H) ** -0.5)) @ k.transpose(0, 2, 1) + gated
        s = s - s.max(axis=-1, keepdims=True)
        pr = np.exp(s)
        pr = pr / pr.sum(axis=-1, keepdims=True)
        o = (pr @ v).transpose(1, 0, 2).reshape(T, D)
        return o @ self.w(p + ".out_proj.weight").T + self.w(p + ".out_proj.bias")

    def ffn(self, x, l):
        p = f"encoder.layers.{l}.feed_forward"
        y = gelu(x @ self.w(p + ".intermediate_dense.weight").T + self.w(p + ".intermediate_dense.bias"))
        return y.astype(self.dt) @ self.w(p + ".output_dense.weight").T + self.w(p + ".output_dense.bias")

    def hidden_states(self, audio, do_normalize: bool, return_stages: bool = False):
        """List of L+1 arrays [T, D] == WavLMModel(..., output_hidden_states=True).hidden_states for one clip."""
        x = zero_mean_unit_var_norm(audio) if do_normalize else np.asarray(audio, np.float32)
        stages = {}
        feats = self.feature_encoder(x)
        stages["conv6"] = feats
        h = self.feature_projection(feats).astype(self.dt)
        stages["feat"] = h
        h = h + self.pos_conv(h)
        if not self.stable:
            h = layer_norm(h, self.w("encoder.layer_norm.weight"), self.w("encoder.layer_norm.bias"))
        T = h.shape[0]
        bias = position_bias(T, self.w("encoder.layers.0.attention.rel_attn_embed.weight"))
        hs = []
        for l in range(self.L):
            hs.append(h)
            p = f"encoder.layers.{l}"
            if self.stable:
                h = h + self.attention(layer_norm(h, self.w(p + ".layer_norm.weight"), self.w(p + ".layer_norm.bias")),
                                       l, bias)
                h = h + self.ffn(layer_norm(h, self.w(p + ".final_layer_norm.weight"),
                                            self.w(p + ".final_layer_norm.bias")), l)
            else:
                h = layer_norm(h + self.attention(h, l, bias), self.w(p + ".layer_norm.weight"),
                               self.w(p + ".layer_norm.bias"))
                h = layer_norm(h + self.ffn(h, l), self.w(p + ".final_layer_norm.weight"),
                               self.w(p + ".final_layer_norm.bias"))
            h = h.astype(self.dt)
        if self.stable:
            h = layer_norm(h, self.w("encoder.layer_norm.weight"), self.w("encoder.layer_norm.bias"))
        hs.append(h)
        return (hs, stages) if return_stages else hs

    def pooled(self, audio, do_normalize: bool) -> np.ndarray:
        """[L+1, D] float32: row i == torch.mean(hidden_states[i], dim=1) (REF/WavLM_embeddings.py:321)."""
        return np.stack([h.mean(axis=0) for h in self.hidden_states(audio, do_normalize)]).astype(np.float32)


def select_layers(pooled: np.ndarray, layer_indices, prefix="layer_") -> dict:
    """The reference's selection loop (REF/WavLM_embeddings.py:315-325): out-of-range indices are skipped."""
    return {f"{prefix}{i}": pooled[i].astype(np.float32).flatten() for i in layer_indices if i < pooled.shape[0]}
