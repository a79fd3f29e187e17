"""Host-side mirror of the reference's model objects for the embedding-extraction hot path.

`WavLMEngine` / `WhisperEncoderEngine` wrap one `ssr_engine*` of the C ABI (include/ssr_b200.h). They are built
from the very objects the reference scripts hold — the HF model (its `state_dict()` + `config`) and the HF
feature extractor (REF/WavLM_embeddings.py:482-483, REF/whisper_embeddings_large.py:431-438) — and return the
per-layer time-mean-pooled tensor `[B, L+1, D]` whose row `i` equals `torch.mean(hidden_states[i], dim=1)`
(REF/WavLM_embeddings.py:321, REF/whisper_embeddings_large.py:278).

PyTorch is used only for device memory, streams and pinned host buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np
import torch

from . import _lib
from .melfilters import whisper_mel_filters


class SsrError(RuntimeError):
    pass


def _as_f32_clip(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    a = np.asarray(a, dtype=np.float32).reshape(-1)
    return np.ascontiguousarray(a)


def row_pitch(t: torch.Tensor) -> int:
    """Elements between consecutive rows of a 2-D tensor. A single-row view (e.g. `x[None]`) may carry stride 0 or any
    other value in its size-1 dimension; the C ABI wants a pitch that covers the row."""
    return int(t.stride(0)) if t.shape[0] > 1 else max(int(t.stride(0)), int(t.shape[1]))


class _EngineBase:
    family = -1

    def __init__(self, desc: _lib.ModelDesc, tensors: dict[str, np.ndarray], device: int = 0):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise SsrError("ssr_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = int(device)
        self.desc = desc
        self.hidden, self.layers = int(desc.hidden), int(desc.layers)
        keep = []  # keep the host arrays alive during ssr_create
        arr = (_lib.Weight * len(tensors))()
        for i, (name, t) in enumerate(tensors.items()):
            t = np.ascontiguousarray(t, dtype=np.float32)
            keep.append(t)
            arr[i].name = name.encode()
            arr[i].data = t.ctypes.data
            arr[i].numel = t.size
        h = C.c_void_p()
        rc = self._lib.ssr_create(C.byref(desc), arr, len(tensors), self.device, C.byref(h))
        if rc != 0:
            raise SsrError(f"ssr_create failed ({rc}): {self._lib.ssr_last_error(None).decode()}")
        self._h = h
        self._pin_in = None
        self._pin_out = None

    # ------------------------------------------------------------------ plumbing
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.ssr_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self) -> str:
        return self._lib.ssr_last_error(self._h).decode()

    def set_option(self, key: str, value: int):
        if self._lib.ssr_set_option(self._h, key.encode(), int(value)) != 0:
            raise SsrError(self._err())

    def num_frames(self, n_samples: int) -> int:
        return int(self._lib.ssr_num_frames(self._h, int(n_samples)))

    @property
    def launch_count(self) -> int:
        return int(self._lib.ssr_launch_count(self._h))

    def profile_fetch(self) -> dict:
        """Per-kernel-group device-event times since the last fetch (needs set_option('profile', 1))."""
        import json

        s = self._lib.ssr_profile_fetch(self._h)
        return json.loads(s.decode()) if s else {}

    def debug_fetch(self, name: str) -> np.ndarray:
        """Copy a named internal buffer of the last run (bf16 buffers are returned as float32)."""
        dims = (C.c_int64 * 4)()
        dt = C.c_int32()
        n = self._lib.ssr_debug_fetch(self._h, name.encode(), None, 0, dims, C.byref(dt))
        if n < 0:
            raise SsrError(self._err())
        raw = np.empty(n, dtype=np.uint8)
        if self._lib.ssr_debug_fetch(self._h, name.encode(), raw.ctypes.data, n, dims, C.byref(dt)) < 0:
            raise SsrError(self._err())
        shape = [int(d) for d in dims]
        while len(shape) > 1 and shape[-1] == 1:
            shape.pop()
        if dt.value == 0:
            return raw.view(np.float32).reshape(shape)
        u16 = raw.view(np.uint16).astype(np.uint32) << 16
        return u16.view(np.float32).reshape(shape)

    # ------------------------------------------------------------------ batching helpers
    def _stage(self, clips: Sequence) -> tuple[torch.Tensor, np.ndarray]:
        """Pack ragged clips into a pinned [B, ld] float32 buffer (ld = max length rounded up to 8)."""
        arrs = [_as_f32_clip(c) for c in clips]
        n = np.array([a.size for a in arrs], dtype=np.int32)
        ld = int(max(8, (int(n.max()) + 7) // 8 * 8)) if len(arrs) else 8
        B = len(arrs)
        if self._pin_in is None or self._pin_in.shape[0] < B or self._pin_in.shape[1] != ld:
            self._pin_in = torch.zeros((B, ld), dtype=torch.float32).pin_memory()
        buf = self._pin_in[:B]
        nb = buf.numpy()
        for i, a in enumerate(arrs):
            nb[i, : a.size] = a
            nb[i, a.size:] = 0.0
        return buf, n

    def _call_dev(self, fn, audio: torch.Tensor, n_samples: np.ndarray, out: torch.Tensor, stream=None):
        n_samples = np.ascontiguousarray(n_samples, dtype=np.int32)
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        rc = fn(self._h, audio.data_ptr(), row_pitch(audio), n_samples.ctypes.data_as(_lib.c_i32p), audio.shape[0],
                out.data_ptr(), st.cuda_stream)
        if rc != 0:
            raise SsrError(self._err())

    # ------------------------------------------------------------------ public API
    def pooled_device(self, audio: torch.Tensor, n_samples, out: torch.Tensor | None = None,
                      stream=None) -> torch.Tensor:
        """audio: CUDA float32 [B, ld]; n_samples: host ints [B]. Returns CUDA float32 [B, L+1, D] (asynchronous)."""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2 and audio.stride(1) == 1
        B = audio.shape[0]
        if out is None:
            out = torch.empty((B, self.layers + 1, self.hidden), dtype=torch.float32, device=audio.device)
        self._call_dev(self._pooled_fn, audio, np.asarray(n_samples, dtype=np.int32), out, stream)
        return out

    def pooled_pinned(self, audio_host: torch.Tensor, n_samples, out_host: torch.Tensor) -> torch.Tensor:
        """Host buffers in and out through the C ABI's *_host entry point (H2D, run, D2H, stream sync inside).
        Pass pinned tensors for full PCIe speed: audio_host float32 [B, ld], out_host float32 [B, L+1, D]."""
        assert not audio_host.is_cuda and audio_host.dtype == torch.float32 and audio_host.stride(1) == 1
        assert not out_host.is_cuda and out_host.dtype == torch.float32 and out_host.is_contiguous()
        n = np.ascontiguousarray(n_samples, dtype=np.int32)
        rc = self._pooled_host_fn(self._h, audio_host.data_ptr(), row_pitch(audio_host), n.ctypes.data_as(_lib.c_i32p),
                                  audio_host.shape[0], out_host.data_ptr())
        if rc != 0:
            raise SsrError(self._err())
        return out_host

    def pooled_stream(self, batches, depth: int = 2):
        """Throughput path over many batches. `batches` yields (audio_host [B, ld] float32 — pinned for full PCIe
        speed —, n_samples [B]); this generator yields one pinned float32 [B, L+1, D] tensor per batch, in order.
        The H2D copy of batch i+1 and the D2H copy of batch i-1 overlap the forward of batch i (three CUDA streams,
        `depth` buffers each), so the copies the synchronous `pooled_pinned` pays per call disappear from the
        critical path. A yielded tensor is a view of a reused pinned buffer: consume (or copy) it before advancing
        the generator again. With a problem in one batch the SsrError surfaces from the generator."""
        dev = torch.device("cuda", self.device)
        L1, D = self.layers + 1, self.hidden
        # streams and (pinned / device) buffers live on the engine: allocating pinned memory per call would cost more
        # than the copies this path hides
        st = getattr(self, "_stream_state", None)
        if st is None or len(st["dev_in"]) != depth:
            st = {"streams": tuple(torch.cuda.Stream(dev) for _ in range(3)), "dev_in": [None] * depth,
                  "dev_out": [None] * depth, "pin_out": [None] * depth}
            self._stream_state = st
        s_in, s_run, s_out = st["streams"]
        dev_in, dev_out, pin_out = st["dev_in"], st["dev_out"], st["pin_out"]
        torch.cuda.synchronize(dev)  # a previous, abandoned generator may have left work on the side streams
        run_done = [torch.cuda.Event() for _ in range(depth)]
        in_ready = [torch.cuda.Event() for _ in range(depth)]
        out_done = [torch.cuda.Event() for _ in range(depth)]
        pending = []
        for k, (ah, n) in enumerate(batches):
            slot = k % depth
            ah = torch.as_tensor(ah)
            assert not ah.is_cuda and ah.dtype == torch.float32 and ah.dim() == 2 and ah.stride(1) == 1
            B, ld = ah.shape
            if dev_in[slot] is None or dev_in[slot].shape[0] < B or dev_in[slot].shape[1] != ld:
                torch.cuda.synchronize(dev)  # (re)allocation: nothing may still be using the old buffer
                dev_in[slot] = torch.empty((B, ld), dtype=torch.float32, device=dev)
            if dev_out[slot] is None or dev_out[slot].shape[0] < B:
                torch.cuda.synchronize(dev)
                dev_out[slot] = torch.empty((B, L1, D), dtype=torch.float32, device=dev)
                pin_out[slot] = torch.empty((B, L1, D), dtype=torch.float32).pin_memory()
            with torch.cuda.stream(s_in):
                s_in.wait_event(run_done[slot])      # the forward that last read this input buffer has finished
                dev_in[slot][:B].copy_(ah, non_blocking=True)
                in_ready[slot].record(s_in)
            s_run.wait_event(in_ready[slot])
            s_run.wait_event(out_done[slot])         # the D2H that last read this output buffer has finished
            self.pooled_device(dev_in[slot][:B], n, out=dev_out[slot][:B], stream=s_run)
            run_done[slot].record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(run_done[slot])
                pin_out[slot][:B].copy_(dev_out[slot][:B], non_blocking=True)
                out_done[slot].record(s_out)
            pending.append((slot, B))
            if len(pending) == depth:
                s0, b0 = pending.pop(0)
                out_done[s0].synchronize()
                yield pin_out[s0][:b0]
        for s0, b0 in pending:
            out_done[s0].synchronize()
            yield pin_out[s0][:b0]

    def pooled(self, clips: Sequence) -> np.ndarray:
        """Host in, host out (the call the drop-in shim makes): list of 1-D float clips -> float32 [B, L+1, D]."""
        if len(clips) == 0:
            return np.zeros((0, self.layers + 1, self.hidden), dtype=np.float32)
        buf, n = self._stage(clips)
        B = len(clips)
        if self._pin_out is None or self._pin_out.shape[0] < B:
            self._pin_out = torch.empty((B, self.layers + 1, self.hidden), dtype=torch.float32).pin_memory()
        out = self._pin_out[:B]
        rc = self._pooled_host_fn(self._h, buf.data_ptr(), buf.stride(0), n.ctypes.data_as(_lib.c_i32p), B,
                                  out.data_ptr())
        if rc != 0:
            raise SsrError(self._err())
        return out.numpy().copy()


class WavLMEngine(_EngineBase):
    family = _lib.SSR_WAVLM

    def __init__(self, state_dict: dict, config, do_normalize: bool, device: int = 0):
        desc = _lib.ModelDesc()
        desc.family = _lib.SSR_WAVLM
        desc.hidden = config.hidden_size
        desc.layers = config.num_hidden_layers
        desc.heads = config.num_attention_heads
        desc.ffn = config.intermediate_size
        desc.feat_norm = _lib.SSR_FEAT_NORM_LAYER if config.feat_extract_norm == "layer" else _lib.SSR_FEAT_NORM_GROUP
        desc.stable_ln = int(bool(config.do_stable_layer_norm))
        desc.do_normalize = int(bool(do_normalize))
        desc.n_mels = 0
        self._check_config(config)
        tensors = {}
        for k, v in state_dict.items():
            if k.startswith("wavlm."):
                k = k[len("wavlm."):]
            if k.startswith(("feature_extractor.", "feature_projection.", "encoder.")):
                tensors[k] = v.detach().to(torch.float32).cpu().numpy()
        super().__init__(desc, tensors, device)
        self._pooled_fn = self._lib.ssr_wavlm_pooled
        self._pooled_host_fn = self._lib.ssr_wavlm_pooled_host

    @staticmethod
    def _check_config(cfg):
        want = dict(conv_dim=(512,) * 7, conv_stride=(5, 2, 2, 2, 2, 2, 2), conv_kernel=(10, 3, 3, 3, 3, 2, 2),
                    num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16, num_buckets=320,
                    max_bucket_distance=800, conv_bias=False)
        for k, v in want.items():
            got = getattr(cfg, k)
            got = tuple(got) if isinstance(got, (list, tuple)) else got
            if got != v:
                raise SsrError(f"unsupported WavLM config: {k}={got!r} (this engine implements {v!r})")
        if cfg.hidden_act != "gelu" or cfg.feat_extract_activation != "gelu":
            raise SsrError("unsupported WavLM config: activations must be erf-GELU")
        if float(cfg.layer_norm_eps) != 1e-5:
            raise SsrError("unsupported WavLM config: layer_norm_eps must be 1e-5")

    @classmethod
    def from_hf(cls, model, feature_extractor=None, device: int = 0) -> "WavLMEngine":
        do_norm = bool(getattr(feature_extractor, "do_normalize", False)) if feature_extractor is not None else False
        return cls(model.state_dict(), model.config, do_norm, device)


class WhisperEncoderEngine(_EngineBase):
    family = _lib.SSR_WHISPER_ENC

    def __init__(self, state_dict: dict, config, mel_filters: np.ndarray | None = None, device: int = 0):
        desc = _lib.ModelDesc()
        desc.family = _lib.SSR_WHISPER_ENC
        desc.hidden = config.d_model
        desc.layers = config.encoder_layers
        desc.heads = config.encoder_attention_heads
        desc.ffn = config.encoder_ffn_dim
        desc.n_mels = config.num_mel_bins
        if config.max_source_positions != 1500:
            raise SsrError("unsupported Whisper config: max_source_positions must be 1500")
        if config.activation_function != "gelu":
            raise SsrError("unsupported Whisper config: activation must be erf-GELU")
        tensors = {}
        has_decoder = False
        for k, v in state_dict.items():
            if k.startswith("model."):
                k = k[len("model."):]
            if k.startswith("decoder."):
                # decoder start-token probe (REF/whisper_embeddings_large.py:257-262): only row 0 of the two
                # embedding tables is needed; the single-token self-attention never touches q_proj / k_proj.
                if k in ("decoder.embed_tokens.weight", "decoder.embed_positions.weight"):
                    tensors[k + "[0]"] = v[0].detach().to(torch.float32).cpu().numpy()
                    has_decoder = True
                elif ".self_attn.q_proj." in k or ".self_attn.k_proj." in k:
                    continue
                else:
                    tensors[k] = v.detach().to(torch.float32).cpu().numpy()
                continue
            if k.startswith("encoder."):
                k = k[len("encoder."):]
            elif k.startswith("proj_out."):
                continue
            if k.startswith(("conv1.", "conv2.", "embed_positions.", "layers.", "layer_norm.")):
                tensors[k] = v.detach().to(torch.float32).cpu().numpy()
        self.decoder_layers = 0
        if has_decoder:
            if config.decoder_attention_heads != config.encoder_attention_heads:
                raise SsrError("unsupported Whisper config: decoder and encoder head counts differ")
            if getattr(config, "scale_embedding", False):
                raise SsrError("unsupported Whisper config: scale_embedding")
            desc.reserved[0] = config.decoder_layers
            desc.reserved[1] = config.decoder_ffn_dim
            self.decoder_layers = int(config.decoder_layers)
        if mel_filters is None:
            mel_filters = whisper_mel_filters(config.num_mel_bins)
        mel_filters = np.asarray(mel_filters, dtype=np.float32)
        if mel_filters.shape != (201, config.num_mel_bins):
            raise SsrError(f"mel_filters must be [201, {config.num_mel_bins}], got {mel_filters.shape}")
        tensors["mel_filters"] = mel_filters
        super().__init__(desc, tensors, device)
        self._pooled_fn = self._lib.ssr_whisper_enc_pooled
        self._pooled_host_fn = self._lib.ssr_whisper_enc_pooled_host

    @classmethod
    def from_hf(cls, model, feature_extractor=None, device: int = 0) -> "WhisperEncoderEngine":
        """`model` may be WhisperModel, WhisperForConditionalGeneration or a bare WhisperEncoder."""
        enc = model
        for attr in ("model", "encoder"):
            if hasattr(enc, attr) and not hasattr(enc, "conv1"):
                enc = getattr(enc, attr)
        if not hasattr(enc, "conv1") and hasattr(enc, "encoder"):
            enc = enc.encoder
        fe = getattr(feature_extractor, "feature_extractor", feature_extractor)  # WhisperProcessor -> its FE
        mel = getattr(fe, "mel_filters", None) if fe is not None else None
        if fe is not None:
            for k, v in dict(n_fft=400, hop_length=160, chunk_length=30, sampling_rate=16000).items():
                if getattr(fe, k, v) != v:
                    raise SsrError(f"unsupported WhisperFeatureExtractor: {k}={getattr(fe, k)} (engine implements {v})")
            if float(getattr(fe, "dither", 0.0)) != 0.0:
                raise SsrError("unsupported WhisperFeatureExtractor: dither must be 0")
            if bool(getattr(fe, "do_normalize", False)):
                raise SsrError("unsupported WhisperFeatureExtractor: do_normalize must be False")
        sd = {"encoder." + k: v for k, v in enc.state_dict().items()}
        full = getattr(model, "model", model)  # WhisperForConditionalGeneration -> WhisperModel
        dec = getattr(full, "decoder", None)
        if dec is not None and hasattr(dec, "state_dict"):
            sd.update({"decoder." + k: v for k, v in dec.state_dict().items()})
        return cls(sd, enc.config, mel, device)

    def pooled_with_decoder(self, clips: Sequence) -> tuple[np.ndarray, np.ndarray]:
        """(encoder pooled [B, Le+1, D], decoder start-token hidden states [B, Ld+1, D]) — everything
        REF/whisper_embeddings_large.py:234-299 derives its `encoder_layer_*` / `decoder_layer_*` outputs from."""
        if self.decoder_layers == 0:
            raise SsrError("this engine was built without decoder weights")
        B = len(clips)
        if B == 0:
            return (np.zeros((0, self.layers + 1, self.hidden), np.float32),
                    np.zeros((0, self.decoder_layers + 1, self.hidden), np.float32))
        buf, n = self._stage(clips)
        enc = torch.empty((B, self.layers + 1, self.hidden), dtype=torch.float32).pin_memory()
        dec = torch.empty((B, self.decoder_layers + 1, self.hidden), dtype=torch.float32).pin_memory()
        rc = self._lib.ssr_whisper_full_host(self._h, buf.data_ptr(), buf.stride(0), n.ctypes.data_as(_lib.c_i32p), B,
                                             enc.data_ptr(), dec.data_ptr())
        if rc != 0:
            raise SsrError(self._err())
        return enc.numpy().copy(), dec.numpy().copy()

    def logmel_device(self, audio: torch.Tensor, n_samples, stream=None) -> torch.Tensor:
        """== WhisperFeatureExtractor(audio).input_features : CUDA float32 [B, 80, 3000]."""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2 and audio.stride(1) == 1
        out = torch.empty((audio.shape[0], int(self.desc.n_mels), 3000), dtype=torch.float32, device=audio.device)
        self._call_dev(self._lib.ssr_logmel, audio, np.asarray(n_samples, dtype=np.int32), out, stream)
        return out

    def logmel(self, clips: Sequence) -> np.ndarray:
        buf, n = self._stage(clips)
        dev = buf.to(f"cuda:{self.device}", non_blocking=True)
        out = self.logmel_device(dev, n)
        return out.cpu().numpy()


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous clip shard of rank `rank`: gathered shards are already in global order."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def iter_batches(n_items: int, batch: int) -> Iterable[tuple[int, int]]:
    for lo in range(0, n_items, batch):
        yield lo, min(lo + batch, n_items)
