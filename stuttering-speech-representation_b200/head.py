"""Classifier head on the pooled embeddings, trained data-parallel on the GPUs (BASELINE configs[3]).

NEW component — the reference trains sklearn estimators (`Pipeline([StandardScaler, SVC | RandomForest])`,
REF/model_training_1.py:630-680). Kept from it: the scaler semantics, `class_weight='balanced'` (REF :576-589), the
`fit(X, y)` / `predict(X)` shape of the estimator and the metrics (balanced accuracy, REF :672). The model is a
one-hidden-layer MLP under class-weighted softmax cross-entropy and Adam; every FLOP runs in `csrc/head.cu` behind
the C ABI (`ssr_head_*`), torch is used for device buffers and the NCCL all-reduce only.

Data parallelism: one process per GPU (`torch.distributed`, NCCL); rank r holds a contiguous shard of the rows, in
rank order. Every step ALL ranks derive the same seeded global minibatch (a slice of a global permutation), each
computes unnormalised gradient sums over the rows of that minibatch it owns, ONE all-reduce(sum) of the flat
`[P + 2]` buffer (gradient | weighted loss sum | weight sum) follows, and the Adam kernel divides by the reduced weight
sum on the device. The update is therefore the same function of the same global minibatch at any world size: weights
after training agree between 1 and N GPUs up to the all-reduce's summation order.
"""
from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .engine import SsrError

logger = logging.getLogger("ssr_b200")


@dataclass
class HeadConfig:
    hidden: int = 256
    epochs: int = 20
    batch_size: int = 1024          # GLOBAL minibatch (all ranks together)
    lr: float = 1e-3
    weight_decay: float = 1e-4
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    class_weight: str | None = "balanced"
    seed: int = 0


def balanced_class_weights(counts) -> np.ndarray:
    """sklearn compute_class_weight('balanced'): n_samples / (n_classes * count_c)."""
    counts = np.asarray(counts, np.float64)
    return counts.sum() / (len(counts) * counts)


def init_params(D: int, H: int, n_classes: int, seed: int) -> np.ndarray:
    """torch.nn.Linear's default init (uniform +-1/sqrt(fan_in)), drawn on the host so every rank starts identical."""
    rng = np.random.default_rng([seed, 0x4845_4144])
    b1, b2 = 1.0 / np.sqrt(D), 1.0 / np.sqrt(H)
    parts = [rng.uniform(-b1, b1, H * D), rng.uniform(-b1, b1, H), rng.uniform(-b2, b2, n_classes * H),
             rng.uniform(-b2, b2, n_classes)]
    return np.concatenate(parts).astype(np.float32)


def epoch_permutation(N: int, seed: int, epoch: int) -> np.ndarray:
    return np.random.default_rng([seed, epoch]).permutation(N)


def local_rows(global_idx: np.ndarray, offset: int, n_local: int) -> np.ndarray:
    """Rows of a global minibatch owned by the shard [offset, offset + n_local), as local indices, order kept."""
    sel = global_idx[(global_idx >= offset) & (global_idx < offset + n_local)]
    return (sel - offset).astype(np.int32)


def _dist():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


class GpuHead:
    """`fit(X_local, y_local)` / `predict(X)` / `predict_proba(X)` / `score(X, y)` (balanced accuracy)."""

    def __init__(self, config: HeadConfig | None = None, device: int = 0, distributed: bool | None = None, **kw):
        """distributed: None = use the initialised torch.distributed group if there is one; False = train on this
        process's rows only even inside a group (e.g. a single-GPU control run)."""
        self.cfg = config or HeadConfig(**kw)
        self.distributed = distributed
        if not torch.cuda.is_available():
            raise SsrError("ssr_b200 classifier head needs a CUDA device; there is no CPU fallback")
        self._lib = _lib.load()
        self.device = torch.device("cuda", int(device))
        self.n_classes = self.D = None
        self.params = self.mean = self.inv_std = None
        self.losses: list[float] = []
        self._work = None

    # ------------------------------------------------------------------ helpers
    def _group(self):
        return None if self.distributed is False else _dist()

    def _call(self, fn, *args):
        err = C.create_string_buffer(256)
        st = torch.cuda.current_stream(self.device).cuda_stream
        rc = fn(*args, st, err, 256)
        if rc != 0:
            raise SsrError(err.value.decode(errors="replace"))

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._work is None or self._work.numel() < nbytes:
            self._work = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=self.device)
        return self._work

    def _to_dev(self, X) -> torch.Tensor:
        t = torch.as_tensor(X)
        return t.to(self.device, dtype=torch.float32).contiguous()

    # ------------------------------------------------------------------ scaler
    def _fit_scaler(self, X: torch.Tensor, N: int):
        dist = self._group()
        n, D = X.shape
        s = torch.zeros(D, dtype=torch.float64, device=self.device)
        self._call(self._lib.ssr_head_scaler_stats, X.data_ptr(), n, D, X.stride(0), None, s.data_ptr())
        if dist:
            dist.all_reduce(s)
        mean = s / N
        ss = torch.zeros(D, dtype=torch.float64, device=self.device)
        self._call(self._lib.ssr_head_scaler_stats, X.data_ptr(), n, D, X.stride(0), mean.data_ptr(), ss.data_ptr())
        if dist:
            dist.all_reduce(ss)
        scale = torch.sqrt(ss / N)
        scale[scale < 10 * np.finfo(np.float64).eps] = 1.0   # sklearn _handle_zeros_in_scale
        self.mean64, self.scale64 = mean.cpu().numpy(), scale.cpu().numpy()
        self.mean = mean.float().contiguous()
        self.inv_std = (1.0 / scale).float().contiguous()

    # ------------------------------------------------------------------ training
    def fit(self, X, y, n_classes: int | None = None, max_steps: int | None = None):
        """X: [n_local, D] (numpy or torch, host or device), y: int labels in [0, n_classes). Under an initialised
        torch.distributed group every rank passes ITS contiguous shard, in rank order."""
        cfg = self.cfg
        dist = self._group()
        X = self._to_dev(X)
        y_host = np.ascontiguousarray(np.asarray(torch.as_tensor(y).cpu()), dtype=np.int64)
        n_local, D = X.shape
        assert y_host.shape[0] == n_local
        y_dev = torch.from_numpy(y_host.astype(np.int32)).to(self.device)
        # shard geometry
        if dist:
            sizes = torch.zeros(dist.get_world_size(), dtype=torch.int64, device=self.device)
            sizes[dist.get_rank()] = n_local
            dist.all_reduce(sizes)
            sizes = sizes.cpu().numpy()
            offset, N = int(sizes[: dist.get_rank()].sum()), int(sizes.sum())
        else:
            offset, N = 0, n_local
        # classes and weights
        c_local = int(y_host.max(initial=-1)) + 1
        if n_classes is None:
            cmax = torch.tensor([c_local], dtype=torch.int64, device=self.device)
            if dist:
                dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
            n_classes = int(cmax.item())
        counts = torch.from_numpy(np.bincount(y_host, minlength=n_classes).astype(np.int64)).to(self.device)
        if dist:
            dist.all_reduce(counts)
        counts = counts.cpu().numpy()
        if cfg.class_weight == "balanced":
            if (counts == 0).any():
                raise SsrError("class_weight='balanced' needs every class present in the training labels")
            cw = balanced_class_weights(counts)
        else:
            cw = np.ones(n_classes)
        self.class_weight_ = cw
        cw_dev = torch.from_numpy(cw.astype(np.float32)).to(self.device)
        self.n_classes, self.D, H = n_classes, D, cfg.hidden
        P = int(self._lib.ssr_head_param_count(D, H, n_classes))
        if P < 0:
            raise SsrError(f"unsupported head dimensions D={D} H={H} C={n_classes}")
        self._fit_scaler(X, N)
        self.params = torch.from_numpy(init_params(D, H, n_classes, cfg.seed)).to(self.device)
        m = torch.zeros(P, dtype=torch.float32, device=self.device)
        v = torch.zeros(P, dtype=torch.float32, device=self.device)
        G = torch.zeros(P + 2, dtype=torch.float32, device=self.device)
        loss_log = []
        step = 0
        self.losses = []
        done = False
        for ep in range(cfg.epochs):
            perm = epoch_permutation(N, cfg.seed, ep)
            # this rank's rows of every minibatch of the epoch, uploaded once
            chunks = [local_rows(perm[lo:lo + cfg.batch_size], offset, n_local) for lo in range(0, N, cfg.batch_size)]
            starts = np.cumsum([0] + [len(c) for c in chunks])
            rows_dev = torch.from_numpy(np.concatenate(chunks) if chunks else np.zeros(0, np.int32)).to(self.device)
            for t, rows in enumerate(chunks):
                nb = len(rows)
                work = self._workspace(int(self._lib.ssr_head_work_bytes(max(nb, 1), D, H, n_classes)))
                self._call(self._lib.ssr_head_grad, X.data_ptr(), y_dev.data_ptr(),
                           rows_dev.data_ptr() + 4 * int(starts[t]), nb, D, H, n_classes, self.mean.data_ptr(),
                           self.inv_std.data_ptr(), self.params.data_ptr(), cw_dev.data_ptr(), G.data_ptr(),
                           work.data_ptr(), work.numel())
                if dist:
                    dist.all_reduce(G)
                step += 1
                self._call(self._lib.ssr_head_adam, self.params.data_ptr(), G.data_ptr(), m.data_ptr(), v.data_ptr(),
                           P, cfg.lr, cfg.beta1, cfg.beta2, cfg.eps, cfg.weight_decay, step)
                loss_log.append(G[P:P + 2].clone())
                if max_steps is not None and step >= max_steps:
                    done = True
                    break
            if done:
                break
        if loss_log:
            ll = torch.stack(loss_log).cpu().numpy().astype(np.float64)
            self.losses = list(ll[:, 0] / ll[:, 1])
        self.steps_ = step
        return self

    # ------------------------------------------------------------------ inference
    def _predict(self, X, want_proba: bool):
        if self.params is None:
            raise SsrError("GpuHead is not fitted")
        X = self._to_dev(X)
        n, D = X.shape
        assert D == self.D
        pred = torch.empty(n, dtype=torch.int32, device=self.device)
        proba = torch.empty((n, self.n_classes), dtype=torch.float32, device=self.device) if want_proba else None
        chunk = 1 << 16
        for lo in range(0, n, chunk):
            hi = min(lo + chunk, n)
            work = self._workspace(4 * (hi - lo) * self.cfg.hidden)
            self._call(self._lib.ssr_head_predict, X[lo:hi].data_ptr(), hi - lo, D, self.cfg.hidden, self.n_classes,
                       self.mean.data_ptr(), self.inv_std.data_ptr(), self.params.data_ptr(), pred[lo:hi].data_ptr(),
                       None if proba is None else proba[lo:hi].data_ptr(), work.data_ptr(), work.numel())
        return pred, proba

    def predict(self, X) -> np.ndarray:
        return self._predict(X, False)[0].cpu().numpy().astype(np.int64)

    def predict_proba(self, X) -> np.ndarray:
        return self._predict(X, True)[1].cpu().numpy()

    def score(self, X, y) -> float:
        """Balanced accuracy (mean per-class recall), sklearn.metrics.balanced_accuracy_score."""
        return balanced_accuracy(np.asarray(y), self.predict(X))


def balanced_accuracy(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    y_true = np.asarray(y_true).astype(np.int64)
    y_pred = np.asarray(y_pred).astype(np.int64)
    recalls = [float((y_pred[y_true == c] == c).mean()) for c in np.unique(y_true)]
    return float(np.mean(recalls))
