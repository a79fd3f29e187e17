"""TEST INFRASTRUCTURE — float64 numpy restatement of the data-parallel classifier head.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product never does.

The head is a NEW component (the reference trains only sklearn estimators), so there is no reference output to pin
against — "parity unpinned" for the classifier itself. What IS taken from the reference and checked against the real
thing in tests/test_head_cpu.py:
    StandardScaler semantics      sklearn.preprocessing.StandardScaler  (REF/model_training_1.py:658-661)
    balanced class weights        sklearn.utils.class_weight.compute_class_weight('balanced')  (REF :576-589)
    evaluation metric             sklearn.metrics.balanced_accuracy_score  (REF :672)
The training recipe itself (MLP D->H->C with ReLU, class-weighted softmax cross-entropy normalised by the weight sum,
Adam with L2 weight decay, the seeded global minibatch schedule) is defined here and in DESIGN.md; the CUDA path has
to reproduce this file's float64 result to float32 accuracy.
"""
from __future__ import annotations

import numpy as np


def balanced_class_weights(counts: np.ndarray) -> np.ndarray:
    counts = np.asarray(counts, np.float64)
    return counts.sum() / (len(counts) * counts)


def scaler_fit(X: np.ndarray):
    X = np.asarray(X, np.float64)
    mean = X.mean(0)
    var = ((X - mean) ** 2).mean(0)
    scale = np.sqrt(var)
    scale[scale < 10 * np.finfo(np.float64).eps] = 1.0  # sklearn _handle_zeros_in_scale
    return mean, scale


def init_params(D: int, H: int, C: int, seed: int) -> np.ndarray:
    """torch.nn.Linear's default scheme (uniform +-1/sqrt(fan_in) for weight and bias), from a numpy generator."""
    rng = np.random.default_rng([seed, 0x4845_4144])
    b1, b2 = 1.0 / np.sqrt(D), 1.0 / np.sqrt(H)
    parts = [rng.uniform(-b1, b1, H * D), rng.uniform(-b1, b1, H), rng.uniform(-b2, b2, C * H), rng.uniform(-b2, b2, C)]
    return np.concatenate(parts).astype(np.float32)


def split(params: np.ndarray, D: int, H: int, C: int):
    o = 0
    W1 = params[o:o + H * D].reshape(H, D); o += H * D
    b1 = params[o:o + H]; o += H
    W2 = params[o:o + C * H].reshape(C, H); o += C * H
    b2 = params[o:o + C]
    return W1, b1, W2, b2


def epoch_permutation(N: int, seed: int, epoch: int) -> np.ndarray:
    return np.random.default_rng([seed, epoch]).permutation(N)


def grad_sums(Xs: np.ndarray, y: np.ndarray, params: np.ndarray, H: int, C: int, class_w: np.ndarray) -> np.ndarray:
    """Unnormalised [P + 2]: sum_i w_i dloss_i/dparam | sum_i w_i loss_i | sum_i w_i  over the given (scaled) rows."""
    n, D = Xs.shape
    W1, b1, W2, b2 = split(params.astype(np.float64), D, H, C)
    if n == 0:
        return np.zeros(params.size + 2)
    Z1 = Xs @ W1.T + b1
    A1 = np.maximum(Z1, 0.0)
    Z2 = A1 @ W2.T + b2
    Z2 = Z2 - Z2.max(1, keepdims=True)
    logp = Z2 - np.log(np.exp(Z2).sum(1, keepdims=True))
    w = class_w[y]
    loss = -(w * logp[np.arange(n), y]).sum()
    dZ2 = np.exp(logp)
    dZ2[np.arange(n), y] -= 1.0
    dZ2 *= w[:, None]
    dW2 = dZ2.T @ A1
    db2 = dZ2.sum(0)
    dZ1 = (dZ2 @ W2) * (Z1 > 0)
    dW1 = dZ1.T @ Xs
    db1 = dZ1.sum(0)
    return np.concatenate([dW1.ravel(), db1, dW2.ravel(), db2, [loss, w.sum()]])


def adam_step(params, G, m, v, step, lr, beta1, beta2, eps, wd):
    P = params.size
    g = G[:P] / G[P + 1] + wd * params
    m[:] = beta1 * m + (1 - beta1) * g
    v[:] = beta2 * v + (1 - beta2) * g * g
    params -= lr * (m / (1 - beta1 ** step)) / (np.sqrt(v / (1 - beta2 ** step)) + eps)


def train(X: np.ndarray, y: np.ndarray, C: int, hidden=64, epochs=3, batch_size=256, lr=1e-3, weight_decay=1e-4,
          beta1=0.9, beta2=0.999, eps=1e-8, class_weight="balanced", seed=0, max_steps=None):
    """Single-process float64 training on the WHOLE data set. Returns dict(params, mean, scale, class_w, losses)."""
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.int64)
    N, D = X.shape
    counts = np.bincount(y, minlength=C)
    class_w = balanced_class_weights(counts) if class_weight == "balanced" else np.ones(C)
    mean, scale = scaler_fit(X)
    # the device keeps mean and 1/scale in float32
    mean32 = mean.astype(np.float32).astype(np.float64)
    inv32 = (1.0 / scale).astype(np.float32).astype(np.float64)
    cw32 = class_w.astype(np.float32).astype(np.float64)
    Xs = (X - mean32) * inv32
    params = init_params(D, hidden, C, seed).astype(np.float64)
    m = np.zeros_like(params)
    v = np.zeros_like(params)
    losses, step = [], 0
    for ep in range(epochs):
        perm = epoch_permutation(N, seed, ep)
        for lo in range(0, N, batch_size):
            idx = perm[lo:lo + batch_size]
            G = grad_sums(Xs[idx], y[idx], params, hidden, C, cw32)
            step += 1
            adam_step(params, G, m, v, step, lr, beta1, beta2, eps, weight_decay)
            losses.append(G[-2] / G[-1])
            if max_steps is not None and step >= max_steps:
                return dict(params=params, mean=mean, scale=scale, class_w=class_w, losses=losses)
    return dict(params=params, mean=mean, scale=scale, class_w=class_w, losses=losses)


def predict(X, params, mean, scale, H, C):
    X = np.asarray(X, np.float64)
    W1, b1, W2, b2 = split(np.asarray(params, np.float64), X.shape[1], H, C)
    Xs = (X - mean.astype(np.float32)) * (1.0 / scale).astype(np.float32)
    return (np.maximum(Xs @ W1.T + b1, 0) @ W2.T + b2).argmax(1)


def synthetic_clusters(n: int, D: int, C: int, seed=0, spread=6.0, imbalance=0.6):
    """BASELINE configs[3] stand-in: C Gaussian clusters in R^D with geometrically imbalanced class counts."""
    rng = np.random.default_rng(seed)
    p = imbalance ** np.arange(C)
    p /= p.sum()
    y = rng.choice(C, size=n, p=p)
    for c in range(C):  # every class present
        y[c] = c
    centres = rng.standard_normal((C, D)) * spread / np.sqrt(D)
    offset = rng.standard_normal(D) * 3.0            # non-zero feature means and scales: the scaler matters
    scale = np.exp(rng.standard_normal(D) * 0.5)
    X = (centres[y] + rng.standard_normal((n, D))) * scale + offset
    return X.astype(np.float32), y.astype(np.int32)
