"""CPU tier for the classifier-head row: the float64 oracle against the sklearn pieces the reference
uses (StandardScaler, balanced class weights, balanced accuracy), the product's host-side logic against the oracle, and
a world_size-2 gloo run of the data-parallel decomposition (schedule, row ownership, one all-reduce per step)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scaler_and_class_weights_match_sklearn():
    from sklearn.preprocessing import StandardScaler
    from sklearn.utils.class_weight import compute_class_weight

    from oracle import head_oracle as ho

    X, y = ho.synthetic_clusters(500, 24, 5, seed=3)
    X[:, 7] = 2.5  # zero-variance feature: sklearn leaves it unscaled
    sk = StandardScaler().fit(X.astype(np.float64))
    mean, scale = ho.scaler_fit(X)
    np.testing.assert_allclose(mean, sk.mean_, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(scale, sk.scale_, rtol=1e-10)
    assert scale[7] == 1.0
    want = compute_class_weight("balanced", classes=np.arange(5), y=y)
    np.testing.assert_allclose(ho.balanced_class_weights(np.bincount(y, minlength=5)), want, rtol=1e-12)


def test_product_host_logic_matches_oracle():
    from sklearn.metrics import balanced_accuracy_score

    from oracle import head_oracle as ho
    from ssr_b200 import head

    np.testing.assert_array_equal(head.init_params(40, 16, 5, 7), ho.init_params(40, 16, 5, 7))
    np.testing.assert_array_equal(head.epoch_permutation(1000, 3, 2), ho.epoch_permutation(1000, 3, 2))
    assert not np.array_equal(head.epoch_permutation(1000, 3, 2), head.epoch_permutation(1000, 3, 3))
    np.testing.assert_allclose(head.balanced_class_weights([5, 10, 85]), ho.balanced_class_weights([5, 10, 85]))
    rng = np.random.default_rng(0)
    yt, yp = rng.integers(0, 4, 300), rng.integers(0, 4, 300)
    assert abs(head.balanced_accuracy(yt, yp) - balanced_accuracy_score(yt, yp)) < 1e-12
    # row ownership: the shards of a minibatch partition it, in order
    perm = head.epoch_permutation(103, 0, 0)
    mb = perm[16:48]
    bounds = [(0, 40), (40, 33), (73, 30)]
    parts = [head.local_rows(mb, off, n) + off for off, n in bounds]
    assert sorted(np.concatenate(parts).tolist()) == sorted(mb.tolist())
    for (off, n), p in zip(bounds, parts):
        np.testing.assert_array_equal(p, mb[(mb >= off) & (mb < off + n)])
        assert p.dtype == np.int32 or p.dtype == np.int64


def test_oracle_gradient_is_the_numerical_gradient():
    from oracle import head_oracle as ho

    rng = np.random.default_rng(1)
    D, H, C, n = 6, 5, 3, 20
    Xs = rng.standard_normal((n, D))
    y = rng.integers(0, C, n)
    cw = np.array([0.5, 1.0, 2.0])
    p = ho.init_params(D, H, C, 0).astype(np.float64)
    G = ho.grad_sums(Xs, y, p, H, C, cw)
    for i in rng.choice(p.size, 12, replace=False):
        e = np.zeros_like(p)
        e[i] = 1e-6
        num = (ho.grad_sums(Xs, y, p + e, H, C, cw)[-2] - ho.grad_sums(Xs, y, p - e, H, C, cw)[-2]) / 2e-6
        assert abs(num - G[i]) < 1e-5 * max(1.0, abs(G[i]))


def test_oracle_learns_the_synthetic_clusters():
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    from sklearn.svm import SVC

    from oracle import head_oracle as ho
    from ssr_b200.head import balanced_accuracy

    X, y = ho.synthetic_clusters(1500, 64, 6, seed=0)
    Xtr, ytr, Xte, yte = X[:1000], y[:1000], X[1000:], y[1000:]
    r = ho.train(Xtr, ytr, 6, hidden=32, epochs=12, batch_size=128, lr=3e-3, seed=0)
    assert r["losses"][-1] < 0.5 * r["losses"][0]
    ba = balanced_accuracy(yte, ho.predict(Xte, r["params"], r["mean"], r["scale"], 32, 6))
    svc = Pipeline([("scaler", StandardScaler()), ("classifier", SVC(kernel="rbf", C=10))]).fit(Xtr, ytr)
    ba_svc = balanced_accuracy(yte, svc.predict(Xte))   # the reference's yardstick, REF/model_training_1.py:658-680
    assert ba > 0.8 and ba > ba_svc - 0.05, (ba, ba_svc)


GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SSR_ROOT"])
from oracle import head_oracle as ho
from ssr_b200 import head, shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, D, H, C, B, seed = 203, 12, 8, 4, 32, 5
X, y = ho.synthetic_clusters(N, D, C, seed=1)
lo, hi = shard_range(N, rank, world)
Xl, yl = X[lo:hi].astype(np.float64), y[lo:hi]
# scaler by two all-reduced passes, as the product does it
s = torch.from_numpy(Xl.sum(0)); dist.all_reduce(s); mean = s.numpy() / N
ss = torch.from_numpy(((Xl - mean) ** 2).sum(0)); dist.all_reduce(ss); scale = np.sqrt(ss.numpy() / N)
m_ref, s_ref = ho.scaler_fit(X)
assert np.allclose(mean, m_ref, rtol=1e-12, atol=1e-12) and np.allclose(scale, s_ref, rtol=1e-10)
counts = torch.from_numpy(np.bincount(yl, minlength=C)); dist.all_reduce(counts)
cw = head.balanced_class_weights(counts.numpy())
# data-parallel training with the product's schedule / ownership functions and the oracle's arithmetic
mean32 = mean.astype(np.float32).astype(np.float64); inv32 = (1 / scale).astype(np.float32).astype(np.float64)
cw32 = cw.astype(np.float32).astype(np.float64)
Xs = (Xl - mean32) * inv32
p = head.init_params(D, H, C, seed).astype(np.float64); m = np.zeros_like(p); v = np.zeros_like(p)
step = 0
for ep in range(2):
    perm = head.epoch_permutation(N, seed, ep)
    for b0 in range(0, N, B):
        rows = head.local_rows(perm[b0:b0 + B], lo, hi - lo)
        G = torch.from_numpy(ho.grad_sums(Xs[rows], yl[rows], p, H, C, cw32))
        dist.all_reduce(G)
        step += 1
        ho.adam_step(p, G.numpy(), m, v, step, 1e-3, 0.9, 0.999, 1e-8, 1e-4)
ref = ho.train(X, y, C, hidden=H, epochs=2, batch_size=B, lr=1e-3, weight_decay=1e-4, seed=seed)
err = np.abs(p - ref["params"]).max()
assert err < 1e-10, err
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok", err)
"""


def test_data_parallel_decomposition_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, SSR_ROOT=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29671", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
