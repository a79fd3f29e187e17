"""Data-parallel classifier head on real GPUs (BASELINE configs[3]): run under torchrun with N ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29700 \
        tools/head_dp_check.py [--samples 200000] [--steps 60]

Checks that the weights after K steps at world size N equal a single-GPU control run on the full data (same global
minibatches) up to fp32 summation order, and prints one JSON line with device-timed samples/s (max over ranks).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssr_b200 import shard_range, synth  # noqa: E402
from ssr_b200.head import GpuHead  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=200000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--classes", type=int, default=8)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=60)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    X, y = synth.cluster_embeddings(a.samples, a.dim, a.classes, seed=0)
    lo, hi = shard_range(a.samples, rank, world)
    kw = dict(hidden=a.hidden, epochs=1 + 3 * a.steps * a.batch // a.samples, batch_size=a.batch, lr=1e-3, seed=0)
    Xl = torch.from_numpy(X[lo:hi]).cuda()
    head = GpuHead(device=local, **kw)
    head.fit(Xl, y[lo:hi], n_classes=a.classes, max_steps=3)  # warm-up (NCCL channels, workspace)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    def timed_fit(steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        head.fit(Xl, y[lo:hi], n_classes=a.classes, max_steps=steps)
        ev[1].record()
        torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1])], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # fit() = scaler fit + schedule upload + K steps: the per-step cost is the slope between two step counts
    ms_long = timed_fit(3 * a.steps)
    ms_fit = timed_fit(a.steps)  # run last: p_dp / losses below belong to the a.steps run compared with the control
    ms_step = (ms_long - ms_fit) / (2 * a.steps)
    ms = torch.tensor([ms_fit], device="cuda")
    p_dp = head.params.clone()
    out = {"workload": "classifier head D=%d H=%d C=%d, global batch %d" % (a.dim, a.hidden, a.classes, a.batch),
           "n_gpus": world, "steps": a.steps, "ms_per_step": round(ms_step, 4),
           "samples_per_s": round(a.batch / ms_step * 1e3, 1),
           "ms_fit_total": round(ms_fit, 3),
           "note": "ms_per_step = (fit of 3K steps - fit of K steps) / 2K, device-timed, max over ranks: the scaler fit "
                   "and the schedule upload (in ms_fit_total) are outside the per-step figure",
           "final_loss": round(float(head.losses[-1]), 5), "first_loss": round(float(head.losses[0]), 5)}
    if rank == 0:
        ctrl = GpuHead(device=local, distributed=False, **kw).fit(X, y, n_classes=a.classes, max_steps=a.steps)
        d = (p_dp - ctrl.params).abs()
        out["vs_single_gpu"] = {"max_abs_diff": float(d.max()), "median_abs_diff": float(d.median()),
                                "max_abs_param": float(ctrl.params.abs().max()),
                                "loss_rel_diff": float(np.abs(np.array(head.losses) / np.array(ctrl.losses) - 1).max())}
        out["train_balanced_accuracy"] = round(head.score(X[:20000], y[:20000]), 4)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
