#!/usr/bin/env python
"""Benchmark of the embedding-extraction hot path (BASELINE.json metric: embedded clips/sec, 3 s @ 16 kHz).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (port over HF transformers)

A "step" is one pass of the hot path over one batch of synthetic clips per GPU (headline workload =
BASELINE.json configs[1]: WavLM-Large, 256 clips of 3 s per GPU). One process per GPU (torchrun for N > 1); clips
are sharded by contiguous index range, the only collective is the all-gather of pooled embeddings (scaling: weak).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "embedded clips/sec (3s@16kHz), WavLM-L & Whisper-L encoder, 1/2/4/8 B200"
WAVLM_GFLOP_PER_CLIP = 109.62   # SURVEY.md 8(d), algorithmic 2*MAC
WHISPER_GFLOP_PER_CLIP = 2272.7


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d.get("bf16_tflops"), "bf16_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region: NVML polled every few milliseconds from a
    thread (the timed region of a default run is ~0.4 s, too short for `nvidia-smi -lms`, which is the fallback)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.samples, self.stop_flag, self.th = None, [], False, None

    def _nvml_loop(self, h):
        nv = self.nvml
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.nvml = nv
            self.th = threading.Thread(target=self._nvml_loop, args=(h,), daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=2)
            sm = [s[0] for s in self.samples]
            pw = [s[1] for s in self.samples]
            bits = 0
            for s in self.samples:
                bits |= int(s[2])
            reasons = sorted(k for k, b in self.REASON_BITS.items() if bits & b)
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.mx,
                    "power_w_max": round(max(pw), 2) if pw else None, "samples": len(sm), "reasons": reasons,
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi"}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def parity(got: np.ndarray, ref: np.ndarray) -> dict:
    """Stated tolerance per pooled layer vector: cosine >= 0.9999 and max-abs <= 1e-2 * max|ref| (north_star). The
    per-layer vectors (worst clip per layer) show which depth owns the error budget."""
    g, r = got.astype(np.float64), ref.astype(np.float64)
    cos = (g * r).sum(-1) / np.maximum(np.sqrt((g * g).sum(-1) * (r * r).sum(-1)), 1e-30)
    rel = np.abs(g - r).max(-1) / np.maximum(np.abs(r).max(-1), 1e-30)
    return {"min_cos": float(cos.min()), "max_rel_err": float(rel.max()), "tolerance": "cos>=0.9999, rel<=1e-2",
            "ok": bool(cos.min() >= 0.9999 and rel.max() <= 1e-2), "clips": int(g.shape[0]),
            "max_rel_err_per_layer": [round(float(x), 6) for x in rel.max(0)],
            "min_cos_per_layer": [round(float(x), 7) for x in cos.min(0)]}


def workload_config(batch: int, n_gpus: int) -> dict:
    """The `config` object of BOTH arms (identical dicts: the driver compares them)."""
    return {"workload": "WavLM-Large (24 layers, d=1024) per-layer pooled embeddings, 3 s clips, "
                        f"batch {batch} per GPU (BASELINE configs[1])",
            "clips_per_gpu_per_step": batch, "clip_samples": 48000, "weights": "seeded random init (seed 0)",
            "output": f"[{batch}, 25, 1024] f32 per GPU", "parallelism": f"clip-sharded x{n_gpus}",
            "l2_policy": "no explicit flush: per-step working set (~6 GB of activations) >> 126 MB L2"}


# ---------------------------------------------------------------------------------------------- our arm
class StepRunner:
    """One bench step = one forward of the local shard (+ for N > 1 the all-gather of the pooled embeddings).
    Gather modes: `inline` — NCCL on the compute stream; `overlap` — NCCL on a side stream into double-buffered outputs,
    overlapping the next step's forward; `ce` — the same overlap, but the gather is N peer copies out of symmetric
    memory on the COPY ENGINES (torch.distributed._symmetric_memory: the forward writes its pooled output into a
    symmetric buffer, two device-side barriers bracket the copies), so that no SM is taken from the forward's
    persistent kernels."""

    def __init__(self, eng, audio, n_samples, world, dev, gather_mode):
        import torch

        self.torch, self.eng, self.audio, self.n, self.world, self.dev = torch, eng, audio, n_samples, world, dev
        B, L1, D = audio.shape[0], eng.layers + 1, eng.hidden
        self.shape = (B, L1, D)
        self.mode = gather_mode if world > 1 else "none"
        nb = 2 if self.mode in ("overlap", "ce") else 1
        self.hdl = None
        if self.mode == "ce":
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem

            self.outs = [symm_mem.empty(B, L1, D, dtype=torch.float32, device=dev) for _ in range(nb)]
            self.hdl = [symm_mem.rendezvous(t, dist.group.WORLD) for t in self.outs]
            self.rank = dist.get_rank()
        else:
            self.outs = [torch.empty((B, L1, D), dtype=torch.float32, device=dev) for _ in range(nb)]
        self.gath = [torch.empty((world * B, L1, D), dtype=torch.float32, device=dev) for _ in range(nb)] \
            if world > 1 else []
        self.side = torch.cuda.Stream(dev) if nb == 2 else None
        self.run_done = [torch.cuda.Event() for _ in range(nb)]
        self.gather_done = [torch.cuda.Event() for _ in range(nb)]
        self.k = 0

    def step(self):
        import torch.distributed as dist

        torch = self.torch
        s = self.k % len(self.outs)
        self.k += 1
        cur = torch.cuda.current_stream(self.dev)
        if self.side is not None:
            cur.wait_event(self.gather_done[s])  # the gather issued two steps ago has read outs[s]
        self.eng.pooled_device(self.audio, self.n, out=self.outs[s])
        if self.mode == "inline":
            dist.all_gather_into_tensor(self.gath[0], self.outs[0])
        elif self.mode == "overlap":
            self.run_done[s].record(cur)
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.run_done[s])
                dist.all_gather_into_tensor(self.gath[s], self.outs[s])
                self.gather_done[s].record(self.side)
        elif self.mode == "ce":
            B = self.shape[0]
            self.run_done[s].record(cur)
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.run_done[s])
                h = self.hdl[s]
                h.barrier()  # every rank has finished writing its outs[s]
                for st in range(self.world):
                    r = (self.rank - st) % self.world
                    self.gath[s][r * B:(r + 1) * B].copy_(h.get_buffer(r, self.shape, torch.float32), non_blocking=True)
                h.barrier()  # every rank has finished reading: outs[s] may be overwritten two steps from now
                self.gather_done[s].record(self.side)
        return s

    def drain(self):
        if self.side is not None:
            self.torch.cuda.current_stream(self.dev).wait_stream(self.side)


def bench_engine(eng, clips_np: np.ndarray, n_samples: np.ndarray, steps: int, warmup: int, world: int, local: int,
                 gather_mode: str, sustain_s: float = 0.0, ragged: bool = True):
    """Returns a dict: device-timed seconds for `steps` steps (max over ranks), e2e seconds, launches, clocks, pooled
    result of the e2e call, sustained clips/s over >= sustain_s seconds, gather identity check."""
    import torch
    import torch.distributed as dist

    B, ld = clips_np.shape
    L1, D = eng.layers + 1, eng.hidden
    dev = torch.device(f"cuda:{local}")
    pin_in = torch.from_numpy(clips_np).pin_memory()
    audio = pin_in.to(dev)
    pin_out = torch.empty((B, L1, D), dtype=torch.float32).pin_memory()
    run = StepRunner(eng, audio, n_samples, world, dev, gather_mode)
    res = {}

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    for _ in range(warmup):
        run.step()
    run.drain()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        last = run.step()
    run.drain()
    ev1.record()
    barrier()
    res["clocks"] = sampler.stop()
    res["launches"] = eng.launch_count - l0
    res["dev_s"] = max_over_ranks([ev0.elapsed_time(ev1) / 1e3])[0]

    # multi-GPU identity: every rank finds its own shard, bit for bit, at its place in the gathered tensor; rank 0
    # recomputes rank 1's whole shard (same clips, same batch positions) and compares it with what arrived
    if world > 1:
        rank = dist.get_rank()
        mine = run.gath[last][rank * B:(rank + 1) * B]
        ok = torch.equal(mine, run.outs[last])
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        res["gather_ok"] = bool(flag.item() == 1.0)
        res["gathered"] = run.gath[last]

    # sustained: the same step back to back for >= sustain_s seconds (the 20-step region above is < 1 s)
    if sustain_s > 0:
        n_sus = max(steps, int(np.ceil(sustain_s / max(res["dev_s"] / steps, 1e-6))))
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sus = ClockSampler(local)
        sus.start()
        s0.record()
        for _ in range(n_sus):
            run.step()
        run.drain()
        s1.record()
        barrier()
        res["sustained_clocks"] = sus.stop()
        res["sustained_s"] = max_over_ranks([s0.elapsed_time(s1) / 1e3])[0]
        res["sustained_steps"] = n_sus

    # end to end through the host-buffer C-ABI call: pinned H2D + run + D2H every step
    for _ in range(2):
        eng.pooled_pinned(pin_in, n_samples, pin_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.pooled_pinned(pin_in, n_samples, pin_out)
    e2e_local = time.perf_counter() - t0
    res["pooled"] = pin_out.numpy().copy()
    # the same work through the streaming API (host buffers in and out every step; copies overlap the neighbouring
    # steps' compute on separate CUDA streams)
    for _ in eng.pooled_stream((pin_in, n_samples) for _ in range(2)):
        pass
    barrier()
    t1 = time.perf_counter()
    for _ in eng.pooled_stream((pin_in, n_samples) for _ in range(steps)):
        pass
    piped_local = time.perf_counter() - t1
    # ragged streaming: every batch carries NEW per-clip lengths (2 s .. full), so every step uploads lengths
    ragged_local, ragged_clips = 0.0, 0
    if ragged:
        rng = np.random.default_rng(7)
        lens = [rng.integers(ld * 2 // 3, ld + 1, size=B).astype(np.int32) for _ in range(steps + 2)]
        for _ in eng.pooled_stream((pin_in, lens[i]) for i in range(2)):
            pass
        barrier()
        t2 = time.perf_counter()
        for _ in eng.pooled_stream((pin_in, lens[2 + i]) for i in range(steps)):
            pass
        ragged_local = time.perf_counter() - t2
        ragged_clips = B * steps
    res["e2e_s"], res["piped_s"], res["ragged_s"] = max_over_ranks([e2e_local, piped_local, ragged_local])
    res["ragged_clips"] = ragged_clips
    return res


def profile_engine(eng, clips_np, n_samples, local, reps=2):
    import torch

    dev = torch.device(f"cuda:{local}")
    audio = torch.from_numpy(clips_np).to(dev)
    eng.pooled_device(audio, n_samples)
    torch.cuda.synchronize(dev)
    eng.set_option("profile", 1)
    for _ in range(reps):
        eng.pooled_device(audio, n_samples)
    prof = eng.profile_fetch()
    eng.set_option("profile", 0)
    for v in prof.values():
        v["ms"] /= reps
        v["launches"] //= reps
        v["flops"] /= reps
    return prof


def roofline_from_profile(prof: dict, peaks: dict) -> tuple[dict, dict]:
    # the dominant kernel is gemm_tc2_kernel: every launch of it in one step (the fused conv + LayerNorm kernel, the
    # positional-conv kernel and the decoder's token GEMMs are other kernels and are listed in kernels_ms_per_step)
    dominant = ("gemm_qkv", "gemm_out", "gemm_ffn1", "gemm_ffn2", "gemm_proj", "gemm_conv", "gemm_conv1", "gemm_conv2")
    gemm = {k: v for k, v in prof.items() if k in dominant}
    ms = sum(v["ms"] for v in gemm.values())
    fl = sum(v["flops"] for v in gemm.values())
    n = sum(v["launches"] for v in gemm.values())
    total_ms = sum(v["ms"] for v in prof.values())
    ach = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    peak = peaks["bf16_sustained"] or 1400.0
    traffic, tnote = None, None
    for tag in ("r02", "r01"):
        tp = os.path.join(ROOT, "profiles", f"{tag}_gemm_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("traffic_bytes_per_launch_avg")
            tnote = f"dram read+write bytes per launch, ncu --set full average over one layer's four GEMMs " \
                    f"(profiles/{tag}_gemm_traffic.json)"
            break
    roof = {"bound": "tensor", "kernel": "gemm_tc2_kernel (tcgen05 bf16 CTA-pair GEMM: every qkv / out-proj / ffn1 / ffn2 / "
                                         "projection launch of one step)",
            "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4),
            "peak_source": peaks["source"] + ", sustained bf16", "launches_per_step": n,
            "avg_launch_ms": round(ms / max(n, 1), 4), "share_of_step": round(ms / max(total_ms, 1e-9), 4),
            "traffic": traffic, "traffic_note": tnote}
    kern = {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1) if v["flops"] else None}
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    return roof, kern


def hbm_entry(kern: dict, name: str, bytes_per_step: float, peaks: dict, what: str):
    """Roofline entry of a memory-bound kernel group from the in-run event profile."""
    if name not in kern or not kern[name]["ms"]:
        return None
    gbs = bytes_per_step / (kern[name]["ms"] * 1e-3) / 1e9
    peak = peaks["hbm_gbs"] or 6650.0
    return {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4),
            "algorithmic_bytes_per_step": int(bytes_per_step), "ms_per_step": kern[name]["ms"], "what": what}


def hf_gpu_baseline(model, prep, x_np, steps: int, local: int) -> dict:
    """The stock HF module on the same GPU (what the reference's `--device cuda` gives a user,
    REF/WavLM_embeddings.py:442-445, :483): fp32 as the reference runs it, and bf16; same batch, same pooled output;
    the CPU-side feature extraction is outside the timed region (favours this baseline)."""
    import torch

    import copy

    dev = torch.device(f"cuda:{local}")
    out = {}
    x0 = prep(torch.from_numpy(x_np).to(dev))
    model = copy.deepcopy(model)  # the caller's fp32 weights must survive the bf16 round trip below
    for name, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        try:
            m = model.to(dev).to(dtype).eval()
            x = x0.to(dtype)

            def step():
                with torch.no_grad():
                    hs = m(x, output_hidden_states=True, return_dict=True).hidden_states
                    return torch.stack([h.float().mean(1) for h in hs], 1)

            step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"clips_per_s": round(x_np.shape[0] / ms * 1e3, 1), "ms_per_step": round(ms, 2), "steps": steps}
        except Exception as e:  # noqa: BLE001 - e.g. out of memory: reported, not fatal
            out[name] = {"error": str(e)[:160]}
        torch.cuda.empty_cache()
    del model
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="WavLM clips per GPU per step")
    ap.add_argument("--whisper", default="on", choices=["on", "off"],
                    help="also measure Whisper-large (the metric names both models)")
    ap.add_argument("--whisper-batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-HF-on-this-GPU baseline")
    ap.add_argument("--gather", default="inline", choices=["overlap", "inline", "ce"],
                    help="N > 1: all-gather on the compute stream (default), or on a side stream overlapping the next "
                         "step (measured: +1.3 %% at N=2, -1.7 %% at N=4, -1.1 %% at N=8: the NCCL CTAs that run next to "
                         "the forward take SMs away from its persistent 148-CTA kernels), or `ce`: overlapped peer "
                         "copies out of symmetric memory on the copy engines")
    ap.add_argument("--sustain", type=float, default=5.0, help="seconds of the back-to-back sustained loop (0 = skip)")
    ap.add_argument("--ref-clips", type=int, default=8, help="--impl reference: clips per step")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE",
                    help="A/B measurements: process-wide kernel knob passed to ssr_tuning_set (e.g. pdl=0)")
    args = ap.parse_args()
    rank, local, world = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    import torch

    from ssr_b200 import synth

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference(args)
        return

    import torch.distributed as dist

    import ssr_b200
    from ssr_b200 import WavLMEngine, WhisperEncoderEngine

    torch.cuda.set_device(local)
    if world > 1:
        if args.gather == "overlap":
            # The gather has a whole step (~35 ms) to land ~26 MB per peer: two channels are plenty, and every NCCL
            # CTA that runs next to the forward takes an SM away from a persistent 148-CTA kernel for that long.
            os.environ.setdefault("NCCL_MAX_NCHANNELS", "2")
            os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    peaks = load_peaks()
    for kv in args.tune:
        from ssr_b200 import _lib

        key, _, val = kv.partition("=")
        if _lib.load().ssr_tuning_set(key.encode(), int(val)) != 0:
            raise SystemExit(f"unknown tuning knob {key!r}")
    side_legs = rank == 0 and world == 1  # CPU / library baselines and parity: N = 1 only

    # ---- headline: WavLM-Large, B clips of 3 s per GPU (BASELINE.json configs[1]) ----
    model, fe = synth.build_wavlm("large", seed=0)
    eng = WavLMEngine.from_hf(model, fe, device=local)
    B = args.batch
    clips = np.stack([synth.clip_by_index(rank * B + i, 48000) for i in range(B)])
    n_samples = np.full(B, 48000, np.int32)
    r = bench_engine(eng, clips, n_samples, args.steps, warmup, world, local, args.gather, sustain_s=args.sustain)
    pooled = r["pooled"]
    total_clips = world * B * args.steps
    value = total_clips / r["dev_s"]
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": round(r["dev_s"] / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(B, world),
        "clocks": r["clocks"],
        "e2e": {"value": round(total_clips / r["e2e_s"], 1), "unit": "clips/s",
                "h2d_bytes_per_step": int(clips.nbytes + n_samples.nbytes), "d2h_bytes_per_step": int(pooled.nbytes),
                "api": "ssr_wavlm_pooled_host (C ABI, pinned host buffers)",
                "streamed_value": round(total_clips / r["piped_s"], 1),
                "streamed_api": "WavLMEngine.pooled_stream: same host buffers and copies every step, H2D / D2H on side "
                                "streams overlapping the neighbouring steps' ssr_wavlm_pooled calls",
                "streamed_ragged_value": round(world * r["ragged_clips"] / r["ragged_s"], 1) if r["ragged_s"] else None,
                "streamed_ragged_note": "pooled_stream with new random per-clip lengths (2 s .. 3 s) every batch: "
                                        "a length upload through the pinned ring every step"},
        "gpu_launches": int(r["launches"]),
    }
    if args.tune:
        line["tuning"] = args.tune
    if args.sustain > 0:
        line["sustained_value"] = round(world * B * r["sustained_steps"] / r["sustained_s"], 1)
        line["sustained"] = {"seconds": round(r["sustained_s"], 2), "steps": r["sustained_steps"],
                             "clocks": r["sustained_clocks"]}
    if world > 1:
        line["gather"] = {"mode": args.gather, "nccl_max_nchannels": os.environ.get("NCCL_MAX_NCHANNELS"),
                          "gather_ok": r["gather_ok"],
                          "bytes_landed_per_rank_per_step": int(world * pooled.nbytes)}
        if rank == 0:
            # rank 0 recomputes rank 1's shard (same clips, same batch positions) and compares with what arrived
            other = np.stack([synth.clip_by_index(1 * B + i, 48000) for i in range(B)])
            again = eng.pooled_device(torch.from_numpy(other).to(f"cuda:{local}"), n_samples)
            line["gather"]["recompute_of_rank1_shard_bit_identical"] = bool(torch.equal(again, r["gathered"][B:2 * B]))
        r.pop("gathered", None)
    if rank == 0:
        prof = profile_engine(eng, clips, n_samples, local)
        roof, kern = roofline_from_profile(prof, peaks)
        line["roofline"] = roof
        line["kernels_ms_per_step"] = kern
        line["model_tflops"] = round(value / world * WAVLM_GFLOP_PER_CLIP / 1e3, 1)
        line["model_frac_of_bf16_sustained"] = round(line["model_tflops"] / (peaks["bf16_sustained"] or 1400.0), 4)
        # memory-bound stages: bf16 conv0 output written once; every LayerNorm reads fp32 rows and writes bf16 rows
        M = B * 149
        rl = {}
        ent = hbm_entry(kern, "wavlm_conv0", B * (48000 * 4 + 9599 * 512 * 2), peaks,
                        "waveform read + [B, 9599, 512] bf16 written once")
        if ent:
            rl["wavlm_conv0"] = ent
        ent = hbm_entry(kern, "layernorm", 48 * M * 1024 * (4 + 2) + M * 512 * 4 + M * 1024 * 8, peaks,
                        "48 pre-LN rows passes (fp32 in, bf16 out) + projection LN + final LN")
        if ent:
            rl["layernorm"] = ent
        line["roofline_memory_bound"] = rl

    if side_legs:
        # the reference-shaped per-clip call (B = 1, host numpy in, dict out): what swapping only the import gives
        c0 = clips[0]
        for _ in range(3):
            ssr_b200.extract_embeddings_from_audio_wavlm(c0, model, fe, f"cuda:{local}", [24, 23, 22, 12])
        t0 = time.perf_counter()
        n_pc = 100
        for i in range(n_pc):
            ssr_b200.extract_embeddings_from_audio_wavlm(clips[i % B], model, fe, f"cuda:{local}", [24, 23, 22, 12])
        line["percall_ms"] = round((time.perf_counter() - t0) / n_pc * 1e3, 3)
        line["percall_note"] = "extract_embeddings_from_audio_wavlm, one 3 s clip per call (the reference's loop), " \
                               "host numpy in, dict of numpy out"

    # ---- CPU baseline: the reference's per-clip loop on this box's host cores (rank 0, N = 1 only) ----
    if side_legs and not args.no_cpu_baseline:
        from oracle import ref_port

        torch.set_num_threads(os.cpu_count() or 1)
        idx = list(range(model.config.num_hidden_layers + 1))
        ref_port.wavlm_extract_one(clips[0], model, fe, idx)  # warm-up
        t0 = time.perf_counter()
        ref_out, n_done = [], 0
        while n_done < min(B, 64) and (time.perf_counter() - t0 < 12.0 or n_done < 4):
            emb = ref_port.wavlm_extract_one(clips[n_done], model, fe, idx)
            ref_out.append(np.stack([emb[f"layer_{i}"] for i in idx]))
            n_done += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": round(n_done / dt, 3), "unit": "clips/s", "cores": torch.get_num_threads(),
                                "kind": "port",
                                "sample": f"first {n_done} clips of the same batch, one clip per HF forward "
                                          "(the reference's loop), fp32, all 25 layers pooled"}
        line["parity"] = parity(pooled[:n_done], np.stack(ref_out))
    if side_legs and not args.no_gpu_baseline:
        def prep(x):
            return (x - x.mean(1, keepdim=True)) / torch.sqrt(x.var(1, unbiased=False, keepdim=True) + 1e-7)

        del eng
        torch.cuda.empty_cache()
        eng = None
        line["gpu_library_baseline"] = hf_gpu_baseline(model, prep, clips, 3, local)
        line["gpu_library_baseline"]["note"] = "stock HF WavLMModel.cuda() on this GPU, same batch, same pooled output"

    # ---- Whisper-large encoder, 30 s window (BASELINE configs[2]); measured at every N ----
    if args.whisper == "on":
        eng = None
        model = None
        torch.cuda.empty_cache()
        enc, wfe = synth.build_whisper_encoder("large", seed=0)
        weng = WhisperEncoderEngine.from_hf(enc, wfe, device=local)
        WB = args.whisper_batch
        wclips = np.stack([synth.clip_by_index(rank * WB + i, 48000) for i in range(WB)])
        wn = np.full(WB, 48000, np.int32)
        wsteps = max(2, args.steps // 3)
        wr = bench_engine(weng, wclips, wn, wsteps, 3, world, local, args.gather, sustain_s=args.sustain,
                          ragged=False)
        wpooled = wr["pooled"]
        sec = {"workload": f"Whisper-large encoder (32 layers, d=1280), 3 s clips in the 30 s window, batch {WB} per "
                           "GPU (BASELINE configs[2])",
               "value": round(world * WB * wsteps / wr["dev_s"], 2), "unit": "clips/s", "n_gpus": world,
               "ms_per_step": round(wr["dev_s"] / wsteps * 1e3, 2), "steps": wsteps,
               "e2e": round(world * WB * wsteps / wr["e2e_s"], 2),
               "streamed_value": round(world * WB * wsteps / wr["piped_s"], 2),
               "gpu_launches": int(wr["launches"]), "clocks": wr["clocks"]}
        if args.sustain > 0:
            sec["sustained_value"] = round(world * WB * wr["sustained_steps"] / wr["sustained_s"], 2)
            sec["sustained"] = {"seconds": round(wr["sustained_s"], 2), "steps": wr["sustained_steps"],
                                "clocks": wr["sustained_clocks"]}
        if world > 1:
            sec["gather_ok"] = wr["gather_ok"]
            wr.pop("gathered", None)
        if rank == 0:
            wprof = profile_engine(weng, wclips, wn, local, reps=1)
            wroof, wkern = roofline_from_profile(wprof, peaks)
            sec["roofline"] = wroof
            sec["kernels_ms_per_step"] = wkern
            sec["model_tflops"] = round(sec["value"] / world * WHISPER_GFLOP_PER_CLIP / 1e3, 1)
            sec["model_frac_of_bf16_sustained"] = round(sec["model_tflops"] / (peaks["bf16_sustained"] or 1400.0), 4)
        if rank == 0:
            # BASELINE configs[2], second half: 64 FULL-LENGTH (480 000-sample) clips — only the log-mel stage sees the
            # difference (every frame is live). Device-timed, same step otherwise.
            fclips = np.stack([synth.clip_by_index(rank * WB + i, 480000) for i in range(WB)])
            fn = np.full(WB, 480000, np.int32)
            fdev = torch.from_numpy(fclips).to(f"cuda:{local}")
            fout = None
            for _ in range(2):
                fout = weng.pooled_device(fdev, fn, out=fout)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(wsteps):
                weng.pooled_device(fdev, fn, out=fout)
            e1.record()
            torch.cuda.synchronize()
            fms = e0.elapsed_time(e1) / wsteps
            fprof = profile_engine(weng, fclips, fn, local, reps=1)
            _, fkern = roofline_from_profile(fprof, peaks)
            sec["full_length"] = {"workload": f"{WB} clips of 480 000 samples (the whole 30 s window live)",
                                  "value": round(WB / fms * 1e3, 2), "unit": "clips/s (1 GPU)",
                                  "ms_per_step": round(fms, 2),
                                  "logmel_roofline": hbm_entry(
                                      fkern, "logmel", WB * (480000 * 4 + 3000 * 80 * 2), peaks,
                                      "480 000 f32 samples read + [3000, 80] bf16 log-mel written, per clip")}
            sec["logmel_roofline"] = hbm_entry(wkern, "logmel", WB * (48000 * 4 + 3000 * 80 * 2), peaks,
                                               "48 000 f32 samples read + [3000, 80] bf16 log-mel written, per clip "
                                               "(frames in the zero padding are synthesised, not read)")
            del fdev, fout
        if side_legs and not args.no_cpu_baseline:
            from oracle import ref_port

            idx = list(range(enc.config.encoder_layers + 1))
            t0 = time.perf_counter()
            refs = []
            for i in range(2):
                emb = ref_port.whisper_extract_one(wclips[i], enc, wfe, idx)
                refs.append(np.stack([emb[f"encoder_layer_{j}"] for j in idx]))
            dt = time.perf_counter() - t0
            sec["cpu_baseline"] = {"value": round(2 / dt, 4), "unit": "clips/s", "cores": torch.get_num_threads(),
                                   "kind": "port", "sample": "first 2 clips, one clip per HF encoder forward, fp32"}
            sec["parity"] = parity(wpooled[:2], np.stack(refs))
        if side_legs and not args.no_gpu_baseline:
            del weng
            torch.cuda.empty_cache()
            feats = wfe([c for c in wclips], sampling_rate=16000, return_tensors="np").input_features
            sec["gpu_library_baseline"] = hf_gpu_baseline(enc, lambda x: x, np.asarray(feats, np.float32), 2, local)
            sec["gpu_library_baseline"]["note"] = "stock HF WhisperEncoder.cuda() on this GPU, same batch, log-mel " \
                                                  "features precomputed on the CPU outside the timed region"
        line["whisper_large"] = sec

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def run_reference(args):
    """The reference's own CPU implementation of the path (port over HF transformers; /root/reference itself is
    Python glue that is not present on the GPU box), all host threads, bounded sample per step."""
    import torch

    from oracle import ref_port
    from ssr_b200 import synth

    torch.set_num_threads(os.cpu_count() or 1)
    model, fe = synth.build_wavlm("large", seed=0)
    idx = list(range(model.config.num_hidden_layers + 1))
    n = args.ref_clips
    clips = [synth.clip_by_index(i, 48000) for i in range(n)]
    for _ in range(max(args.warmup, 1)):
        ref_port.time_clips(ref_port.wavlm_extract_one, clips[:2], model, fe, idx)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref_port.time_clips(ref_port.wavlm_extract_one, clips, model, fe, idx)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    cb = {"value": round(value, 3), "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
          "sample": f"{n} clips of that workload per step, one clip per HF forward on the host cores, fp32 (the "
                    "reference's own loop, REF/WavLM_embeddings.py:583-594)"}
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.batch, args.gpus),
            "cpu_baseline": cb,
            "e2e": {"value": round(value, 3), "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.whisper == "on":
        del model
        enc, wfe = synth.build_whisper_encoder("large", seed=0)
        widx = list(range(enc.config.encoder_layers + 1))
        ref_port.whisper_extract_one(clips[0], enc, wfe, widx)
        t0 = time.perf_counter()
        nw = 2
        for i in range(nw):
            ref_port.whisper_extract_one(clips[i], enc, wfe, widx)
        wdt = time.perf_counter() - t0
        line["whisper_large"] = {"value": round(nw / wdt, 4), "unit": "clips/s", "cores": torch.get_num_threads(),
                                 "kind": "port", "sample": f"{nw} clips, one clip per HF encoder forward, fp32"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
