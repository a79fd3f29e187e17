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
    g, r = got.astype(np.float64), ref.astype(np.float64)
    cos = (g * r).sum(-1) / np.maximum(np.sqrt((g * g).sum(-1) * (r * r).sum(-1)), 1e-30)
    rel = np.abs(g - r).max(-1) / np.maximum(np.abs(r).max(-1), 1e-30)
    return {"min_cos": float(cos.min()), "max_rel_err": float(rel.max()), "tolerance": "cos>=0.9999, rel<=1e-2",
            "ok": bool(cos.min() >= 0.9999 and rel.max() <= 1e-2)}


# ---------------------------------------------------------------------------------------------- our arm
def bench_engine(eng, clips_np: np.ndarray, n_samples: np.ndarray, steps: int, warmup: int, world: int, local: int,
                 gather: bool):
    """Returns (device-timed seconds for `steps` steps [max over ranks], e2e seconds, launches, clocks, pooled)."""
    import torch
    import torch.distributed as dist

    B, ld = clips_np.shape
    L1, D = eng.layers + 1, eng.hidden
    dev = torch.device(f"cuda:{local}")
    pin_in = torch.from_numpy(clips_np).pin_memory()
    audio = pin_in.to(dev)
    out = torch.empty((B, L1, D), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * B, L1, D), dtype=torch.float32, device=dev) if gather else None
    pin_out = torch.empty((B, L1, D), dtype=torch.float32).pin_memory()

    def step():
        eng.pooled_device(audio, n_samples, out=out)
        if gather:
            dist.all_gather_into_tensor(gathered, out)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    launches = eng.launch_count - l0
    t = torch.tensor([ev0.elapsed_time(ev1) / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s = float(t.item())

    # end to end through the host-buffer C-ABI call: pinned H2D + run + D2H every step
    for _ in range(2):
        eng.pooled_pinned(pin_in, n_samples, pin_out)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.pooled_pinned(pin_in, n_samples, pin_out)
    e2e_local = time.perf_counter() - t0
    result = pin_out.numpy().copy()
    # the same work through the streaming API (host buffers in and out every step; copies overlap the neighbouring
    # steps' compute on separate CUDA streams)
    for _ in eng.pooled_stream((pin_in, n_samples) for _ in range(2)):
        pass
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    for _ in eng.pooled_stream((pin_in, n_samples) for _ in range(steps)):
        pass
    piped_local = time.perf_counter() - t1
    t = torch.tensor([e2e_local, piped_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    bench_engine.pipelined_s = float(t[1].item())
    return dev_s, float(t[0].item()), launches, clocks, result


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
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("traffic_bytes_per_launch_avg")
    roof = {"bound": "tensor", "kernel": "gemm_tc2_kernel (tcgen05 bf16 CTA-pair GEMM: every qkv / out-proj / ffn1 / ffn2 / "
                                         "projection launch of one step)",
            "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4),
            "peak_source": peaks["source"] + ", sustained bf16", "launches_per_step": n,
            "avg_launch_ms": round(ms / max(n, 1), 4), "share_of_step": round(ms / max(total_ms, 1e-9), 4),
            "traffic": traffic,
            "traffic_note": "dram read+write bytes per launch, ncu --set full average over one layer's four GEMMs "
                            "(profiles/r01_gemm_traffic.json)" if traffic else None}
    kern = {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1) if v["flops"] else None}
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    return roof, kern


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="WavLM clips per GPU per step")
    ap.add_argument("--whisper", default="auto", choices=["auto", "on", "off"],
                    help="also measure Whisper-large (secondary); auto = only at --gpus 1")
    ap.add_argument("--whisper-batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-clips", type=int, default=8, help="--impl reference: clips per step")
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

    from ssr_b200 import WavLMEngine, WhisperEncoderEngine

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    peaks = load_peaks()

    # ---- headline: WavLM-Large, B clips of 3 s per GPU (BASELINE.json configs[1]) ----
    model, fe = synth.build_wavlm("large", seed=0)
    eng = WavLMEngine.from_hf(model, fe, device=local)
    B = args.batch
    clips = np.stack([synth.clip_by_index(rank * B + i, 48000) for i in range(B)])
    n_samples = np.full(B, 48000, np.int32)
    dev_s, e2e_s, launches, clocks, pooled = bench_engine(eng, clips, n_samples, args.steps, warmup, world, local,
                                                          gather=world > 1)
    total_clips = world * B * args.steps
    value = total_clips / dev_s
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": round(dev_s / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "WavLM-Large (24 layers, d=1024) per-layer pooled embeddings, 3 s clips, "
                               f"batch {B} per GPU (BASELINE configs[1])",
                   "clips_per_gpu_per_step": B, "clip_samples": 48000, "weights": "seeded random init (seed 0)",
                   "output": f"[{B}, 25, 1024] f32 per GPU", "parallelism": f"clip-sharded x{world}",
                   "l2_policy": "no explicit flush: per-step working set (~6 GB of activations) >> 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": round(total_clips / e2e_s, 1), "unit": "clips/s",
                "h2d_bytes_per_step": int(clips.nbytes + n_samples.nbytes), "d2h_bytes_per_step": int(pooled.nbytes),
                "api": "ssr_wavlm_pooled_host (C ABI, pinned host buffers)",
                "streamed_value": round(total_clips / bench_engine.pipelined_s, 1),
                "streamed_api": "WavLMEngine.pooled_stream: same host buffers and copies every step, H2D / D2H on side "
                                "streams overlapping the neighbouring steps' ssr_wavlm_pooled calls"},
        "gpu_launches": int(launches),
    }
    if rank == 0:
        prof = profile_engine(eng, clips, n_samples, local)
        roof, kern = roofline_from_profile(prof, peaks)
        line["roofline"] = roof
        line["kernels_ms_per_step"] = kern
        line["model_tflops"] = round(value / world * WAVLM_GFLOP_PER_CLIP / 1e3, 1)
        line["model_frac_of_bf16_sustained"] = round(line["model_tflops"] / (peaks["bf16_sustained"] or 1400.0), 4)

    # ---- CPU baseline: the reference's per-clip loop on this box's host cores (rank 0, N = 1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
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

    # ---- secondary: Whisper-large encoder, 30 s window (BASELINE configs[2]) ----
    want_whisper = args.whisper == "on" or (args.whisper == "auto" and world == 1)
    if want_whisper:
        del eng
        torch.cuda.empty_cache()
        enc, wfe = synth.build_whisper_encoder("large", seed=0)
        weng = WhisperEncoderEngine.from_hf(enc, wfe, device=local)
        WB = args.whisper_batch
        wclips = np.stack([synth.clip_by_index(rank * WB + i, 48000) for i in range(WB)])
        wn = np.full(WB, 48000, np.int32)
        wsteps = max(2, args.steps // 3)
        wdev, we2e, wl, wclk, wpooled = bench_engine(weng, wclips, wn, wsteps, 3, world, local, gather=world > 1)
        sec = {"workload": f"Whisper-large encoder (32 layers, d=1280), 3 s clips in the 30 s window, batch {WB}",
               "value": round(world * WB * wsteps / wdev, 2), "unit": "clips/s",
               "ms_per_step": round(wdev / wsteps * 1e3, 2), "steps": wsteps,
               "e2e": round(world * WB * wsteps / we2e, 2), "gpu_launches": int(wl), "clocks": wclk}
        if rank == 0:
            wprof = profile_engine(weng, wclips, wn, local, reps=1)
            wroof, wkern = roofline_from_profile(wprof, peaks)
            sec["roofline"] = wroof
            sec["kernels_ms_per_step"] = wkern
            sec["model_tflops"] = round(sec["value"] / world * WHISPER_GFLOP_PER_CLIP / 1e3, 1)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
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
          "sample": f"{n} clips of 3 s per step, one clip per HF forward (the reference's loop), fp32"}
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "WavLM-Large (24 layers, d=1024) per-layer pooled embeddings, 3 s clips, "
                                   f"batch {args.batch} per GPU (BASELINE configs[1])",
                       "clip_samples": 48000, "weights": "seeded random init (seed 0)",
                       "reference_sample": f"{n} clips of that workload per step, one clip per HF forward on the "
                                           "host cores (the reference's own loop, REF/WavLM_embeddings.py:583-594)"},
            "cpu_baseline": cb,
            "e2e": {"value": round(value, 3), "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
