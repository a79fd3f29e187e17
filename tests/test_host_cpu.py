"""CPU tier: the C-ABI library loads and exports every declared symbol; host-side logic (layer selection, sharding,
error convention, WAV loader); world_size-2 gloo test of the sharded extraction plumbing. No compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ssr_b200 import _lib

    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "ssr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ssr_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ssr_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_abi_argument_validation_without_gpu():
    from ssr_b200 import _lib

    lib = _lib.load()
    h = C.c_void_p()
    assert lib.ssr_create(None, None, 0, 0, C.byref(h)) != 0
    assert b"null" in lib.ssr_last_error(None)
    assert lib.ssr_set_option(None, b"simt_gemm", 1) == -1
    assert lib.ssr_launch_count(None) == -1
    assert lib.ssr_num_frames(None, 48000) == -1
    lib.ssr_destroy(None)  # must be a no-op


def test_struct_layout_matches_header():
    from ssr_b200 import _lib

    assert C.sizeof(_lib.ModelDesc) == 16 * 4
    assert C.sizeof(_lib.Weight) == 24


def test_engine_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ssr_b200 import SsrError, WavLMEngine, synth

    model, fe = synth.build_wavlm("tiny_post")
    with pytest.raises(SsrError):
        WavLMEngine.from_hf(model, fe)


def test_dropin_returns_none_instead_of_raising(caplog):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present: covered by the gpu tier")
    import ssr_b200
    from ssr_b200 import synth

    model, fe = synth.build_wavlm("tiny_post")
    out = ssr_b200.extract_embeddings_from_audio_wavlm(np.zeros(16000, np.float32), model, fe, "cpu", [2, 1])
    assert out is None  # reference convention: log + None (REF/WavLM_embeddings.py:329-341)


def test_layer_selection_semantics():
    from ssr_b200 import pooled_to_layer_dict

    pooled = np.arange(13 * 4, dtype=np.float32).reshape(13, 4)
    n = 13
    idx = [n - 1, n - 2, n - 3, n // 2]  # REF/WavLM_embeddings.py:506
    out = pooled_to_layer_dict(pooled, idx + [13, 40], "layer_")
    assert list(out) == ["layer_12", "layer_11", "layer_10", "layer_6"]
    assert out["layer_6"].dtype == np.float32 and out["layer_6"].shape == (4,)
    np.testing.assert_array_equal(out["layer_12"], pooled[12])


def test_shard_range_partitions_contiguously():
    from ssr_b200 import iter_batches, shard_range

    for n in (0, 1, 7, 100000):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert list(iter_batches(10, 4)) == [(0, 4), (4, 8), (8, 10)]


def test_clip_by_index_is_sharding_independent():
    from ssr_b200 import synth

    a = synth.clip_by_index(12345)
    b = synth.clip_by_index(12345)
    np.testing.assert_array_equal(a, b)
    assert not np.array_equal(a, synth.clip_by_index(12346))


def test_wav_loader(tmp_path):
    from ssr_b200.extract import load_audio

    x = (np.sin(np.arange(1600) / 10.0) * 20000).astype("<i2")
    p = tmp_path / "a.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(np.stack([x, x], 1).tobytes())
    y = load_audio(p)
    assert y.shape == (1600,) and y.dtype == np.float32
    np.testing.assert_allclose(y, x / 32768.0, atol=1e-6)
    assert load_audio(p, max_length=0.05).shape == (800,)
    assert load_audio(tmp_path / "missing.wav") is None


def test_wav_loader_float_and_resample(tmp_path):
    """The reference loads ANY sample rate / format torchaudio reads and resamples to 16 kHz
    (REF/WavLM_embeddings.py:87-125). Here: IEEE-float WAV (which the `wave` module refuses) and an 8 kHz file that
    must come back resampled by torchaudio's Resample exactly as the reference would do it."""
    import struct

    import torch
    import torchaudio

    from ssr_b200.extract import load_audio

    x = (0.5 * np.sin(np.arange(4000) / 7.0)).astype("<f4")
    p = tmp_path / "f32.wav"
    hdr = b"RIFF" + struct.pack("<I", 36 + x.nbytes) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 3, 1, 16000,
                                                                                         64000, 4, 32)
    p.write_bytes(hdr + b"data" + struct.pack("<I", x.nbytes) + x.tobytes())
    y = load_audio(p)
    assert y is not None and y.dtype == np.float32
    np.testing.assert_array_equal(y, x)

    q = tmp_path / "8k.wav"
    xi = (np.sin(np.arange(8000) / 9.0) * 12000).astype("<i2")
    with wave.open(str(q), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(8000)
        w.writeframes(xi.tobytes())
    y8 = load_audio(q)
    want = torchaudio.transforms.Resample(8000, 16000)(torch.from_numpy(xi.astype(np.float32) / 32768.0)[None])
    assert y8 is not None and y8.shape == (16000,)
    np.testing.assert_allclose(y8, want.squeeze().numpy(), atol=1e-6)
    bad = tmp_path / "bad.flac"
    bad.write_bytes(b"fLaC" + bytes(64))
    assert load_audio(bad) is None  # undecodable here: logged and skipped, like the reference


def test_bench_config_identical_for_both_arms_and_parity_vectors():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.workload_config(256, 4) == bench.workload_config(256, 4)
    assert set(bench.workload_config(256, 1)) >= {"workload", "clips_per_gpu_per_step", "parallelism"}
    ref = np.random.default_rng(0).standard_normal((3, 5, 16)).astype(np.float32)
    got = ref.copy()
    got[1, 3] *= 1.02
    par = bench.parity(got, ref)
    assert len(par["max_rel_err_per_layer"]) == 5 and len(par["min_cos_per_layer"]) == 5
    assert par["max_rel_err_per_layer"][3] == pytest.approx(0.02, rel=1e-3) and not par["ok"]
    assert par["max_rel_err_per_layer"][0] == 0.0


GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SSR_ROOT"])
from ssr_b200 import shard_range, synth
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, L1, D = 11, 3, 8
lo, hi = shard_range(N, rank, world)
# stand-in for the engine: a deterministic function of the clip index, so the gather order is checkable
def fake_pooled(i):
    c = synth.clip_by_index(i, 64)
    return np.outer(np.arange(1, L1 + 1), c[:D]).astype(np.float32)
local = np.stack([fake_pooled(i) for i in range(lo, hi)]) if hi > lo else np.zeros((0, L1, D), np.float32)
sizes = [shard_range(N, r, world) for r in range(world)]
maxn = max(h - l for l, h in sizes)
pad = np.zeros((maxn, L1, D), np.float32); pad[: hi - lo] = local
out = [torch.zeros((maxn, L1, D)) for _ in range(world)]
dist.all_gather(out, torch.from_numpy(pad))
full = np.concatenate([o.numpy()[: h - l] for o, (l, h) in zip(out, sizes)])
want = np.stack([fake_pooled(i) for i in range(N)])
assert np.array_equal(full, want), "gathered shards are not in global clip order"
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharded_gather_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, SSR_ROOT=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29653", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_pipeline_on_disk_contract_and_resume(tmp_path):
    """files the reference's training scripts read (REF/model_training_1.py:112-161)."""
    import pickle

    from ssr_b200 import pipeline

    clips = {f"/data/train_{i}.wav": np.full(100 + i, float(i), np.float32) for i in range(7)}
    clips["/data/train_3.wav"] = None  # a clip that fails to load is skipped, as in REF/WavLM_embeddings.py:596-598
    rows = [{"filename": os.path.basename(p), "path": p, "label": i % 2, "split": "train"}
            for i, p in enumerate(clips)]
    calls = []

    def pooled_fn(batch):  # stand-in engine: [B, 3, 4], value = clip's first sample
        calls.append(len(batch))
        return np.stack([np.full((3, 4), c[0], np.float32) + np.arange(3, dtype=np.float32)[:, None] for c in batch])

    out = str(tmp_path)
    res = pipeline.extract_split(rows, pooled_fn, [2, 1, 9], out, "train", clips.get, batch_size=4,
                                 checkpoint_interval=4)
    assert calls == [3, 3] and len(res) == 6
    meta, emb = pipeline.load_split(out, "train")
    assert list(meta.columns) == ["filename", "path", "label", "split"] and len(meta) == 6
    assert sorted(emb) == ["layer_1", "layer_2"]
    assert emb["layer_2"].dtype == np.float32 and emb["layer_2"].shape == (6, 4)
    order = [0, 1, 2, 4, 5, 6]
    np.testing.assert_array_equal(emb["layer_1"][:, 0], np.array(order, np.float32) + 1)
    assert list(meta["path"]) == [f"/data/train_{i}.wav" for i in order]
    ck = sorted(os.listdir(os.path.join(out, "checkpoints")))
    assert ck == ["checkpoint_train_0.pkl", "checkpoint_train_1.pkl"]
    with open(os.path.join(out, "checkpoints", ck[-1]), "rb") as f:
        saved = pickle.load(f)
    assert len(saved) == 6 and set(saved[0]) == {"filename", "path", "label", "split", "layer_2", "layer_1"}
    # resume: already processed paths are filtered out, nothing new is computed except the new file
    clips["/data/train_7.wav"] = np.full(50, 7.0, np.float32)
    rows.append({"filename": "train_7.wav", "path": "/data/train_7.wav", "label": 1, "split": "train"})
    clips["/data/train_3.wav"] = None
    calls.clear()
    res2 = pipeline.extract_split(rows, pooled_fn, [2, 1], out, "train", clips.get, batch_size=4, resume=True)
    assert calls == [1] and len(res2) == 7
    assert pipeline.find_latest_checkpoint(out, "train") == 2


def test_product_bucket_function_bit_exact_vs_hf():
    """A6 is the one piece of integer work on the path: the PRODUCT's own C function (csrc/engine.cu rel_bucket,
    exported as ssr_wavlm_rel_bucket; it fills the device relative-bias table) against HF
    WavLMAttention._relative_positions_bucket (modeling_wavlm.py:252-271) for every rel in [-4000, 4000], and
    against the known answers of SURVEY.md 8(c). Bit-exact, no tolerance."""
    import hashlib

    import torch
    from transformers.models.wavlm.modeling_wavlm import WavLMAttention

    from ssr_b200 import _lib

    lib = _lib.load()
    att = WavLMAttention(embed_dim=64, num_heads=1)
    rel = torch.arange(-4000, 4001)
    want = att._relative_positions_bucket(rel).numpy()
    got = np.array([lib.ssr_wavlm_rel_bucket(int(r)) for r in rel.tolist()], dtype=want.dtype)
    np.testing.assert_array_equal(got, want)
    kat = {1: 161, -1: 1, 79: 239, -79: 79, 80: 240, 81: 240, -81: 80, 100: 247, -100: 87, 120: 254, -120: 94,
           148: 261, -148: 101, 0: 0}
    for r, b in kat.items():
        assert lib.ssr_wavlm_rel_bucket(r) == b, r
    lut = np.array([lib.ssr_wavlm_rel_bucket(r) for r in range(-148, 149)], dtype=np.int16)
    assert hashlib.sha256(lut.tobytes()).hexdigest().startswith("67859a4d9be1a425")


_CURSOR_HARNESS = r"""
#include <cstdio>
#include <cstdlib>
#include <vector>
#define __device__
#define __forceinline__ inline
struct { int x; } blockIdx;
struct AttentionArgs { int B, slot, H, D; };
constexpr int QT = 128;
struct ItemPre { int b, h, q0; };
%s
static int ceil_div(int a, int b) { return (a + b - 1) / b; }
int main(int argc, char** argv) {
  AttentionArgs a;
  a.B = atoi(argv[1]); a.slot = atoi(argv[2]); a.H = atoi(argv[3]); a.D = a.H * 64;
  const int sms = atoi(argv[4]), paired = atoi(argv[5]), grouped = atoi(argv[6]);
  const int nqt = ceil_div(a.slot, QT);
  const long long items = (long long)nqt * a.H * a.B;
  const int grid = (int)(items < 2LL * sms ? items : 2LL * sms);
  const Step st = make_step(a, grid, paired, grouped, 0);
  std::vector<int> seen(items, 0);
  long long adjacent = 0;  // items of one (clip, head) that run in the same round on consecutive CTAs
  for (int c = 0; c < grid; ++c) {
    blockIdx.x = c;
    Cursor cur;
    cur.init(a, st);
    for (long long idx = c; idx < items; idx += grid) {
      if (cur.qt < 0 || cur.qt >= nqt || cur.b < 0 || cur.b >= a.B || cur.h < 0 || cur.h >= a.H) {
        printf("out of range: cta %%d idx %%lld -> qt %%d b %%d h %%d\n", c, idx, cur.qt, cur.b, cur.h);
        return 1;
      }
      ++seen[((long long)cur.b * a.H + cur.h) * nqt + cur.qt];
      if (idx + grid < items) cur.advance(a, st);
    }
  }
  for (long long i = 0; i < items; ++i)
    if (seen[i] != 1) { printf("item %%lld visited %%d times\n", i, seen[i]); return 1; }
  printf("ok mode %%d\n", st.paired ? 1 : st.grouped ? 2 : 0);
  return 0;
}
"""


def test_attention_item_orders_visit_every_item_once(tmp_path):
    """The persistent attention CTAs walk their items with the mixed-radix Cursor of csrc/attention_tc.cu (no division
    per item). Its three orders (query-tile-major, paired, grouped) are compiled for the host from the product's own
    source text and must each visit every (clip, head, query tile) exactly once, for grids smaller and larger than
    the item count."""
    src = open(os.path.join(ROOT, "stuttering-speech-representation_b200", "csrc", "attention_tc.cu")).read()
    m = re.search(r"struct Step \{.*?\n\};\n", src, re.S)
    c0 = src.index("struct Cursor {")
    c1 = src.index("  // The clip length of the item goes global", c0)
    f0 = src.index("static Step make_step(")
    f1 = src.index("\n}\n", f0) + 3
    code = _CURSOR_HARNESS % (m.group(0) + src[c0:c1] + "};\n" + src[f0:f1])
    (tmp_path / "cursor.cpp").write_text(code)
    exe = str(tmp_path / "cursor")
    subprocess.run(["g++", "-O1", "-o", exe, str(tmp_path / "cursor.cpp")], check=True)
    modes = set()
    for B, slot, H in [(64, 1500, 20), (1, 1500, 20), (3, 1500, 6), (5, 300, 4), (7, 385, 3), (256, 150, 16),
                       (9, 256, 16), (40, 128, 16), (2, 150, 2), (33, 640, 12), (1, 1500, 1)]:
        for sms in (148, 7, 1):
            for paired, grouped in [(0, 0), (1, 0), (0, 1), (1, 1)]:
                r = subprocess.run([exe, str(B), str(slot), str(H), str(sms), str(paired), str(grouped)],
                                   capture_output=True, text=True)
                assert r.returncode == 0, (B, slot, H, sms, paired, grouped, r.stdout)
                modes.add(r.stdout.strip())
    assert modes == {"ok mode 0", "ok mode 1", "ok mode 2"}
