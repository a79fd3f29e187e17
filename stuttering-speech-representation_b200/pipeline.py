"""Batched split extraction + the reference's on-disk contract (a "next" row of the scope table).

Feeds the GPU engine in real batches instead of the reference's clip-at-a-time loop and writes exactly what the
reference's extraction scripts write, so that the UNMODIFIED training scripts (`load_data`,
REF/model_training_1.py:99-165) can consume the output:

    <output_dir>/<split>/embedding_metadata.csv          all non-embedding columns, row order = processing order
    <output_dir>/<split>/<layer>_embeddings.npy          float32 [N, D], same row order
    <output_dir>/checkpoints/checkpoint_<split>_<n>.pkl  pickled list of result dicts (row metadata + layer arrays)

Mirrors: save_embeddings REF/WavLM_embeddings.py:343-387 (WavLM columns `layer_*`), REF/whisper_embeddings_large.py:301-348
(`encoder_layer_*`); checkpoints REF/WavLM_embeddings.py:389-434; the resume filter (:555-564) and the checkpoint
save condition `((i + batch) % checkpoint_interval == 0) or last` (:633). Failed clips are skipped with a warning, as
the reference does (:596-598). numpy / pandas / pickle only — no GPU code lives here.
"""
from __future__ import annotations

import logging
import os
import pickle
from typing import Callable, Iterable, Sequence

import numpy as np

logger = logging.getLogger("ssr_b200")

EMBED_PREFIXES = ("layer_", "encoder_layer_", "decoder_layer_")


def _is_embedding_col(name: str) -> bool:
    return name.startswith(EMBED_PREFIXES)


def save_embeddings(results: Sequence[dict], output_dir: str, split: str | None = None, expected_dim: int | None = None):
    """Write metadata CSV + one `.npy` per layer for a list of result dicts (the reference passes a DataFrame built
    from the same list, REF/WavLM_embeddings.py:638-647)."""
    import pandas as pd

    if len(results) == 0:
        logger.warning("No embeddings to save")
        return
    split_dir = os.path.join(output_dir, split) if split and split != "all" else output_dir
    os.makedirs(split_dir, exist_ok=True)
    df = pd.DataFrame(list(results))
    meta_cols = [c for c in df.columns if not _is_embedding_col(c)]
    df[meta_cols].to_csv(os.path.join(split_dir, "embedding_metadata.csv"), index=False)
    for col in [c for c in df.columns if _is_embedding_col(c)]:
        arr = np.stack(df[col].values).astype(np.float32)
        if expected_dim is not None and arr.shape[1] != expected_dim:
            logger.warning(f"WARNING: {col} has dimension {arr.shape[1]} but expected {expected_dim}")
        np.save(os.path.join(split_dir, f"{col}_embeddings.npy"), arr)
        logger.info(f"Saved {col} embeddings with shape {arr.shape}")


def save_checkpoint(results: Sequence[dict], output_dir: str, split: str, checkpoint_num: int):
    d = os.path.join(output_dir, "checkpoints")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, f"checkpoint_{split}_{checkpoint_num}.pkl"), "wb") as f:
        pickle.dump(list(results), f)


def find_latest_checkpoint(output_dir: str, split: str):
    d = os.path.join(output_dir, "checkpoints")
    if not os.path.exists(d):
        return None
    nums = [int(f.split("_")[-1].split(".")[0]) for f in os.listdir(d)
            if f.startswith(f"checkpoint_{split}_") and f.endswith(".pkl")]
    return max(nums) if nums else None


def load_checkpoint(output_dir: str, split: str, checkpoint_num: int) -> list:
    p = os.path.join(output_dir, "checkpoints", f"checkpoint_{split}_{checkpoint_num}.pkl")
    if not os.path.exists(p):
        return []
    with open(p, "rb") as f:
        return pickle.load(f)


def load_split(output_dir: str, split: str):
    """Reader used by tests: (metadata DataFrame, {layer: ndarray}) as REF/model_training_1.py:112-161 builds them."""
    import pandas as pd

    d = os.path.join(output_dir, split)
    meta = pd.read_csv(os.path.join(d, "embedding_metadata.csv"))
    emb = {os.path.splitext(f)[0].replace("_embeddings", ""): np.load(os.path.join(d, f))
           for f in sorted(os.listdir(d)) if f.endswith("_embeddings.npy")}
    return meta, emb


def extract_split(rows: Iterable[dict], pooled_fn: Callable[[list], np.ndarray], layer_indices: Sequence[int],
                  output_dir: str, split: str, load_audio: Callable[[str], np.ndarray | None], prefix: str = "layer_",
                  batch_size: int = 256, checkpoint_interval: int = 50, resume: bool = False,
                  expected_dim: int | None = None) -> list:
    """One split of the reference's main loop (REF/WavLM_embeddings.py:535-650), batched.

    rows: metadata records, each with at least 'path' (REF create_metadata_from_files); pooled_fn: clips -> [B, L+1, D]
    (e.g. `engine.pooled`); prefix: 'layer_' (WavLM) or 'encoder_layer_' (Whisper).
    """
    rows = list(rows)
    results: list = []
    ckpt = 0
    if resume:
        latest = find_latest_checkpoint(output_dir, split)
        if latest is not None:
            results = load_checkpoint(output_dir, split, latest)
            done = {r["path"] for r in results if "path" in r}
            rows = [r for r in rows if r["path"] not in done]
            ckpt = latest + 1
    for i in range(0, len(rows), batch_size):
        batch = rows[i:i + batch_size]
        clips, kept = [], []
        for row in batch:
            a = load_audio(row["path"])
            if a is None:
                logger.warning(f"Failed to extract embeddings for {row['path']}")
                continue
            clips.append(a)
            kept.append(row)
        pooled = None
        if clips:
            try:
                pooled = pooled_fn(clips)
            except Exception as e:  # noqa: BLE001 - a bad clip must not kill the batch: fall back to one by one
                logger.error(f"Batch failed ({e}); retrying clip by clip")
        for j, row in enumerate(kept):
            p = None
            if pooled is not None:
                p = pooled[j]
            else:
                try:
                    p = pooled_fn([clips[j]])[0]
                except Exception as e:  # noqa: BLE001
                    logger.error(f"Error processing {row['path']}: {e}")
            if p is None:
                logger.warning(f"Failed to extract embeddings for {row['path']}")
                continue
            res = dict(row)
            for idx in layer_indices:
                if idx < p.shape[0]:
                    # a fresh array per selected layer: a view would keep the whole [B, L+1, D] batch alive for as
                    # long as `results` lives (all 25 / 33 layers instead of the 3-4 selected)
                    res[f"{prefix}{idx}"] = np.array(p[idx], dtype=np.float32, copy=True)
                else:
                    logger.warning(f"Layer {idx} is out of range (max: {p.shape[0] - 1})")
            results.append(res)
        if ((i + batch_size) % checkpoint_interval == 0) or ((i + batch_size) >= len(rows)):
            save_checkpoint(results, output_dir, split, ckpt)
            ckpt += 1
    if results:
        save_embeddings(results, output_dir, split, expected_dim)
    else:
        logger.warning(f"No embeddings were extracted for {split} split")
    return results
