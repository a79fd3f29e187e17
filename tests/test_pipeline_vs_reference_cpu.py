"""The on-disk contract checked against the REFERENCE's own writer / checkpoint functions, imported
unmodified from /root/reference. Runs only where the reference is mounted (the build container); skipped elsewhere."""
import filecmp
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "WavLM_embeddings.py")),
                                reason="/root/reference is not mounted on this box")


def _import_ref(tmp_path, name):
    cwd = os.getcwd()
    os.chdir(tmp_path)  # the scripts create ./logs at import time (REF/WavLM_embeddings.py:16-25)
    sys.path.insert(0, REF)
    try:
        return __import__(name)
    finally:
        sys.path.remove(REF)
        os.chdir(cwd)


def _results(prefix, n=5, d=6):
    rng = np.random.default_rng(0)
    return [{"filename": f"f{i}.wav", "path": f"/x/f{i}.wav", "label": int(i % 3), "split": "train",
             f"{prefix}4": rng.standard_normal(d).astype(np.float32),
             f"{prefix}2": rng.standard_normal(d).astype(np.float32)} for i in range(n)]


@pytest.mark.parametrize("script,prefix", [("WavLM_embeddings", "layer_"), ("whisper_embeddings_large", "encoder_layer_")])
def test_writer_and_checkpoints_match_reference(tmp_path, script, prefix):
    import pandas as pd

    from ssr_b200 import pipeline

    ref = _import_ref(tmp_path, script)
    res = _results(prefix)
    a, b = str(tmp_path / "ref"), str(tmp_path / "ours")
    ref.save_embeddings(pd.DataFrame(res), a, "train", None)
    pipeline.save_embeddings(res, b, "train")
    assert sorted(os.listdir(os.path.join(a, "train"))) == sorted(os.listdir(os.path.join(b, "train")))
    for f in os.listdir(os.path.join(a, "train")):
        pa, pb = os.path.join(a, "train", f), os.path.join(b, "train", f)
        if f.endswith(".npy"):
            x, y = np.load(pa), np.load(pb)
            assert x.dtype == y.dtype == np.float32 and np.array_equal(x, y)
        else:
            assert filecmp.cmp(pa, pb, shallow=False), f
    # checkpoints: each side can resume from the other's file
    ref.save_checkpoint(res, a, "train", 3)
    pipeline.save_checkpoint(res, b, "train", 3)
    assert ref.find_latest_checkpoint(b, "train") == pipeline.find_latest_checkpoint(a, "train") == 3
    back = ref.load_checkpoint(b, "train", 3)
    assert [r["path"] for r in back] == [r["path"] for r in res]
    np.testing.assert_array_equal(back[1][f"{prefix}4"], res[1][f"{prefix}4"])
    assert len(pipeline.load_checkpoint(a, "train", 3)) == len(res)
