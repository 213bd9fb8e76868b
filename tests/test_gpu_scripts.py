"""GPU: the entry-point scripts run end to end with the reference's flags and YAML files (synthetic data of the configs'
shapes), single process.  They are the drop-in for script/train_dae_on_*.py and the stage-IV entry point."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG

pytestmark = pytest.mark.gpu
SCRIPT = os.path.join(PKG, "script")
CONFIG = os.path.join(PKG, "config")


def run(args, cwd):
    r = subprocess.run([sys.executable] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_train_dae_on_embedding_script(tmp_path):
    out = run([os.path.join(SCRIPT, "train_dae_on_embedding.py"), "--output_path", str(tmp_path), "--config",
               os.path.join(CONFIG, "embedding.yaml"), "--synthetic", "1024", "--epochs", "2", "--rank", "1", "--graph"], str(tmp_path))
    assert "TRAINING HAS ENDED." in out and "VALIDATION RANKING ERROR" in out
    d = [p for p in os.listdir(str(tmp_path)) if p.endswith("_train_EMBEDDING")]
    assert len(d) == 1
    m = np.load(os.path.join(str(tmp_path), d[0], "metrics.npz"))
    assert m["ftl"].shape == (2,) and np.all(np.isfinite(m["ftl"])) and np.all(np.isfinite(m["pvl"]))
    assert m["ftl"][1] < m["ftl"][0]                    # the full training error goes down
    assert 0.0 <= float(m["rl"][-1]) <= 1.0             # normalised rank of the true item
    assert os.path.exists(os.path.join(str(tmp_path), d[0], "model.pt"))


def test_modanet_config_without_trunk_grad_key(tmp_path):
    out = run([os.path.join(SCRIPT, "train_dae_on_embedding.py"), "--output_path", str(tmp_path), "--config",
               os.path.join(CONFIG, "modanet_merge_top_bottom_shoe.yaml"), "--synthetic", "256", "--epochs", "1"], str(tmp_path))
    assert "TRAINING HAS ENDED." in out                  # the reference raises KeyError('TRUNK_GRAD') here


def test_train_dae_on_abalone_script(tmp_path):
    out = run([os.path.join(SCRIPT, "train_dae_on_abalone.py"), "--output_path", str(tmp_path), "--config",
               os.path.join(CONFIG, "abalone.yaml"), "--synthetic", "512", "--epochs", "2", "--nb_missing", "2"], str(tmp_path))
    assert "TRAINING HAS ENDED." in out and "VALIDATION PARTIAL ERROR" in out and "k=2" in out


def test_complementarity_inference_script(tmp_path):
    out = run([os.path.join(SCRIPT, "4_complementarity_inference.py"), "--config", os.path.join(CONFIG, "embedding.yaml"),
               "--synthetic", "5000", "--slot", "1", "--k", "7", "--queries", "3", "--metric", "cosine"], str(tmp_path))
    res = json.loads(out.strip().splitlines()[-1])
    assert res["slot"] == 1 and len(res["indices"]) == 3 and all(len(r) == 7 for r in res["indices"])
    assert all(0 <= i < 5000 for r in res["indices"] for i in r)
    assert all(r[j] >= r[j + 1] for r in res["scores"] for j in range(6))       # cosine: best first


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_complementarity_inference_script_swap_mode(tmp_path, dtype):
    out = run([os.path.join(SCRIPT, "4_complementarity_inference.py"), "--config", os.path.join(CONFIG, "embedding.yaml"),
               "--synthetic", "3000", "--slot", "2", "--k", "5", "--queries", "2", "--mode", "swap", "--compute_dtype", dtype],
              str(tmp_path))
    res = json.loads(out.strip().splitlines()[-1])
    assert res["mode"] == "swap" and len(res["indices"]) == 2 and all(len(r) == 5 for r in res["indices"])
    assert all(0 <= i < 3000 for r in res["indices"] for i in r)
    assert all(r[j] <= r[j + 1] for r in res["scores"] for j in range(4))       # reconstruction error: best first
