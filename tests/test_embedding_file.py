"""CPU: the binary embedding format (.cemb) carries exactly what the reference's JSON wire format carries."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

from codae.dataset import ConcatenatedEmbeddingDataset
from codae.tool import convert_json_to_cemb, load_dataset_of_embeddings, read_cemb, write_cemb
from codae.tool.embedding_file import read_cemb_header


def _json_embeddings(tmp_path):
    from oracle.gen_golden import synth_embeddings
    g = np.load(os.path.join(GOLDEN, "emb_small.npz"))
    gen = torch.Generator().manual_seed(int(g["seed"]))
    emb = synth_embeddings(96, ["top", "bottom", "shoe"], 16, gen)
    emb["incomplete"] = {"top": [0.0] * 16}
    path = os.path.join(str(tmp_path), "emb.json")
    with open(path, "w") as f:
        json.dump(emb, f)
    return path, emb, g


def test_json_to_cemb_round_trip_matches_reference_dataset(tmp_path):
    path, emb, g = _json_embeddings(tmp_path)
    cats = ["top", "bottom", "shoe"]
    out = os.path.join(str(tmp_path), "emb.cemb")
    nbytes = convert_json_to_cemb(path, out, cats)
    assert nbytes == os.path.getsize(out) and nbytes < os.path.getsize(path) / 3     # binary is several x smaller
    h = read_cemb_header(out)
    assert (h["N"], h["S"], h["E"]) == (96, 3, 16) and h["categories"] == cats and "incomplete" not in h["ids"]
    config = {"DATASET": {"USED_CATEGORY": cats}}
    ds_bin = load_dataset_of_embeddings(out, config)
    ds_json = ConcatenatedEmbeddingDataset(embeddings=emb, used_category=cats)
    assert torch.equal(ds_bin.data, ds_json.data) and ds_bin.scale == ds_json.scale
    assert np.array_equal(ds_bin.data.numpy(), g["data"])                             # == the reference's dataset
    for c in range(3):
        assert torch.equal(ds_bin.data_per_category[c], ds_json.data_per_category[c])
    assert ds_bin.index == ds_json.index and [a["position"] for a in ds_bin.arch] == [0, 16, 32]
    # category subset and order are chosen at read time
    planes, _ = read_cemb(out, ["shoe", "top"])
    assert torch.equal(planes[0], ds_json.data_per_category[2]) and torch.equal(planes[1], ds_json.data_per_category[0])


def test_bf16_catalog_and_errors(tmp_path):
    out = os.path.join(str(tmp_path), "cat.cemb")
    x = torch.rand(50, 32)
    write_cemb(out, [x], ["shoe"], dtype="bf16")
    planes, h = read_cemb(out)
    assert h["dtype"] == 1 and planes[0].dtype == torch.bfloat16
    assert torch.equal(planes[0], x.to(torch.bfloat16))
    with pytest.raises(Exception, match="category top is not in the file"):
        read_cemb(out, ["top"])
    with open(out, "r+b") as f:
        f.truncate(os.path.getsize(out) - 10)
    with pytest.raises(Exception, match="truncated"):
        read_cemb(out)
    bad = os.path.join(str(tmp_path), "bad.cemb")
    open(bad, "wb").write(b"NOTCEMB!" + b"\0" * 100)
    with pytest.raises(Exception, match="bad magic"):
        read_cemb(bad)
