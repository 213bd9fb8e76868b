"""Binary embedding file (.cemb) -- replacement for the reference's JSON wire format on the way INTO the hot path.

The reference stores `{obs_id: {category: [E floats]}}` as JSON text (script/encode_coco.py:65-78, encode_polyvore.py:67-86)
and rebuilds the dataset with O(N*S) torch.cat calls (codae/dataset/concatenated_embedding_dataset.py:56-67), cached as a
pickle keyed on st_ctime (codae/tool/data_tool.py:114-162).  A .cemb file holds the same information as one header plus
a raw row-major [N, S, E] array that is memory-mapped and copied to the device in one piece; it is also the catalog
format of stage IV (one [N, E] plane per category).

layout (little endian):
    bytes 0..7    magic  b"CODAEMB1"
    bytes 8..39   uint64 N, uint64 S, uint64 E, uint64 dtype (0 = float32, 1 = bfloat16)
    bytes 40..47  uint64 M = length of the JSON metadata block
    M bytes       JSON {"categories": [...], "ids": [...] | null}
    pad to 64     zero bytes up to the next multiple of 64
    data          N * S * E elements, row-major [N, S, E]  (UN-scaled values, like data_per_category in the reference)
"""
import json
import os
import struct

import numpy as np
import torch

MAGIC = b"CODAEMB1"
_DTYPES = {0: (np.float32, torch.float32), 1: (np.uint16, torch.bfloat16)}


def write_cemb(path, per_category, categories, ids=None, dtype="fp32"):
    """per_category: list of S arrays/tensors [N, E] (un-scaled).  Returns the number of bytes written."""
    planes = [torch.as_tensor(c, dtype=torch.float32) for c in per_category]
    N, E = planes[0].shape
    S = len(planes)
    data = torch.stack(planes, dim=1).contiguous()          # [N, S, E]
    code = 0 if dtype == "fp32" else 1
    if code == 1:
        raw = data.to(torch.bfloat16).view(torch.int16).numpy().tobytes()
    else:
        raw = data.numpy().astype("<f4", copy=False).tobytes()
    meta = json.dumps({"categories": list(categories), "ids": None if ids is None else [str(i) for i in ids]}).encode("utf-8")
    head = MAGIC + struct.pack("<5Q", N, S, E, code, len(meta)) + meta
    head += b"\0" * ((-len(head)) % 64)
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(head)
        f.write(raw)
    os.replace(tmp, path)
    return len(head) + len(raw)


def read_cemb_header(path):
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise Exception("Error while reading embedding file: bad magic.")
        N, S, E, code, M = struct.unpack("<5Q", f.read(40))
        if code not in _DTYPES:
            raise Exception("Error while reading embedding file: unknown dtype code %d." % code)
        meta = json.loads(f.read(M).decode("utf-8"))
    offset = 48 + M
    offset += (-offset) % 64
    return dict(N=N, S=S, E=E, dtype=code, categories=meta["categories"], ids=meta["ids"], offset=offset)


def read_cemb(path, used_category=None, device=None):
    """Memory-maps the file and returns (per_category list of [N, E] tensors, header).  With `device`, the selected
    planes are staged through pinned memory and copied asynchronously."""
    h = read_cemb_header(path)
    np_dt, t_dt = _DTYPES[h["dtype"]]
    expect = h["offset"] + h["N"] * h["S"] * h["E"] * np.dtype(np_dt).itemsize
    if os.path.getsize(path) != expect:
        raise Exception("Error while reading embedding file: truncated (%d != %d bytes)." % (os.path.getsize(path), expect))
    mm = np.memmap(path, dtype=np_dt, mode="r", offset=h["offset"], shape=(h["N"], h["S"], h["E"]))
    cats = h["categories"] if used_category is None else used_category
    out = []
    for c in cats:
        if c not in h["categories"]:
            raise Exception("Error while reading embedding file: category %s is not in the file." % c)
        plane = np.ascontiguousarray(mm[:, h["categories"].index(c), :])
        t = torch.from_numpy(plane)
        if h["dtype"] == 1:
            t = t.view(torch.bfloat16)
        if device is not None and torch.device(device).type == "cuda":
            t = t.pin_memory().to(device, non_blocking=True)
        out.append(t)
    return out, h


def convert_json_to_cemb(json_path, out_path, used_category, dtype="fp32"):
    """JSON wire format -> .cemb, keeping only observations that have every used category (the reference's filter,
    concatenated_embedding_dataset.py:28-38) in file order."""
    with open(json_path, "r") as f:
        emb = json.load(f)
    ids = [k for k, v in emb.items() if all(c in v for c in used_category)]
    planes = [np.asarray([emb[i][c] for i in ids], dtype=np.float32) for c in used_category]
    return write_cemb(out_path, planes, used_category, ids=ids, dtype=dtype)


def load_cemb_dataset(path, used_category):
    """ConcatenatedEmbeddingDataset from a .cemb file (same `data`, `scale`, `arch`, `data_per_category` as from JSON)."""
    from codae.dataset import ConcatenatedEmbeddingDataset
    planes, h = read_cemb(path, used_category)
    ds = ConcatenatedEmbeddingDataset.from_tensors([p.float() for p in planes], list(used_category))
    if h["ids"] is not None:
        ds.index = list(h["ids"])
    return ds
