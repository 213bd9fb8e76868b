"""Shared helpers of the entry-point scripts: import path, synthetic data of the configs' shapes,
data-parallel bootstrap (one process per GPU, torch.distributed / NCCL)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def init_distributed():
    """(rank, world_size, device).  Under torchrun: NCCL process group, one GPU per rank."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("codae: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return rank, world, torch.device("cuda", local)


def synthetic_categories(n, nb_category, embedding_size, seed, device="cpu"):
    """Un-scaled per-category embeddings shaped like post-ReLU ResNet features: |N(0,1)| * Bernoulli(0.7)."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = []
    for _ in range(nb_category):
        v = torch.randn((n, embedding_size), generator=g, device=device).abs_()
        v *= (torch.rand((n, embedding_size), generator=g, device=device) < 0.7)
        out.append(v)
    return out


def epoch_batches(indices, batch_size, rng, rank=0, world=1):
    """SubsetRandomSampler semantics (a fresh permutation of `indices` per epoch, ragged last batch), with the
    global batch split by rank for data parallelism: yields (local_indices, global_batch_size)."""
    perm = rng.permutation(len(indices))
    idx = np.asarray(indices)[perm]
    for s in range(0, len(idx), batch_size):
        g = idx[s:s + batch_size]
        yield g[rank::world], len(g)
