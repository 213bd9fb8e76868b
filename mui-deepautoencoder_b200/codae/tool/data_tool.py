"""Corrupter, Normalizer, collate and loader -- the reference's codae/tool/data_tool.py API.

`Corrupter` keeps the reference's attributes (`nb_run`, `binary_masks`, `nb_missing_per_run`,
`mask_to_use`, `nb_corruption_per_k`, `nb_subset_per_variable`) but the per-batch work is one CUDA kernel:
masks are computed from (mask id, arch tables) on the device instead of a Python loop over the batch.
"""
import glob
import hashlib
import itertools
import json
import os
import pickle
import random

import torch

from codae import _C
from codae.dataset import ConcatenatedEmbeddingDataset


def get_mask_transformation(observation_mask, loss_mask):
    """[len(observation_mask), len(loss_mask)] 0/1 matrix from column space to variable space: every 1 of
    `observation_mask` opens its own variable, a run of 0s shares one (data_tool.py:16-43)."""
    T = torch.zeros((len(observation_mask), len(loss_mask)))
    opened, c = True, 0
    for i in range(len(observation_mask)):
        if observation_mask[i] == 1:
            opened = True
            T[i, c] = 1
            c += 1
        elif opened:
            opened = False
            T[i, c] = 1
            c += 1
    return T


class Normalizer:
    """Min-max (de)normalisation with a fitted scaler's attributes (data_tool.py:46-90)."""

    def __init__(self, normalizer, device, normalization_type="min_max"):
        self.normalization_type = normalization_type
        self.device = device
        self.min = torch.Tensor(normalizer.data_min_).to(device)
        self.max = torch.Tensor(normalizer.data_max_).to(device)
        self.scale = torch.Tensor(normalizer.data_range_).to(device)

    def do(self, data):
        return (data - self.min) / self.scale

    def undo(self, data):
        return (data * self.scale) + self.min


def collate_embedding(batch):
    """(stacked rows, tuple of indices) (data_tool.py:96-103)."""
    batch, indices = zip(*batch)
    return torch.stack(batch), indices


def simple_collate(batch):
    return torch.stack(batch)


def load_dataset_of_embeddings(embedding_path, config, cache_dir="tmp/"):
    """JSON {obs_id: {category: [floats]}} -> ConcatenatedEmbeddingDataset, with the reference's pickle cache
    keyed on sha1(st_ctime) (data_tool.py:114-162).  A path ending in .cemb is read as the binary format of
    codae.tool.embedding_file (memory-mapped, no cache needed)."""
    if str(embedding_path).endswith(".cemb"):
        from codae.tool.embedding_file import load_cemb_dataset
        return load_cemb_dataset(embedding_path, config["DATASET"]["USED_CATEGORY"])
    using_cache = False
    dataset = None
    dataset_cache = glob.glob(os.path.join(cache_dir, "*_dataset.bin"))
    key = hashlib.sha1(str(os.stat(embedding_path)[9]).encode('utf-8')).hexdigest()
    if len(dataset_cache) > 0:
        path = dataset_cache[0]
        if path.split("/")[-1].split("_")[0] == key:
            using_cache = True
            try:
                with open(path, 'rb') as f:
                    dataset = pickle.load(f)
            except Exception:
                raise Exception("Error while reading embedding json file.")
        else:
            os.remove(path)
    if not using_cache:
        try:
            with open(embedding_path, 'r') as f:
                embeddings = json.load(f)
        except Exception:
            raise Exception("Error while reading embedding json file.")
        dataset = ConcatenatedEmbeddingDataset(embeddings=embeddings, used_category=config["DATASET"]["USED_CATEGORY"])
        os.makedirs(cache_dir, exist_ok=True)
        tmp = os.path.join(cache_dir, "new_dataset_tmp.bin")
        with open(tmp, "wb") as f:
            pickle.dump(dataset, f)
        os.rename(tmp, os.path.join(cache_dir, key + "_dataset.bin"))
    return dataset


class Corrupter:
    """Create, handle, and keep track of corruption masks (data_tool.py:165-262).

    seed=None  : `mask_to_use` is drawn with sequential random.sample calls exactly like the reference
                 (identical table under random.seed).
    seed=int   : `mask_to_use` comes from the Philox4x32-10 kernel (codae_mask_table_philox): ids depend only
                 on (seed, observation), so data-parallel ranks agree without communication.
    `mask_to_use` stays a plain attribute: assigning a new table re-uploads it on next use.
    """

    MAX_VARIABLES = 64

    def __init__(self, nb_observation, arch, k_max, device, seed=None):
        self.nb_observation = nb_observation
        self.arch = arch
        self.k_max = k_max
        self.device = device
        if (k_max < 0) | (k_max > len(self.arch) - 1):
            raise Exception("Invalid k_max number. k_max > 0 && k_max < nb_predictor - 1")
        if len(arch) > self.MAX_VARIABLES:
            raise Exception("Error: at most %d variables are supported by the mask kernels" % self.MAX_VARIABLES)
        self.io_size = sum([v["size"] for v in self.arch])
        self.nb_predictor = len(self.arch)

        subsets = []
        self.nb_corruption_per_k = [0 for _ in range(k_max)]
        for k in range(k_max):
            s = list(itertools.combinations(range(self.nb_predictor), k + 1))
            self.nb_corruption_per_k[k] = len(s)
            subsets.extend(s)
        self.nb_run = sum(self.nb_corruption_per_k)
        self.subsets = subsets
        # the device mask-id table is int16 and the kernels keep one 64-bit word of variables per mask
        if self.nb_run > 32767:
            raise Exception("Too many corruption subsets (%d): the device mask table holds ids up to 32767." % self.nb_run)
        if self.nb_predictor > 64:
            raise Exception("Too many variables (%d): the corruption kernels support at most 64." % self.nb_predictor)

        self.binary_masks = torch.ones((max(self.nb_run, 0), self.io_size))
        for r, sub in enumerate(subsets):
            for v in sub:
                self.binary_masks[r, arch[v]["position"]:arch[v]["position"] + arch[v]["size"]] = 0
        self.nb_missing_per_run = [len(s) for s in subsets]

        self.corrupted_index = [x for x in range(self.nb_run)]
        self.seed = seed
        self._table = None
        self._table_src = None
        if seed is None:
            self.mask_to_use = torch.stack([torch.LongTensor(random.sample(self.corrupted_index, self.nb_run))
                                            for _ in range(nb_observation)]) if nb_observation > 0 else \
                torch.zeros((0, self.nb_run), dtype=torch.long)
        else:
            self._table = _C.mask_table_philox(int(seed), 0, nb_observation, self.nb_run, self._cuda_device())
            self.mask_to_use = self._table.cpu().long()
            self._table_src = self.mask_to_use

        self.nb_subset_per_variable = []
        for i in range(1, self.k_max + 1):
            k_subset = 1
            for j in range(1, i):
                k_subset *= (self.nb_predictor - j) / 2
            self.nb_subset_per_variable.append(k_subset)
        self._tables = None

    # ---- device tables ------------------------------------------------------------------------------
    def _cuda_device(self):
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError("codae: Corrupter needs a CUDA device; the B200 path has no CPU fallback")
        return dev

    def device_tables(self):
        """(mask_table int16 [N, nb_run], mask_bits int64 [nb_run], col_var uint8 [io], nb_missing uint8 [nb_run])."""
        dev = self._cuda_device()
        if self._tables is None:
            bits = []
            for sub in self.subsets:
                b = 0
                for v in sub:
                    b |= 1 << v
                bits.append(b if b < (1 << 63) else b - (1 << 64))
            col_var = torch.zeros(self.io_size, dtype=torch.uint8)
            for v, var in enumerate(self.arch):
                col_var[var["position"]:var["position"] + var["size"]] = v
            self._tables = (torch.tensor(bits, dtype=torch.int64, device=dev), col_var.to(dev),
                            torch.tensor(self.nb_missing_per_run, dtype=torch.uint8, device=dev))
        if self._table is None or self._table_src is not self.mask_to_use:
            self._table = torch.as_tensor(self.mask_to_use).to(torch.int16).contiguous().to(dev)
            self._table_src = self.mask_to_use
        return (self._table,) + self._tables

    def batch_index_tensor(self, batch_indices):
        if torch.is_tensor(batch_indices):
            return batch_indices.to(device=self._cuda_device(), dtype=torch.int64)
        return torch.tensor(list(batch_indices), dtype=torch.int64, device=self._cuda_device())

    def get_masks(self, batch_indices, run):
        """(list of k_max dense masks [B, io], their sum) -- one kernel instead of the reference's Python loop
        over the batch (data_tool.py:239-262)."""
        table, bits, col_var, nmiss = self.device_tables()
        idx = self.batch_index_tensor(batch_indices)
        B = idx.numel()
        masks = torch.empty((self.k_max, B, self.io_size), dtype=torch.float32, device=table.device)
        fmask = torch.empty((B, self.io_size), dtype=torch.float32, device=table.device)
        _C.dense_masks(idx, B, table, run, bits, nmiss, col_var, self.io_size, self.k_max, masks, fmask)
        return [masks[k] for k in range(self.k_max)], fmask
