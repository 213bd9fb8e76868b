"""ConcatenatedEmbeddingDataset -- the boundary object that supplies the resident [N, S*E] fp32 tensor,
`arch` and `data_per_category` (codae/dataset/concatenated_embedding_dataset.py of the reference)."""
import numpy as np
import torch
from torch.utils.data.dataset import Dataset


class ConcatenatedEmbeddingDataset(Dataset):

    def __init__(self, embeddings, used_category, transform=None):
        """embeddings: {obs_id: {category: [E floats]}} (the on-disk wire format, script/encode_coco.py:65-78).
        Observations missing a used category are dropped (…dataset.py:28-38); data = concatenation / (max - min)
        -- scaled, not shifted (…:69-74); data_per_category stays un-scaled (…:45-49,76-78)."""
        self.embeddings = embeddings
        self.transform = transform
        self.used_category = used_category
        self.nb_used_category = len(self.used_category)
        self.filtered_embeddings = {}
        self.index = []
        for k, v in self.embeddings.items():
            if all(uc in v for uc in self.used_category):
                self.index.append(k)
                self.filtered_embeddings[k] = v
        self.nb_observation = len(self.index)
        self.embedding_size = len(self.filtered_embeddings[self.index[0]][self.used_category[0]])
        per_cat = [np.asarray([self.filtered_embeddings[i][c] for i in self.index], dtype=np.float32)
                   for c in self.used_category]
        self._finish(per_cat)

    @classmethod
    def from_tensors(cls, per_category, used_category=None, transform=None):
        """Build from un-scaled [N, E] tensors, one per category (synthetic data, binary loaders)."""
        self = cls.__new__(cls)
        self.embeddings = None
        self.filtered_embeddings = None
        self.transform = transform
        self.used_category = used_category or [str(i) for i in range(len(per_category))]
        self.nb_used_category = len(per_category)
        self.nb_observation = int(per_category[0].shape[0])
        self.index = list(range(self.nb_observation))
        self.embedding_size = int(per_category[0].shape[1])
        self._finish([torch.as_tensor(c, dtype=torch.float32) for c in per_category])
        return self

    def _finish(self, per_cat):
        self.data_per_category = {n: torch.as_tensor(c).clone() for n, c in enumerate(per_cat)}
        self.data = torch.cat([self.data_per_category[n] for n in range(self.nb_used_category)], dim=1)
        self.min = self.data.min()
        self.max = self.data.max()
        self.scale = (self.max - self.min).item()
        self.data = self.data / self.scale
        self.arch = []
        self.io_size = 0
        for name in self.used_category:
            self.arch.append({"name": name, "lambda": 1, "size": self.embedding_size, "type": "regression",
                              "position": self.io_size})
            self.io_size += self.embedding_size
        self.type_mask = torch.ones((self.io_size))
        self.nb_predictor = self.embedding_size * self.nb_used_category

    def __len__(self):
        return self.nb_observation

    def __getitem__(self, idx):
        if self.transform is not None:
            return self.transform(self.data[idx]), idx
        return self.data[idx], idx

    def to(self, device):
        """Move the resident tensors to `device` (…dataset.py:133-143)."""
        self.data = self.data.to(device)
        for i in range(len(self.data_per_category.keys())):
            self.data_per_category[i] = self.data_per_category[i].to(device)
