"""MixedVariableDataset -- tabular boundary object: `data` [N, io] fp32 (one-hot blocks for categorical
columns, first-appearance label order) and `arch` (codae/dataset/mixed_variable_dataset.py of the reference)."""
import numpy as np
import torch
from torch.utils.data.dataset import Dataset


class MixedVariableDataset(Dataset):

    def __init__(self, pd_dataset):
        """pd_dataset: pandas DataFrame; float64/int64 columns are regression variables of size 1, any other
        dtype a classification variable of size nunique() (…dataset.py:28-49)."""
        self.pd_dataset = pd_dataset
        self.variable_names = list(pd_dataset.columns)
        self.nb_predictor = len(pd_dataset.columns)
        self.nb_observation = len(pd_dataset)
        self.io_size = 0
        self.arch = []
        for i, column in enumerate(pd_dataset):
            dtype = str(pd_dataset.dtypes.iloc[i])
            var = {"name": column, "lambda": 1}
            if dtype in ("float64", "int64"):          # exactly the reference's rule (mixed_variable_dataset.py:34)
                var["size"], var["type"] = 1, "regression"
            else:
                var["size"], var["type"] = int(pd_dataset[column].nunique()), "classification"
            var["position"] = self.io_size
            self.io_size += var["size"]
            self.arch.append(var)
        self.type_mask = torch.zeros((self.io_size))
        for var in self.arch:
            if var["type"] == "regression":
                self.type_mask[var["position"]:var["position"] + var["size"]] = 1
        self.map = {}
        data = np.zeros((self.nb_observation, self.io_size))
        for var in self.arch:
            col = pd_dataset[var["name"]].to_numpy()
            if var["type"] == "classification":
                labels = {}
                for v in col:           # first-appearance order (…dataset.py:65-81)
                    if v not in labels:
                        labels[v] = len(labels)
                self.map[var["name"]] = dict(labels, COUNT=len(labels))
                codes = np.fromiter((labels[v] for v in col), dtype=np.int64, count=len(col))
                data[np.arange(self.nb_observation), var["position"] + codes] = 1
            else:
                data[:, var["position"]] = col.astype(np.float64)
        self.data = torch.Tensor(data)

    @classmethod
    def from_arch(cls, arch, data, variable_names=None):
        """Build from a ready `arch` list and an [N, io] tensor (synthetic data)."""
        self = cls.__new__(cls)
        self.pd_dataset = None
        self.arch = arch
        self.variable_names = variable_names or [v["name"] for v in arch]
        self.nb_predictor = len(arch)
        self.data = torch.as_tensor(data, dtype=torch.float32)
        self.nb_observation = int(self.data.shape[0])
        self.io_size = sum(v["size"] for v in arch)
        self.type_mask = torch.zeros((self.io_size))
        for var in arch:
            if var["type"] == "regression":
                self.type_mask[var["position"]:var["position"] + var["size"]] = 1
        self.map = {}
        return self

    def __len__(self):
        return self.nb_observation

    def __getitem__(self, idx):
        return self.data[idx], idx

    def _categorical_to_OHE(self, label, max):
        out = np.zeros(max)
        out[label] = 1
        return out

    def to(self, device):
        self.data = self.data.to(device)
        self.type_mask = self.type_mask.to(device)
