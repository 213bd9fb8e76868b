"""get_rmse, RankingLoss, CombinedCriterion -- the reference's codae/tool/metering.py API on the B200 kernels."""
import numpy as np
import torch

from codae import _C
from codae.tool.data_tool import get_mask_transformation


def get_rmse(x, y):
    """RMSE between two scalars / vectors / arrays (metering.py:24-26)."""
    return np.sqrt(np.mean((x - y) ** 2))


def arch_tables(arch, device):
    """(var_pos, var_size, var_type) int32 device tensors of an `arch` list."""
    pos = torch.tensor([v["position"] for v in arch], dtype=torch.int32, device=device)
    size = torch.tensor([v["size"] for v in arch], dtype=torch.int32, device=device)
    typ = torch.tensor([_C.VAR_REGRESSION if v["type"] == "regression" else _C.VAR_CLASSIFICATION for v in arch],
                       dtype=torch.int32, device=device)
    return pos, size, typ


class RankingLoss:
    """Rank of the true item among the validation items by cosine similarity with the reconstructed slot
    (metering.py:29-79).  The per-sample cosine_similarity + O(|val|) Python loop becomes one
    codae_score_rank sweep per masked category."""

    def __init__(self, dataset, validation_indices, device):
        self.dataset = dataset
        self.device = device
        self.validation_indices = validation_indices
        self._val = None

    def ranks(self, prediction, fmask, indices):
        E = self.dataset.embedding_size
        S = self.dataset.nb_used_category
        dev = prediction.device
        if self._val is None or self._val.device != dev:
            self._val = torch.as_tensor(list(self.validation_indices), dtype=torch.int64, device=dev)
        idx = torch.as_tensor(list(indices), dtype=torch.int64, device=dev) if not torch.is_tensor(indices) \
            else indices.to(dev, torch.int64)
        # masked category of each row: the slot whose first column is zeroed (valid for k_max = 1, metering.py:56)
        cat = torch.argmax(1 - fmask[:, ::E][:, :S], dim=1)
        out = torch.zeros(idx.numel(), dtype=torch.int64, device=dev)
        for c in torch.unique(cat).tolist():
            rows = torch.nonzero(cat == c).flatten()
            q = prediction[rows, c * E:(c + 1) * E].detach().to(torch.float32).contiguous()
            r = torch.zeros(rows.numel(), dtype=torch.int64, device=dev)
            catalog = self.dataset.data_per_category[c]
            if not catalog.is_cuda:
                raise RuntimeError("codae: dataset is not on a CUDA device (dataset.to(device)); no CPU fallback")
            _C.score_rank(catalog, E, q, 1.0, _C.METRIC_COSINE, idx[rows].contiguous(), self._val, r)
            out[rows] = r
        return out

    def get(self, prediction, fmask, indices):
        ranks = self.ranks(prediction, fmask, indices)
        n = len(self.validation_indices)
        return float((1 - ranks.double() / (n - 1)).sum().item())


class _MixedMeanLoss(torch.autograd.Function):

    @staticmethod
    def forward(ctx, y, x, crit):
        xs, ys = x.detach().contiguous(), y.detach().contiguous()
        dy = torch.empty_like(ys)
        out = torch.empty(1 + len(crit.arch), dtype=torch.float32, device=ys.device)
        pos, size, typ, w = crit._tables(ys.device)
        _C.mixed_loss_fwd_bwd(xs, ys, pos, size, typ, w, dy, out)
        ctx.save_for_backward(dy)
        crit.last_per_variable = out[1:]
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (dy,) = ctx.saved_tensors
        return dy * g, None, None


class CombinedCriterion:
    """Mix of per-variable RMSE (regression) and NLL (classification) losses (metering.py:82-204)."""

    def __init__(self, arch, k_max, device, observation_mask, weight=None, reduction="none"):
        # the reference's guard `k_max < 0 | k_max >= len(arch)` is a precedence bug that never fires
        # (metering.py:90); observable behaviour (no raise) is kept.
        self.arch = arch
        self.k_max = k_max
        self.device = device
        self.reduction = reduction
        self.weight = torch.ones(len(self.arch)) if weight is None else torch.Tensor(weight)
        self.observation_mask = observation_mask
        self.io_size = len(self.observation_mask)
        self.loss_mask = [1 if v["type"] == "continuous" else 0 for v in arch]
        self.mask_transformation = get_mask_transformation(observation_mask=self.observation_mask,
                                                           loss_mask=self.loss_mask).cpu().numpy()
        self._dev_tables = None
        self.last_per_variable = None

    def _tables(self, device):
        if self._dev_tables is None or self._dev_tables[0].device != device:
            self._dev_tables = arch_tables(self.arch, device) + (self.weight.to(device=device, dtype=torch.float32),)
        return self._dev_tables

    def __call__(self, x, y, as_numpy=False):
        if self.reduction == "mean":
            return self._mean_loss(x, y, as_numpy)
        elif self.reduction == "none":
            return self._full_loss(x, y, as_numpy)
        else:
            raise Exception("Unknown reduction type.")

    def _full_loss(self, x, y, as_numpy=False):
        """[B, V] per-row squared error / NLL, returned on the host like the reference (metering.py:131-152)."""
        if not y.is_cuda:
            raise RuntimeError("codae: tensors are not on a CUDA device; the B200 path has no CPU fallback")
        xs, ys = x.detach().contiguous(), y.detach().contiguous()
        B, V = xs.shape[0], len(self.arch)
        pos, size, typ, _ = self._tables(ys.device)
        out = torch.empty((B, V), dtype=torch.float32, device=ys.device)
        acc = torch.zeros(2 + 2 * V, dtype=torch.float64, device=ys.device)
        zero_id = torch.zeros(B, dtype=torch.int32, device=ys.device)
        bits = torch.zeros(1, dtype=torch.int64, device=ys.device)
        nmiss = torch.ones(1, dtype=torch.uint8, device=ys.device)
        _C.mixed_monitor(xs, ys, pos, size, typ, None, None, xs.shape[1], zero_id, bits, nmiss, 1, out, acc)
        loss = out.cpu()
        return loss.numpy() if as_numpy else loss

    def _mean_loss(self, x, y, as_numpy=False):
        """sum_i w_i l_i / V, differentiable wrt y (metering.py:155-180)."""
        if not y.is_cuda:
            raise RuntimeError("codae: tensors are not on a CUDA device; the B200 path has no CPU fallback")
        loss = _MixedMeanLoss.apply(y, x, self)
        if as_numpy:  # un-weighted mean of the per-variable losses (metering.py:173-175)
            return (self.last_per_variable.sum() / len(self.arch)).cpu().numpy()
        return loss

    def get_per_k(self, loss, masks):
        """[k_max, V]: column sums of `loss` over the rows whose mask removes exactly k variables
        (metering.py:187-197, written without the io x io ones-matrix product)."""
        out = np.zeros((self.k_max, len(self.arch)))
        for i, mask in enumerate(masks):
            m = mask.detach().cpu().numpy()
            rows = (m.sum(axis=1) > 0).astype(loss.dtype)[:, None]
            out[i, :] = np.sum((rows * np.matmul(np.ones_like(m), self.mask_transformation)) * loss, axis=0)
        return out

    def get_partial(self, loss, mask):
        """(1 - mask . T) * loss (metering.py:200-204)."""
        return (1 - np.matmul(mask.detach().cpu().numpy(), self.mask_transformation)) * loss
