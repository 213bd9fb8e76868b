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
    if len(arch) > _C.MIXED_MAX_VARIABLES or any(int(v["size"]) > _C.MIXED_MAX_CATEGORIES for v in arch):
        # the tabular loss / monitor kernels keep one variable's logits in registers (csrc/loss.cu: kMaxVar)
        raise Exception("The tabular loss kernels support at most %d variables of at most %d categories each."
                        % (_C.MIXED_MAX_VARIABLES, _C.MIXED_MAX_CATEGORIES))
    pos = torch.tensor([v["position"] for v in arch], dtype=torch.int32, device=device)
    size = torch.tensor([v["size"] for v in arch], dtype=torch.int32, device=device)
    typ = torch.tensor([_C.VAR_REGRESSION if v["type"] == "regression" else _C.VAR_CLASSIFICATION for v in arch],
                       dtype=torch.int32, device=device)
    return pos, size, typ


class RankingLoss:
    """Rank of the true item among the validation items by cosine similarity with the reconstructed slot
    (metering.py:29-79).  The per-sample cosine_similarity + O(|val|) Python loop becomes one
    codae_score_rank sweep per masked category."""

    GEMM_MIN_Q = 64      # queries of one category from which the dot products go through the tensor-core contraction

    def __init__(self, dataset, validation_indices, device):
        self.dataset = dataset
        self.device = device
        self.validation_indices = validation_indices
        self._val = None
        self._gemm = {}      # category -> cached operand of the GEMM path (planes of the validation rows, their |c|^2)

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
            if rows.numel() >= self.GEMM_MIN_Q and self._gemm_ok(catalog, E):
                r = self._ranks_gemm(c, catalog, E, q, idx[rows].contiguous())
            else:
                _C.score_rank(catalog, E, q, 1.0, _C.METRIC_COSINE, idx[rows].contiguous(), self._val, r)
            out[rows] = r
        return out

    @staticmethod
    def _gemm_ok(catalog, E):
        return (catalog.dtype == torch.float32 and E % 8 == 0 and
                _C.linear_engine(catalog.device, _C.F32X3, 128, 128, E) == _C.ENGINE_TCGEN05_F32X3)

    def _ranks_gemm(self, c, catalog, E, q, true_idx, max_q=1024):
        """Q >= 64 queries of category c: ONE [Q, E] x [E, n_val + Q] contraction on the fp32-parity tensor-core engine (operands
        as bf16 triples, fp32 accumulation) + codae_rank_count.  The validation rows of the category are gathered and split into
        planes once; the Q true rows are appended per call so that a true item inside the subset scores bit-identically on both
        sides of the strict comparison of metering.py:73."""
        dev = q.device
        n = int(self._val.numel())
        ent = self._gemm.get(c)
        if ent is None:
            rows = catalog[self._val].contiguous()                                  # [n, E] f32
            planes = _C.new_x3((n + max_q, E), dev)
            _C.split_x3(rows, planes[:, :n])
            cc = torch.zeros(n + max_q, dtype=torch.float32, device=dev)
            _C.row_sqnorm(rows, E, cc)
            ent = (planes, cc, torch.zeros(max_q, dtype=torch.float32, device=dev))
            self._gemm[c] = ent
        planes, cc, qq = ent
        out = torch.zeros(q.shape[0], dtype=torch.int64, device=dev)
        for lo in range(0, q.shape[0], max_q):
            qs = q[lo:lo + max_q]
            Q = int(qs.shape[0])
            true_rows = catalog[true_idx[lo:lo + Q]].contiguous()
            _C.split_x3(true_rows, planes[:, n:n + Q])
            _C.row_sqnorm(true_rows, E, cc[n:])
            _C.row_sqnorm(qs, E, qq)
            qp = _C.new_x3((Q, E), dev)
            _C.split_x3(qs, qp)
            ld = (n + Q + 3) // 4 * 4
            scores = torch.empty((Q, ld), dtype=torch.float32, device=dev)
            _C.linear_fwd(qp, planes[:, :n + Q], None, scores[:, :n + Q], Q, n + Q, E, _C.ACT_NONE, _C.F32X3)
            r = torch.zeros(Q, dtype=torch.int64, device=dev)
            _C.rank_count(scores, Q, n, cc, qq, r)
            out[lo:lo + Q] = r
        return out

    def get_tensor(self, prediction, fmask, indices):
        """sum_i 1 - rank_i / (|val| - 1) as a 0-dim float64 DEVICE tensor: a validation loop can accumulate it without a host
        synchronisation per batch (the reference adds a Python float per sample, metering.py:76)."""
        ranks = self.ranks(prediction, fmask, indices)
        n = len(self.validation_indices)
        return (1 - ranks.double() / (n - 1)).sum()

    def get(self, prediction, fmask, indices):
        return float(self.get_tensor(prediction, fmask, indices).item())


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
