"""Stage IV -- complementarity inference (the entry point the reference README names, README.md:14-16,30-32,
whose script is absent from the tree; semantics follow RankingLoss, codae/tool/metering.py:46-79).

context outfit with slot c zeroed -> DAE forward -> predicted embedding p of slot c -> score every catalog
item of category c (squared error against p, or cosine similarity) -> top-k.  The catalog is sharded by rows
across ranks; each rank sweeps its shard (codae_score_topk) and the per-rank lists are all-gathered and merged
deterministically (codae_topk_merge).
"""
import torch

from codae import _C

METRICS = {"sqerr": _C.METRIC_SQERR, "cosine": _C.METRIC_COSINE}


def shard_rows(n_rows, world_size, rank):
    """Contiguous ceil(N/G)-row shards: (row_offset, n_local)."""
    per = (n_rows + world_size - 1) // world_size
    lo = min(rank * per, n_rows)
    return lo, min(per, n_rows - lo)


class ComplementarityScorer:

    def __init__(self, catalog, embedding_size, metric="sqerr", k=10, inv_scale=1.0, row_offset=0, process_group=None):
        """catalog: [n_local, E] fp32 or bf16 CUDA tensor -- this rank's shard (rows row_offset.. of the global
        catalog).  inv_scale multiplies catalog values before the squared error (1/dataset.scale when the
        catalog is un-scaled like dataset.data_per_category)."""
        if metric not in METRICS:
            raise Exception("Unknown metric.")
        if not catalog.is_cuda:
            raise RuntimeError("codae: catalog is not on a CUDA device; the B200 path has no CPU fallback")
        self.catalog = catalog
        self.E = embedding_size
        self.metric = METRICS[metric]
        self.k = k
        self.inv_scale = float(inv_scale)
        self.row_offset = int(row_offset)
        self.pg = process_group
        self._ws = {}

    def _workspace(self, Q):
        ws = self._ws.get(Q)
        if ws is None:
            dev = self.catalog.device
            ws = (_C.score_topk_workspace(dev, Q, self.k), torch.empty((Q, self.k), dtype=torch.float32, device=dev),
                  torch.empty((Q, self.k), dtype=torch.int64, device=dev))
            self._ws[Q] = ws
        return ws

    def topk_local(self, query):
        """(scores [Q,k] f32, global indices [Q,k] i64) of this shard; asynchronous, reuses its buffers."""
        q = query.to(torch.float32).contiguous()
        ws, s, i = self._workspace(q.shape[0])
        _C.score_topk(self.catalog, self.E, self.row_offset, q, self.inv_scale, self.metric, self.k, s, i, ws)
        return s, i

    def topk(self, query):
        """Global top-k: local sweep, all-gather of k x 12 bytes per query, deterministic merge."""
        s, i = self.topk_local(query)
        import torch.distributed as dist
        if self.pg is None and not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return s, i
        G = dist.get_world_size(self.pg)
        gs = torch.empty((G,) + tuple(s.shape), dtype=s.dtype, device=s.device)
        gi = torch.empty((G,) + tuple(i.shape), dtype=i.dtype, device=i.device)
        dist.all_gather_into_tensor(gs, s.contiguous(), group=self.pg)
        dist.all_gather_into_tensor(gi, i.contiguous(), group=self.pg)
        out_s, out_i = torch.empty_like(s), torch.empty_like(i)
        _C.topk_merge(gs, gi, self.metric, out_s, out_i)
        return out_s, out_i


def predict_slot(model, outfits, slot, embedding_size):
    """Zero slot `slot` of the (scaled) outfit rows, run the DAE, return the reconstructed slot [Q, E]."""
    E = embedding_size
    mask = torch.ones_like(outfits)
    mask[:, slot * E:(slot + 1) * E] = 0
    with torch.no_grad():
        y = model(model.corrupt(input_data=outfits, mask=mask))
    return y[:, slot * E:(slot + 1) * E].contiguous()
