"""Stage IV -- complementarity inference (the entry point the reference README names, README.md:14-16,30-32,
whose script is absent from the tree; semantics follow RankingLoss, codae/tool/metering.py:46-79).

context outfit with slot c zeroed -> DAE forward -> predicted embedding p of slot c -> score every catalog
item of category c (squared error against p, or cosine similarity) -> top-k.  The catalog is sharded by rows
across ranks; each rank sweeps its shard (codae_score_topk) and the per-rank lists are all-gathered and merged
deterministically (codae_topk_merge).

SwapScorer is the other reading of "scores candidate item swaps by reconstruction error": every candidate is
substituted into the slot, the whole swapped outfit goes through the DAE (tensor-core GEMMs batched over the
candidates) and the swap is scored by the reconstruction error of the swapped outfit.
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


class SwapScorer:
    """Scores the swap "candidate j into slot c" by || DAE(outfit_j) - outfit_j ||^2 over all io dimensions (lower
    is better) and returns the best k swaps.  Candidates are processed in chunks of `chunk` rows: codae_swap_build
    writes the swapped outfits straight into the first activation buffer, the model's Linear layers run on them,
    codae_swap_error_topk reduces each chunk to k candidates and codae_topk_merge combines chunks (and ranks)."""

    def __init__(self, model, catalog, embedding_size, k=10, inv_scale=1.0, row_offset=0, chunk=8192, process_group=None):
        if not catalog.is_cuda:
            raise RuntimeError("codae: catalog is not on a CUDA device; the B200 path has no CPU fallback")
        model._require_cuda()
        self.model, self.catalog, self.E, self.k = model, catalog, embedding_size, k
        self.inv_scale, self.row_offset, self.chunk, self.pg = float(inv_scale), int(row_offset), int(chunk), process_group
        self.io = model.dims[0][0]
        if self.io % self.E != 0 or catalog.shape[1] < self.E:
            raise Exception("Catalog width or embedding size does not fit the model input.")
        dev = catalog.device
        self._ws = _C.score_topk_workspace(dev, 1, k)
        self._acts = None

    def _buffers(self):
        if self._acts is None:
            m, dev = self.model, self.catalog.device
            eng = m.engine_dtype()
            adt = m.act_dtype(eng)
            acts = [m.new_activation(self.chunk, self.io, adt, dev)]
            for l, (i, o) in enumerate(m.dims):
                acts.append(m.new_activation(self.chunk, o, torch.float32 if l == len(m.dims) - 1 else adt, dev))
            # fp32-parity engine: the swap rows are built in fp32 and split into the engine's three bf16 planes
            stage = m.new_activation(self.chunk, self.io, torch.float32, dev) if eng == _C.F32X3 else None
            self._acts = (eng, acts, stage)
        return self._acts

    def topk_local(self, outfit, slot):
        """outfit: [io] scaled fp32 row.  Returns (errors [k], global candidate indices [k]) of this shard."""
        m, dev = self.model, self.catalog.device
        if not 0 <= slot < self.io // self.E:
            raise Exception("Slot out of range.")
        outfit = outfit.to(torch.float32).contiguous().view(-1)
        m.sync_weights()          # a sharded data-parallel trainer gathers the master weights first (collective)
        eng, acts, stage = self._buffers()
        m.refresh_shadow()
        wflat = m.gemm_weights(eng)
        n = self.catalog.shape[0]
        n_chunks = max(1, (n + self.chunk - 1) // self.chunk)
        ls = torch.full((n_chunks, 1, self.k), float("inf"), dtype=torch.float32, device=dev)
        li = torch.full((n_chunks, 1, self.k), -1, dtype=torch.int64, device=dev)
        for c in range(n_chunks if n else 0):
            first = c * self.chunk
            B = min(self.chunk, n - first)
            if stage is None:
                _C.swap_build(outfit, self.catalog, first, B, self.E, slot, self.io, self.inv_scale, acts[0])
            else:
                _C.swap_build(outfit, self.catalog, first, B, self.E, slot, self.io, self.inv_scale, stage)
                _C.split_x3(stage, acts[0])
            for l, (i, o) in enumerate(m.dims):
                _C.linear_fwd(acts[l], m.aug_view(wflat, l), None, acts[l + 1], B, o, (i + 7) // 8 * 8 + 1,
                              _C.ACT_RELU if m.relu[l] else _C.ACT_NONE, eng)
            _C.swap_error_topk(outfit, self.catalog, first, B, self.E, slot, self.io, self.inv_scale, acts[-1],
                               self.row_offset, self.k, ls[c, 0], li[c, 0], self._ws)
        out_s = torch.empty((1, self.k), dtype=torch.float32, device=dev)
        out_i = torch.empty((1, self.k), dtype=torch.int64, device=dev)
        _C.topk_merge(ls, li, _C.METRIC_SQERR, out_s, out_i)
        return out_s[0], out_i[0]

    def topk(self, outfit, slot):
        s, i = self.topk_local(outfit, slot)
        import torch.distributed as dist
        if self.pg is None and not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return s, i
        G = dist.get_world_size(self.pg)
        gs = torch.empty((G, 1, self.k), dtype=s.dtype, device=s.device)
        gi = torch.empty((G, 1, self.k), dtype=i.dtype, device=i.device)
        dist.all_gather_into_tensor(gs, s.view(1, 1, -1).contiguous(), group=self.pg)
        dist.all_gather_into_tensor(gi, i.view(1, 1, -1).contiguous(), group=self.pg)
        out_s, out_i = torch.empty_like(gs[0]), torch.empty_like(gi[0])
        _C.topk_merge(gs, gi, _C.METRIC_SQERR, out_s, out_i)
        return out_s[0], out_i[0]
