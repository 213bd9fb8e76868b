"""FusedStep -- the training-loop body of the reference scripts as a fixed sequence of libcodae_b200 kernels.

Replaces, per step (script/train_dae_on_embedding.py:198-223; script/train_dae_on_abalone.py:206-236):
  get_masks + corrupt      -> codae_corrupt_fwd      (row gather + slot mask, mask never materialised)
  model(c_input)           -> L x codae_linear_fwd   (bias inside the contraction, ReLU in the epilogue)
  criterion + first bwd    -> codae_mse_loss_fwd_bwd (loss, dL/dy and the full/partial monitor sums, one pass)
                              or codae_mixed_loss_fwd_bwd + codae_mixed_monitor (abalone)
  loss.backward()          -> L x codae_linear_wgrad (bias gradient = the constant-1 column of the augmented
                              contraction, no column-sum kernel), (L-1) x codae_linear_dgrad
  [data parallel]          -> NCCL all-reduce of the flat gradient buffer, issued per layer on a communication stream
                              as soon as that layer's wgrad has finished (overlaps the rest of the backward pass)
  clip_grad_norm_ + Adam   -> codae_clip_adam_step: one cooperative launch over the flat buffers (norm, grid barrier,
                              update); or codae_grad_sqnorm + codae_adam_step; or (wgrad_sqnorm=True, single GPU,
                              tensor-core engine) codae_linear_wgrad_sq + codae_adam_step_partials: the weight-gradient
                              kernels leave sum(dW^2) behind and the optimizer never reads g for the norm
No host synchronisation happens inside a step; monitors stay on the device until read_monitors().
The whole sequence can be captured once per batch size into a CUDA graph (use_graph=True).
"""
import os

import torch

from codae import _C
from codae.tool.metering import arch_tables


def _round_up(x, m):
    return (x + m - 1) // m * m


class FusedStep:

    BUCKET_BYTES = 32 << 20     # gradient all-reduce bucket (bytes of fp32 gradients)

    def __init__(self, model, corrupter, data, lr, weight_decay, clip=True, betas=(0.9, 0.999), eps=1e-8,
                 max_norm=1.0, world_size=1, process_group=None, use_graph=False, mixed=None, overlap_allreduce=True,
                 fused_clip_adam=True, wgrad_sqnorm=None, tiny_mlp=None, dp_mode=None):
        """model: FlatMLP on a CUDA device; corrupter: codae.tool.Corrupter; data: resident [N, io] fp32 CUDA
        tensor (dataset.data).  mixed: None for the embedding loss (MSE mean over all elements) or a dict
        {arch, weight, norm_scale, norm_min, norm_first} for the abalone CombinedCriterion loss + monitors.
        world_size > 1: gradients are summed across `process_group` (each rank computes dL/dy with the GLOBAL
        batch size, so the sum is the global-batch gradient the single-process reference would see).
        dp_mode (world_size > 1): "peer" (default) -- codae_dp_adam_step: reduce-scatter + clip + Adam on this rank's shard +
        all-gather of the new weights as one kernel over NVLink peer memory (gradient / weight buffers live in torch symmetric
        memory; moments are shard-sized; the fp32 master weights of the other shards are gathered by flush()); "nccl" -- NCCL
        all-reduce of the flat gradient buffer in buckets overlapped with the backward pass, whole update on every replica."""
        if model.flat is None:
            raise RuntimeError("codae: FusedStep needs model.to(cuda_device) first; there is no CPU fallback")
        if not data.is_cuda:
            raise RuntimeError("codae: FusedStep needs dataset.to(cuda_device) first; there is no CPU fallback")
        self.model, self.corrupter, self.data = model, corrupter, data
        self.lr, self.wd, self.clip, self.betas, self.eps, self.max_norm = lr, weight_decay, clip, betas, eps, max_norm
        self.world_size, self.pg = world_size, process_group
        self.use_graph = use_graph
        self.mixed = mixed
        self.overlap_allreduce = overlap_allreduce
        self.fused_clip_adam = fused_clip_adam
        self.wgrad_sqnorm = wgrad_sqnorm
        dev = model.flat.device
        self.dev = dev
        self.io = model.dims[0][0]
        self.eng = model.engine_dtype()
        self._compute_dtype = model.compute_dtype      # buffers and the weight shadow are built for this engine: see _check_engine
        self.adt = model.act_dtype(self.eng)       # torch.float32 / torch.bfloat16 / "x3" (three bf16 planes per fp32 tensor)
        self.tc = self.eng != _C.F32               # a tensor-core engine (bf16, or fp32 parity on bf16 triples)
        n = model.flat.numel()
        if self.tc:
            model.refresh_shadow()
        self.dp_mode = None
        self.dp_nvls = False          # peer mode: reduce / broadcast through NVSwitch multicast addresses
        self._master_stale = False
        self._peers = None
        if world_size > 1:
            if mixed is not None:
                raise RuntimeError("codae: the tabular (CombinedCriterion) step is single-GPU (792 parameters: replicas only)")
            self.dp_mode = self._setup_data_parallel(dp_mode, n)
        if self.dp_mode != "peer":
            self.gflat = torch.zeros(n, dtype=torch.float32, device=dev)
            self.m = torch.zeros(n, dtype=torch.float32, device=dev)
            self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.norm_ws = _C.sqnorm_workspace(dev)
        self.loss_ws = _C.loss_workspace(dev)
        self.acc = torch.zeros(4, dtype=torch.float64, device=dev)   # {sum full, sum partial, rows, last step's sum}
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_count = 0
        self.kernel_launches = 0
        if mixed is not None:
            # the tabular loss kernel writes dL/dy in f32 and normalises by the local batch (RMSE over the batch couples all rows)
            if self.eng != _C.F32:
                raise RuntimeError("codae: the tabular (CombinedCriterion) loss runs on the fp32 engine only")
            V = len(mixed["arch"])
            self.var_tables = arch_tables(mixed["arch"], dev)
            self.weight = torch.tensor(list(mixed["weight"]), dtype=torch.float32, device=dev)
            self.norm_scale = None if mixed.get("norm_scale") is None else mixed["norm_scale"].to(dev, torch.float32).contiguous()
            self.norm_min = None if mixed.get("norm_min") is None else mixed["norm_min"].to(dev, torch.float32).contiguous()
            self.norm_first = int(mixed.get("norm_first", self.io))
            self.mixed_loss = torch.zeros(1 + V, dtype=torch.float32, device=dev)
            self.mixed_acc = torch.zeros(2 + 2 * corrupter.k_max * V, dtype=torch.float64, device=dev)
        for l, lin in enumerate(model.linears()):   # .grad views into the flat gradient buffer
            lin.weight.grad = model.weight_view(self.gflat, l)
            lin.bias.grad = model.bias_view(self.gflat, l)
        lay, total = model.layout()
        self._layer_span = [(lay[l][0], lay[l + 1][0] if l + 1 < len(lay) else total) for l in range(len(lay))]
        if wgrad_sqnorm is None:
            # on wherever it applies (embedding.yaml step: 0.3446 -> 0.3367 ms); CODAE_WGRAD_SQNORM=0 switches it off
            self.wgrad_sqnorm = (os.environ.get("CODAE_WGRAD_SQNORM", "1") != "0" and world_size == 1 and self.tc
                                 and fused_clip_adam)
        lay = model.layout()[0]
        tiny_ok = (self.eng == _C.F32 and world_size == 1 and len(model.dims) <= _C.TINY_MAX_LAYERS
                   and max(max(i, o) for i, o in model.dims) <= 64)
        if tiny_mlp is None:
            # tabular widths: ONE forward and ONE backward launch (codae_tiny_mlp_fwd / _bwd), 22 -> 7 launches per abalone step;
            # CODAE_TINY_MLP=0 keeps the per-layer FFMA kernels (A/B)
            tiny_mlp = os.environ.get("CODAE_TINY_MLP", "1") != "0" and tiny_ok
        if tiny_mlp and not tiny_ok:
            raise RuntimeError("codae: tiny_mlp needs the fp32 engine, a single GPU, at most %d layers of width <= 64" % _C.TINY_MAX_LAYERS)
        self.tiny_mlp = bool(tiny_mlp)
        self._tiny_layers = [_C.TinyLayer(lay[l][0], lay[l][1], lay[l][2], i, o, 1 if model.relu[l] else 0)
                             for l, (i, o) in enumerate(model.dims)] if self.tiny_mlp else None
        if self.wgrad_sqnorm and (world_size > 1 or not self.tc):
            raise RuntimeError("codae: wgrad_sqnorm needs a single GPU (the norm of a data-parallel run is taken after the "
                               "all-reduce) and the tensor-core engine")
        self._comm_stream = torch.cuda.Stream(device=dev) if world_size > 1 else None
        self._wgrad_stream = torch.cuda.Stream(device=dev)
        self._bufs = {}
        self._graphs = {}
        self._calls = {}

    # ---- data parallel plumbing ----------------------------------------------------------------------
    def _setup_data_parallel(self, want, n):
        """Chooses the data-parallel schedule.  "peer": the flat gradient buffer, the weight buffer the GEMMs read and a signal
        pad are allocated in torch symmetric memory and rendezvoused over the process group, which yields every rank's buffer as
        a device pointer valid in this process (PyTorch is the plumbing; the data path is codae_dp_adam_step).  All ranks take
        the same decision: if the rendezvous fails anywhere and "peer" was not requested explicitly, everyone uses "nccl"."""
        import torch.distributed as dist
        explicit = want is not None or os.environ.get("CODAE_DP_MODE") is not None
        want = want or os.environ.get("CODAE_DP_MODE", "peer")
        if want not in ("peer", "nccl"):
            raise RuntimeError("codae: dp_mode must be 'peer' or 'nccl'")
        if want == "nccl":
            return "nccl"
        group = self.pg if self.pg is not None else dist.group.WORLD
        dev, model = self.dev, self.model
        ok, err = 1, None
        try:
            if self.world_size > _C.DP_MAX_WORLD or n % 8:
                raise RuntimeError("peer mode supports at most %d ranks and flat sizes that are multiples of 8" % _C.DP_MAX_WORLD)
            import torch.distributed._symmetric_memory as symm
            rank = dist.get_rank(group)
            g = symm.empty(n, dtype=torch.float32, device=dev)
            g.zero_()
            w = symm.empty(n, dtype=torch.bfloat16 if self.eng == _C.BF16 else torch.float32, device=dev)
            w.copy_(model.flat_bf16 if self.eng == _C.BF16 else model.flat)
            sig = symm.empty(_C.DP_SIGNAL_BYTES // 8, dtype=torch.int64, device=dev)
            sig.zero_()
            handles = [symm.rendezvous(t, group) for t in (g, w, sig)]
            ptrs = [[int(x) for x in h.buffer_ptrs] for h in handles]
            assert all(len(x) == self.world_size for x in ptrs) and ptrs[0][rank] == g.data_ptr()
        except Exception as ex:            # noqa: BLE001 -- reported below, every rank must reach the vote
            ok, err = 0, ex
        vote = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN, group=group)
        if int(vote.item()) == 0:
            if explicit:
                raise RuntimeError("codae: dp_mode='peer' is not available on this system: %r" % (err,))
            import warnings
            warnings.warn("codae: symmetric-memory rendezvous failed (%r); data parallel falls back to the NCCL all-reduce schedule" % (err,))
            return "nccl"
        if self.eng == _C.BF16:
            model.flat_bf16 = w
        else:
            model.rebind_flat(w)
        self.gflat = g
        S = _C.dp_shard_elems(n, self.world_size)
        self._shard = (min(n, rank * S), min(n, (rank + 1) * S), S)
        self.m = torch.zeros(S, dtype=torch.float32, device=dev)            # moments exist for this rank's shard only
        self.v = torch.zeros(S, dtype=torch.float32, device=dev)
        # NVSwitch multicast (NVLS) addresses of the gradient and weight buffers when the fabric offers them: in-switch reduction
        mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in handles[:2]]
        have = torch.tensor([1 if all(mc) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(have, op=dist.ReduceOp.MIN, group=group)            # every rank takes the same path
        if int(have.item()) == 0:
            mc = [0, 0]
        self.dp_nvls = bool(all(mc))
        self._peers = _C.dp_peers(self.world_size, rank, *ptrs, grads_mc=mc[0], w_mc=mc[1])
        self._symm = (g, w, sig, handles)                                    # keep the mappings alive
        self.dp_ws = _C.dp_workspace(dev)
        self._dp_group = group
        model._flush_hook = self.flush                                       # reading the weights gathers the master shards first
        torch.cuda.synchronize()
        dist.barrier(group=group)                                            # every pad is zeroed before anyone's first step
        return "peer"

    def _gather_master(self):
        """COLLECTIVE: all-gather of the fp32 master-weight shards (peer mode on the tensor-core engine keeps only this rank's
        shard of model.flat current; the bf16 weights every GEMM reads are always complete)."""
        import torch.distributed as dist
        lo, hi, S = self._shard
        n = self.model.flat.numel()
        mine = torch.zeros(S, dtype=torch.float32, device=self.dev)
        mine[:hi - lo] = self.model.flat[lo:hi]
        full = torch.empty(S * self.world_size, dtype=torch.float32, device=self.dev)
        dist.all_gather_into_tensor(full, mine, group=self._dp_group)
        self.model.flat.copy_(full[:n])
        self._master_stale = False

    # ---- buffers ------------------------------------------------------------------------------------
    def _new_grad(self, B, w):
        if self.adt == "x3":
            return _C.new_x3((B, w), self.dev)
        return torch.zeros((B, w), dtype=self.adt, device=self.dev)

    def _buffers(self, B):
        b = self._bufs.get(B)
        if b is None:
            dev, adt, dims = self.dev, self.adt, self.model.dims
            M = self.model
            wmax = max(M.act_width(max(i, o)) for i, o in dims)
            # tabular (mixed) mode keeps one pitch for every buffer: the mixed-loss kernel takes a single ld
            width = (lambda w: wmax) if self.mixed is not None else (lambda w: None)
            acts = [M.new_activation(B, dims[0][0], adt, dev, width(dims[0][0]))]
            for l, (i, o) in enumerate(dims):
                last = l == len(dims) - 1
                acts.append(M.new_activation(B, o, torch.float32 if last else adt, dev, width(o)))
            b = dict(acts=acts,
                     g0=self._new_grad(B, wmax), g1=self._new_grad(B, wmax), g2=self._new_grad(B, wmax),
                     mask_id=torch.zeros(B, dtype=torch.int32, device=dev),
                     idx=torch.zeros(B, dtype=torch.int64, device=dev),
                     x=torch.zeros((B, wmax), dtype=torch.float32, device=dev) if self.mixed is not None else None,
                     mon=torch.zeros((B, len(self.mixed["arch"])), dtype=torch.float32, device=dev) if self.mixed is not None else None)
            if self.wgrad_sqnorm:
                # one slot per CTA of every weight-gradient launch of this batch size
                slots = [_C.linear_wgrad_sq_slots(dev, B, o, _round_up(i, 8) + 1, self.eng) for i, o in dims]
                if min(slots) < 1:
                    raise RuntimeError("codae: wgrad_sqnorm: the tensor-core engine cannot tile every layer of this model")
                offs = [0]
                for c in slots:
                    offs.append(offs[-1] + c)
                b["sq_off"] = offs
                b["sq_partials"] = torch.zeros(offs[-1], dtype=torch.float64, device=dev)
            self._bufs[B] = b
        return b

    # ---- the kernel sequence ----------------------------------------------------------------------------
    def _enqueue(self, B, b, run, global_batch, data, idx, table, train=True):
        model, dims, eng = self.model, self.model.dims, self.eng
        L = len(dims)
        _, bits, col_var, nmiss = self.corrupter.device_tables()
        acts = b["acts"]
        n = 0
        _C.corrupt_fwd(data, idx, B, table, run, bits, col_var, self.io, acts[0], b["x"], b["mask_id"]); n += 1
        wflat = model.gemm_weights(eng)
        tiny = self.tiny_mlp and len({a.stride(0) for a in acts}) == 1
        if tiny:
            # tabular widths: every layer of the forward pass in one launch (rows are independent: one CTA per 32 rows)
            _C.tiny_mlp_fwd(self._tiny_layers, model.flat, acts, B); n += 1
        else:
            for l, (i, o) in enumerate(dims):
                _C.linear_fwd(acts[l], model.aug_view(wflat, l), None, acts[l + 1], B, o, _round_up(i, 8) + 1,
                              _C.ACT_RELU if model.relu[l] else _C.ACT_NONE, eng); n += 1
        y = acts[L]
        o_last = dims[L - 1][1]
        gbuf = [b["g0"], b["g1"], b["g2"]]            # dL/d(output of layer l) lives in gbuf[l % 3]
        g = gbuf[(L - 1) % 3][..., :_round_up(o_last, 8)]
        if self.mixed is None:
            _C.mse_loss_fwd_bwd(data, idx, y, b["mask_id"], bits, col_var, B, self.io, 2.0 / (global_batch * self.io),
                                g if train else None, self.acc, self.loss_ws); n += 1
        else:
            pos, size, typ = self.var_tables
            _C.mixed_loss_fwd_bwd(b["x"], y, pos, size, typ, self.weight, g, self.mixed_loss); n += 1
            _C.mixed_monitor(b["x"], y, pos, size, typ, self.norm_scale, self.norm_min, self.norm_first, b["mask_id"], bits,
                             nmiss, self.corrupter.k_max, b["mon"], self.mixed_acc); n += 1
        if not train:
            return n
        if tiny:
            # every weight gradient and the input-gradient chain in one single-CTA launch, then the update
            _C.tiny_mlp_bwd(self._tiny_layers, model.flat, self.gflat, acts, gbuf, B); n += 1
            return n + self._enqueue_update(b.get("sq_partials"))
        # Backward.  The input-gradient chain dgrad(L-1) -> ... -> dgrad(1) is the critical path; every weight gradient
        # only needs dL/d(out_l) and the stored activation, so wgrad(l) runs on a second stream next to dgrad(l).
        # A tcgen05 kernel owns its SM (one CTA per SM whatever its resources), so the two only overlap when their grids fit the
        # 148 SMs together: the library launches small-batch input gradients as 12 clusters of 8 = 96 CTAs for that reason
        # (csrc/gemm_tcgen05.cu: pick_bn; embedding.yaml bf16 0.341 -> 0.321 ms/step).
        overlap_comm = self.world_size > 1 and self.overlap_allreduce and self.dp_mode == "nccl"
        if overlap_comm:
            import torch.distributed as dist
        main = torch.cuda.current_stream()
        # second stream only for the tensor-core engine: its small-batch kernels leave most SMs idle; the FFMA engine's
        # kernels fill the GPU and only slow each other down (measured 7.1 vs 5.4 ms/step)
        # (two side streams were measured no better than one; at large batch both contractions fill the GPU and
        # concurrency only disturbs L2 locality, so the second stream is a small-batch device)
        side = self._wgrad_stream if (self.tc and B <= 1024) else main
        wdone = [None] * L
        bucket_hi = None
        for l in range(L - 1, -1, -1):
            i, o = dims[l]
            gl = gbuf[l % 3][..., :_round_up(o, 8)]
            if side is not main:
                ready = torch.cuda.Event()
                ready.record(main)                      # dL/d(out_l) has been produced (loss or dgrad(l+1))
                side.wait_event(ready)
            with torch.cuda.stream(side):
                if self.wgrad_sqnorm:
                    sq = b["sq_partials"][b["sq_off"][l]:b["sq_off"][l + 1]]
                    _C.linear_wgrad_sq(gl, acts[l], model.aug_view(self.gflat, l), B, o, _round_up(i, 8) + 1, eng, sq)
                else:
                    _C.linear_wgrad(gl, acts[l], model.aug_view(self.gflat, l), None, B, o, _round_up(i, 8) + 1, eng)
                n += 1
                wdone[l] = torch.cuda.Event()
                wdone[l].record(side)
            if overlap_comm:
                # Finished layers form a contiguous tail of the flat gradient buffer.  Reduce it on the communication
                # stream while the remaining backward GEMMs run -- in buckets of at least BUCKET_BYTES: per-layer calls
                # (9.4 MB at 1536 wide) cost more in NCCL latency than they hide (measured 0.77 vs 0.70 ms/step, 2 GPUs).
                lo = self._layer_span[l][0]
                if bucket_hi is None:
                    bucket_hi = self._layer_span[l][1]
                if (bucket_hi - lo) * 4 >= self.BUCKET_BYTES or l == 0:
                    comm = self._comm_stream
                    comm.wait_event(wdone[l])
                    with torch.cuda.stream(comm):
                        dist.all_reduce(self.gflat[lo:bucket_hi], op=dist.ReduceOp.SUM, group=self.pg)
                    bucket_hi = None
            if l > 0:
                if l + 2 <= L - 1 and side is not main:
                    main.wait_event(wdone[l + 2])       # dgrad(l) overwrites the buffer wgrad(l+2) was reading
                gp = gbuf[(l - 1) % 3][..., :_round_up(i, 8)]
                _C.linear_dgrad(gl, model.weight_view(wflat, l), acts[l] if model.relu[l - 1] else None, gp, B, o, i, eng); n += 1
        if side is not main:
            main.wait_stream(side)
        if overlap_comm:
            main.wait_stream(self._comm_stream)
        elif self.world_size > 1 and self.dp_mode == "nccl":
            import torch.distributed as dist
            dist.all_reduce(self.gflat, op=dist.ReduceOp.SUM, group=self.pg)
        return n + self._enqueue_update(b.get("sq_partials"))

    def _enqueue_update(self, sq_partials=None):
        """Step counter + clip + Adam over the flat buffers (after the gradients are final)."""
        model, eng = self.model, self.eng
        n = 0
        _C.counter_add(self.step_dev, 1); n += 1
        pb = model.gemm_weights(eng) if self.tc else None       # the copy of the weights the tensor-core GEMMs read
        if self.dp_mode == "peer":
            # gradients of every rank -> this rank's shard: reduce, global clip scale, Adam, new weights to every rank (one kernel)
            wdt = _C.BF16 if eng == _C.BF16 else _C.F32
            _C.dp_adam_step(self._peers, model.flat, self.m, self.v, wdt, self.lr, self.betas[0], self.betas[1], self.eps, self.wd, 0,
                            self.max_norm if self.clip else -1.0, self.sqnorm, self.dp_ws, 1.0, self.step_dev); n += 1
            self._master_stale = eng == _C.BF16
            if eng == _C.F32X3:
                # fp32-parity engine: the peers exchanged fp32 weights (complete when the kernel ends); re-split them locally
                _C.split_x3(model.flat, model.flat_x3); n += 1
            return n
        if sq_partials is not None:
            # the weight-gradient launches of this step left sum(dW^2) per CTA: no norm pass, no grid barrier
            _C.adam_step_partials(model.flat, self.gflat, self.m, self.v, pb, self.lr, self.betas[0], self.betas[1], self.eps,
                                  self.wd, 0, self.max_norm if self.clip else -1.0, sq_partials, self.sqnorm, 1.0,
                                  self.step_dev); n += 1
        elif self.fused_clip_adam:
            # ||g||^2, clip scale and Adam in one cooperative launch (the second read of g comes from L2).  Also without
            # clipping (max_norm < 0 -> scale 1): inside the step it measured faster than the plain Adam kernel
            # (modanet: 0.305 vs 0.319 ms/step) and the gradient norm comes out as a monitor.
            _C.clip_adam_step(model.flat, self.gflat, self.m, self.v, pb, self.lr, self.betas[0], self.betas[1], self.eps,
                              self.wd, 0, self.max_norm if self.clip else -1.0, self.sqnorm, self.norm_ws, 1.0,
                              self.step_dev); n += 1
        else:
            if self.clip:
                _C.grad_sqnorm(self.gflat, self.sqnorm, self.norm_ws); n += 1
            _C.adam_step(model.flat, self.gflat, self.m, self.v, pb, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                         0, self.max_norm if self.clip else -1.0, self.sqnorm if self.clip else None, 1.0, self.step_dev); n += 1
        return n

    def _step_without_samples(self):
        """This rank holds no sample of the (ragged, last) global batch: contribute zero gradients to the reduction and
        take part in the same update as every other rank."""
        self.gflat.zero_()
        if self.dp_mode == "nccl":
            import torch.distributed as dist
            if self.overlap_allreduce:      # same bucket boundaries as the ranks that do have samples
                L, hi = len(self.model.dims), None
                for l in range(L - 1, -1, -1):
                    lo = self._layer_span[l][0]
                    hi = self._layer_span[l][1] if hi is None else hi
                    if (hi - lo) * 4 >= self.BUCKET_BYTES or l == 0:
                        dist.all_reduce(self.gflat[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
                        hi = None
            else:
                dist.all_reduce(self.gflat, op=dist.ReduceOp.SUM, group=self.pg)
        self.kernel_launches = self._enqueue_update()

    def _check_engine(self):
        if self.model.compute_dtype != self._compute_dtype:
            raise RuntimeError("codae: model.set_compute_dtype(%r) after this FusedStep was built for %r; build a new FusedStep"
                               % (self.model.compute_dtype, self._compute_dtype))

    # ---- public API ------------------------------------------------------------------------------------------
    def step(self, batch_idx, run=0, global_batch=None, staged=None):
        """One training step on observations `batch_idx` (int64 CUDA tensor [B]).
        staged=(rows [B, io] f32, table_rows [B, nb_run] i16): host-staged batch (end-to-end path): the kernels
        then read the staged rows instead of gathering from the resident dataset."""
        self._check_engine()
        B = int(batch_idx.numel()) if staged is None else int(staged[0].shape[0])
        gb = B * self.world_size if global_batch is None else global_batch
        if B == 0:
            self.step_count += 1
            self._step_without_samples()
            return
        b = self._buffers(B)
        table = self.corrupter.device_tables()[0]
        if staged is not None:
            data, idx, table = staged[0], None, staged[1]
        else:
            data, idx = self.data, b["idx"]
            idx.copy_(batch_idx, non_blocking=True)
        self.step_count += 1
        key = (B, run, gb, None if staged is None else (staged[0].data_ptr(), staged[1].data_ptr()))
        if not self.use_graph:
            self.kernel_launches = self._enqueue(B, b, run, gb, data, idx, table)
            return
        calls = self._calls.get(key, 0)
        self._calls[key] = calls + 1
        if calls == 0:            # first step of this shape runs eagerly (lazy attribute setup, warm caches)
            self.kernel_launches = self._enqueue(B, b, run, gb, data, idx, table)
            return
        gr = self._graphs.get(key)
        if gr is None:
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            self._static = getattr(self, "_static", {})
            self._static[key] = (data, idx, table)
            with torch.cuda.graph(gr):
                self.kernel_launches = self._enqueue(B, b, run, gb, data, idx, table)
            self._graphs[key] = gr
        gr.replay()

    def train_steps(self, idx_rows, run=0, global_batch=None, graph_steps=32):
        """n consecutive training steps, idx_rows: int64 CUDA tensor [n, B] (row j = the observations of step j).  Chunks of
        `graph_steps` steps are ONE CUDA-graph launch each (no per-step Python call, no per-step index copy); the remainder runs
        through step().  Same kernels on the same values as step(): bit-identical weights and monitors.
        (Measured and rejected on a B200, profiles/r02_notes.md visit 8: running the optimizer update of step s beside the forward
        pass of step s + 1 inside these graphs, each forward GEMM gated on per-layer completion flags of the update.)"""
        self._check_engine()
        n, Bl = int(idx_rows.shape[0]), int(idx_rows.shape[1])
        gb = Bl * self.world_size if global_batch is None else global_batch
        s = 0
        if n and not self._calls.get(("epoch_warm", Bl)):
            # the first step of a batch size runs eagerly (lazy attribute setup) -- it is a real step
            self._calls[("epoch_warm", Bl)] = 1
            saved, self.use_graph = self.use_graph, False
            try:
                self.step(idx_rows[0], run=run, global_batch=gb)
            finally:
                self.use_graph = saved
            s = 1
        K = max(1, int(graph_steps))
        b = self._buffers(Bl)
        table = self.corrupter.device_tables()[0]
        while n - s >= K:
            key = ("steps", Bl, run, gb, K)
            ent = self._graphs.get(key)
            if ent is None:
                win = torch.zeros((K, Bl), dtype=torch.int64, device=self.dev)
                win.copy_(idx_rows[s:s + K])
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    for j in range(K):
                        self.kernel_launches = self._enqueue(Bl, b, run, gb, self.data, win[j], table)
                ent = (gr, win)
                self._graphs[key] = ent
            gr, win = ent
            win.copy_(idx_rows[s:s + K], non_blocking=True)
            gr.replay()
            self.step_count += K
            s += K
        for j in range(s, n):
            self.step(idx_rows[j], run=run, global_batch=gb)
        return n

    def train_epoch(self, train_idx, batch_size, generator=None, rank=0, run=0, graph_steps=32):
        """One epoch with the sampler on the device (train_dae_on_embedding.py:118-128: SubsetRandomSampler + DataLoader(batch_size),
        i.e. a fresh permutation of the training observations, consecutive batches, ragged last batch): the permutation is drawn
        on the GPU and the full batches go through train_steps() -- `graph_steps` consecutive steps per CUDA-graph launch, no
        per-step Python call, no per-step host->device index copy.
        train_idx: int64 CUDA tensor of observation ids; batch_size: GLOBAL batch (data parallel: every rank draws the same
        permutation from an identically seeded `generator` and takes rows rank::world_size of each batch, as the host sampler
        of script/_common.py does).  Returns the number of steps taken."""
        dev, W = self.dev, self.world_size
        n = int(train_idx.numel())
        if n:
            lo, hi = int(train_idx.min()), int(train_idx.max())       # one read-back per epoch: the kernels gather rows unchecked
            if lo < 0 or hi >= self.data.shape[0]:
                raise Exception("Observation index out of range: [%d, %d] for a dataset of %d rows." % (lo, hi, self.data.shape[0]))
        perm = train_idx[torch.randperm(n, device=dev, generator=generator)]
        local, tail = self.epoch_rows(perm, batch_size, W, rank)
        steps = 0
        if local is not None:
            steps = self.train_steps(local, run=run, global_batch=batch_size, graph_steps=graph_steps)
        # ragged tail (or a global batch the ranks cannot split evenly): the per-step path
        for rows, gb in tail:
            self.step(rows, run=run, global_batch=gb)
            steps += 1
        return steps

    @staticmethod
    def epoch_rows(perm, batch_size, world_size=1, rank=0):
        """Splits an epoch's permutation the way the host sampler does (script/_common.py: epoch_batches): consecutive global batches
        of batch_size observations, rank r takes elements r::world_size of each.  Returns (local [n_full, batch_size // world_size]
        or None, [(rows of this rank, global batch size)] for the batches that are ragged or not divisible by world_size)."""
        n = int(perm.numel())
        n_full = n // batch_size if batch_size % world_size == 0 else 0
        local = None
        if n_full:
            local = perm[:n_full * batch_size].view(n_full, batch_size // world_size, world_size)[:, :, rank].contiguous()
        tail = []
        for lo in range(n_full * batch_size, n, batch_size):
            g = perm[lo:lo + batch_size]
            tail.append((g[rank::world_size], int(g.numel())))
        return local, tail

    def flush(self):
        """Makes model.flat (the fp32 master weights) current on this rank.  Single GPU and the NCCL all-reduce schedule: no-op
        (every replica applies the whole update).  Sharded data-parallel update (dp_mode="peer"): COLLECTIVE -- every rank
        gathers the other ranks' master-weight shards; call it on all ranks before reading or saving the weights."""
        if self.dp_mode == "peer" and self._master_stale:
            self._gather_master()

    def evaluate(self, batch_idx, run=0):
        """Validation pass: corruption + forward + monitor sums only (train_dae_on_embedding.py:241-259),
        without building any autograd state.  Returns the reconstruction [B, io] (device, fp32).  Reads the weight buffer the
        GEMMs use, which is complete on every rank after every step (no master-weight gather needed)."""
        self._check_engine()
        B = int(batch_idx.numel())
        b = self._buffers(B)
        b["idx"].copy_(batch_idx, non_blocking=True)
        self._enqueue(B, b, run, B, self.data, b["idx"], self.corrupter.device_tables()[0], train=False)
        return b["acts"][-1][:, :self.io]

    def last_mask_ids(self, B):
        return self._bufs[B]["mask_id"]

    def reset_monitors(self):
        self.acc.zero_()
        if self.mixed is not None:
            self.mixed_acc.zero_()

    def read_monitors(self):
        """One D2H copy: {'full', 'partial', 'rows', 'last_loss'} (embedding) -- the quantities the reference
        accumulates on the host every step (train_dae_on_embedding.py:217-228)."""
        a = self.acc.cpu().tolist()
        out = dict(full=a[0], partial=a[1], rows=a[2], last_sum=a[3])
        if self.mixed is not None:
            V, K = len(self.mixed["arch"]), self.corrupter.k_max
            m = self.mixed_acc.cpu()
            out.update(ftl=float(m[0]), ptl=float(m[1]), ftl_per_k=m[2:2 + K * V].view(K, V).numpy().copy(),
                       ptl_per_k=m[2 + K * V:].view(K, V).numpy().copy(), loss=self.mixed_loss.cpu().tolist())
        return out

    def last_loss(self, B):
        """Loss of the most recent step (synchronises)."""
        if self.mixed is not None:
            return float(self.mixed_loss[0].item())
        return float(self.acc[3].item()) / (B * self.io)
