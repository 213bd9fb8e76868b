"""CPU, world_size 2, gloo: the host-side data-parallel logic of the N>1 path.
 - every global batch is partitioned across ranks (ragged last batch included);
 - per-rank gradients computed with the GLOBAL batch size in the loss scale, summed by all-reduce, equal the
   single-process gradient (what FusedStep does with its flat gradient buffer over NCCL);
 - catalog row shards + all-gather of per-shard top-k + deterministic merge equal the unsharded top-k.
The kernels are stood in for by the oracle here; the same checks run through the CUDA path in the gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, PKG, os.path.join(PKG, "script")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import codae_oracle as O
    from oracle.philox import philox_mask_table
    from _common import epoch_batches
    from codae.tool.inference import shard_rows
    torch.manual_seed(0)
    S, E, N, Bg = 3, 8, 64, 10
    io = S * E
    dims, _ = O.layer_dims(io, io, 2, 2, False)
    W = [torch.randn(o, i) * 0.2 for i, o, _ in dims]
    b = [torch.zeros(o) for _, o, _ in dims]
    relu = [r for _, _, r in dims]
    data = torch.rand(N, io)
    arch = [dict(size=E, position=p) for p in range(0, io, E)]
    bm, nm, _ = O.binary_masks(arch, 1)
    tbl = torch.from_numpy(philox_mask_table(7, N, 3).astype(np.int64))     # same on every rank: f(seed, obs) only
    rng = np.random.RandomState(3)                                        # shared seed -> same global permutation
    seen = []
    for local, gb in epoch_batches(list(range(N)), Bg, rng, rank, world):
        seen += list(local)
        if len(local) == 0:
            continue
        x = data[list(local)]
        _, fmask = O.get_masks(bm, nm, tbl, list(local), 0, 1)
        y, acts = O.forward(W, b, relu, O.corrupt(x, fmask), keep=True)
        dy = (y - x) * (2.0 / (gb * io))                                   # GLOBAL batch in the scale
        gW, gb_ = O.backward(W, relu, acts, dy)
        flat = torch.cat([t.reshape(-1) for pair in zip(gW, gb_) for t in pair])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if rank == 0:
            # single-process reference on the whole global batch
            perm_rng = np.random.RandomState(3)
        last = (flat, gb)
    # recompute the last global batch on one process
    rng2 = np.random.RandomState(3)
    perm = rng2.permutation(N)
    gidx = list(np.asarray(range(N))[perm][-(N % Bg or Bg):])
    xg = data[gidx]
    _, fm = O.get_masks(bm, nm, tbl, gidx, 0, 1)
    yg, actsg = O.forward(W, b, relu, O.corrupt(xg, fm), keep=True)
    _, dyg = O.mse_mean_loss_and_grad(xg, yg)
    gWg, gbg = O.backward(W, relu, actsg, dyg)
    want = torch.cat([t.reshape(-1) for pair in zip(gWg, gbg) for t in pair])
    ok_grad = bool(torch.allclose(last[0], want, rtol=1e-5, atol=1e-8))
    all_seen = [None] * world
    dist.all_gather_object(all_seen, seen)
    ok_part = sorted(sum(all_seen, [])) == list(range(N))
    # catalog sharding + merge
    n, k = 1003, 7
    cat = torch.rand(n, E)
    q = torch.rand(E)
    lo, c = shard_rows(n, world, rank)
    s, i = O.topk(O.score_candidates(cat[lo:lo + c], q), k, row_offset=lo)
    s = s.float()
    gs = [torch.empty(k) for _ in range(world)]
    gi = [torch.empty(k, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gs, s)
    dist.all_gather(gi, i)
    ms, mi = O.topk_merge(gs, gi, k)
    ws, wi = O.topk(O.score_candidates(cat, q), k)
    ok_topk = mi.tolist() == wi.tolist()
    with open(os.path.join(out_dir, "r%d.txt" % rank), "w") as f:
        f.write("%d %d %d" % (ok_grad, ok_part, ok_topk))
    dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "r%d.txt" % r)).read() == "1 1 1"
