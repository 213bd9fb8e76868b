"""Multi-GPU parity check (run under torchrun, NCCL, one rank per GPU; not collected by pytest).

  * data-parallel invariance: G ranks x B/G samples with one all-reduce of the flat gradient buffer == one process on the
    whole global batch (loss, weights after 2 steps), eagerly and under CUDA-graph capture of the step (NCCL included);
  * sharded catalog: per-rank top-k + all-gather + codae_topk_merge == unsharded top-k (bit-exact indices and scores);
  * mask ids are a function of (seed, observation) only: every rank holds the same table.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

from codae.dataset import ConcatenatedEmbeddingDataset
from codae.model import EmbeddingDenoisingAutoencoder
from codae.tool import Corrupter, FusedStep
from codae.tool.inference import ComplementarityScorer, shard_rows


def build(dtype, ws, graph, overlap=True, deferred=False):
    torch.manual_seed(3)
    S, E, N = 3, 128, 1024
    cats = [torch.randn(N, E).abs() for _ in range(S)]
    ds = ConcatenatedEmbeddingDataset.from_tensors(cats)
    m = EmbeddingDenoisingAutoencoder(S * E, S * E, E, 2, 2, False)
    m.set_compute_dtype(dtype)
    m.to(dev)
    ds.to(dev)
    cor = Corrupter(N, ds.arch, 1, dev, seed=77)
    fs = FusedStep(m, cor, ds.data, lr=1e-3, weight_decay=1e-4, clip=True, world_size=ws, use_graph=graph, overlap_allreduce=overlap,
                   deferred_update=deferred)
    return ds, m, cor, fs


def flat(m):
    return m.flat.detach().clone()


ok = True
for dtype, graph, tol in [("fp32", False, 1e-5), ("bf16", False, 1e-2), ("bf16", True, 1e-2)]:
  for overlap in (True, False):
      GB = 64 * world
      rng = np.random.RandomState(5)
      batches = [rng.permutation(1024)[:GB] for _ in range(4)]
      batches.append(rng.permutation(1024)[:1])      # ragged last global batch: every rank but 0 holds no sample
      ds, m, cor, fs = build(dtype, world, graph, overlap)
      tbl32 = cor.device_tables()[0].to(torch.int32)          # NCCL has no int16
      tables = [torch.empty_like(tbl32) for _ in range(world)]
      dist.all_gather(tables, tbl32)
      same_table = all(torch.equal(t, tables[0]) for t in tables)
      for gidx in batches:
          local_idx = torch.as_tensor(gidx[rank::world], dtype=torch.int64, device=dev)
          fs.step(local_idx, global_batch=len(gidx))
      w_dp = flat(m)
      gathered = [torch.empty_like(w_dp) for _ in range(world)]
      dist.all_gather(gathered, w_dp)
      replicas_equal = all(torch.equal(g, gathered[0]) for g in gathered)       # every rank applied the same update
      # single-process reference on the whole global batch (same device, world_size=1)
      ds1, m1, cor1, fs1 = build(dtype, 1, False)
      for gidx in batches:
          fs1.step(torch.as_tensor(gidx, dtype=torch.int64, device=dev), global_batch=len(gidx))
      w_1 = flat(m1)
      err = float((w_dp - w_1).abs().max() / w_1.abs().max())
      good = same_table and replicas_equal and err < tol
      ok &= good
      if rank == 0:
          print("DP %s graph=%s overlap=%s: tables_equal=%s replicas_bitwise_equal=%s |w_dp - w_1|/|w| = %.2e (tol %.0e) -> %s"
                % (dtype, graph, overlap, same_table, replicas_equal, err, tol, "OK" if good else "FAIL"), flush=True)
      del fs, fs1

# deferred update under data parallelism (opt-in schedule): same weights as the immediate update up to the summation order of
# the norm (cooperative kernel there, codae_grad_sqnorm here: the clip scale may differ by an fp32 ulp)
if os.environ.get("CODAE_EXPERIMENTAL") == "1":
    for graph in (False, True):
        rng = np.random.RandomState(6)
        batches = [rng.permutation(1024)[:64 * world] for _ in range(5)]
        res = {}
        for deferred in (False, True):
            ds, m, cor, fs = build("bf16", world, graph, True, deferred)
            for gidx in batches:
                fs.step(torch.as_tensor(gidx[rank::world], dtype=torch.int64, device=dev), global_batch=len(gidx))
            fs.flush()
            torch.cuda.synchronize()
            res[deferred] = flat(m)
            del fs
        err = float((res[False] - res[True]).abs().max() / res[False].abs().max())
        good = err < 1e-4
        ok &= good
        if rank == 0:
            print("DP deferred update graph=%s: |w_deferred - w_immediate|/|w| = %.2e -> %s" % (graph, err, "OK" if good else "FAIL"), flush=True)

# sharded catalog
torch.manual_seed(9)
n, E, k = 200_003, 512, 10
g = torch.Generator(device=dev).manual_seed(9)
catalog = torch.rand((n, E), generator=g, device=dev)           # same seed on every rank -> same global catalog
q = torch.rand((3, E), generator=g, device=dev)
lo, c = shard_rows(n, world, rank)
s_sh, i_sh = ComplementarityScorer(catalog[lo:lo + c].contiguous(), E, "sqerr", k, row_offset=lo).topk(q)
s_full, i_full = ComplementarityScorer(catalog, E, "sqerr", k).topk_local(q)
good = torch.equal(i_sh, i_full) and torch.equal(s_sh, s_full)
ok &= good
if rank == 0:
    print("sharded top-k over %d ranks == unsharded: %s" % (world, "OK" if good else "FAIL"), flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST CHECK", "PASSED" if int(t.item()) == 1 else "FAILED", flush=True)
code = 0 if int(t.item()) == 1 else 1
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(code)      # skip NCCL/graph teardown order issues at interpreter exit
