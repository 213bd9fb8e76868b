"""GPU parity tests of complementarity inference: top-k rankings bit-exact against the fp64 oracle, rank mode against
the reference's RankingLoss golden values, shard/merge invariance, and size-independent properties at scale."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


@pytest.mark.parametrize("n,E,Q,k,metric,bf16", [(5000, 512, 1, 10, "sqerr", False), (5000, 512, 1, 10, "cosine", False),
                                                 (20000, 512, 5, 100, "sqerr", False), (3000, 64, 4, 7, "cosine", False),
                                                 (7, 512, 1, 10, "sqerr", False), (40000, 512, 2, 10, "sqerr", True),
                                                 (1, 16, 1, 1, "cosine", False), (9000, 4096, 1, 128, "sqerr", False)])
def test_topk_matches_oracle(n, E, Q, k, metric, bf16):
    from oracle import codae_oracle as O
    from codae.tool.inference import ComplementarityScorer
    torch.manual_seed(n + E + k)
    cat = torch.randn(n, E).abs() * (torch.rand(n, E) < 0.7)
    if bf16:
        cat = cat.to(torch.bfloat16)
    q = torch.randn(Q, E).abs() * 0.3
    if n > 10:
        cat[n // 2] = cat[3]                      # exact duplicate rows: ties must resolve to the lower index
        q[0] = cat[3].float() * 0.5 + 0.01
    sc = ComplementarityScorer(cat.to(DEV).contiguous(), E, metric=metric, k=k, inv_scale=0.5, row_offset=1000)
    s, i = sc.topk_local(q.to(DEV))
    s, i = s.cpu(), i.cpu()
    for qi in range(Q):
        scores = O.score_candidates(cat.float(), q[qi], metric, inv_scale=0.5)
        ws, wi = O.topk(scores, k, metric, row_offset=1000)
        m = min(k, n)
        assert i[qi, :m].tolist() == wi.tolist(), (qi, i[qi, :m], wi)
        assert torch.all(i[qi, m:] == -1)
        assert np.allclose(s[qi, :m].numpy(), ws.numpy(), rtol=2e-5, atol=1e-6)


def test_topk_shard_merge_equals_unsharded():
    from oracle import codae_oracle as O
    from codae import _C
    from codae.tool.inference import ComplementarityScorer, shard_rows
    torch.manual_seed(0)
    n, E, k, G = 30011, 512, 10, 4
    cat = torch.randn(n, E).abs().to(DEV)
    q = torch.randn(3, E).abs().to(DEV)
    full_s, full_i = ComplementarityScorer(cat, E, "sqerr", k).topk_local(q)
    full_s, full_i = full_s.clone(), full_i.clone()
    parts_s, parts_i = [], []
    for r in range(G):
        lo, c = shard_rows(n, G, r)
        s, i = ComplementarityScorer(cat[lo:lo + c], E, "sqerr", k, row_offset=lo).topk_local(q)
        parts_s.append(s.clone()); parts_i.append(i.clone())
    ms, mi = torch.empty_like(full_s), torch.empty_like(full_i)
    _C.topk_merge(torch.stack(parts_s), torch.stack(parts_i), _C.METRIC_SQERR, ms, mi)
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)
    # and the same merge restated on the host
    for qi in range(3):
        _, wi = O.topk_merge([p[qi].cpu() for p in parts_s], [p[qi].cpu() for p in parts_i], k)
        assert wi.tolist() == mi[qi].cpu().tolist()


@pytest.mark.parametrize("name", ["emb_small"])
def test_ranking_loss_matches_reference(name):
    """RankingLoss.get (metering.py:46-79) on the golden validation batch: same ranks, same loss."""
    from oracle import codae_oracle as O
    from codae.dataset import ConcatenatedEmbeddingDataset
    from codae.tool import RankingLoss
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    e = int(g["e"])
    ncat = int(g["io"]) // e
    ds = ConcatenatedEmbeddingDataset.from_tensors([torch.from_numpy(g["cat%d" % c]) for c in range(ncat)])
    ds.to(DEV)
    val = [int(v) for v in g["rank_val"]]
    rl = RankingLoss(ds, val, device=DEV)
    pred = torch.from_numpy(g["rank_pred"]).to(DEV)
    fmask = torch.from_numpy(g["rank_fmask"]).to(DEV)
    idx = tuple(int(i) for i in g["rank_idx"])
    total, ranks = O.ranking_loss([torch.from_numpy(g["cat%d" % c]) for c in range(ncat)], e, val,
                                  torch.from_numpy(g["rank_pred"]), torch.from_numpy(g["rank_fmask"]), idx)
    assert rl.ranks(pred, fmask, idx).cpu().tolist() == ranks
    assert abs(rl.get(pred, fmask, idx) - float(g["rank_loss"])) < 1e-9


def test_scale_properties_planted_winners():
    """2M x 512 fp32 (4 GB) catalog, too big for the oracle in seconds: plant k near-copies of the query at known
    rows -> the top-k must be exactly those rows in order of their planted distance; bf16 catalog agrees."""
    from codae.tool.inference import ComplementarityScorer
    torch.manual_seed(5)
    n, E, k = 2_000_000, 512, 10
    g = torch.Generator(device=DEV).manual_seed(5)
    cat = torch.rand((n, E), generator=g, device=DEV)
    q = torch.rand((1, E), generator=g, device=DEV)
    rows = torch.tensor([1_999_999, 7, 1_000_003, 55_555, 1_234_567, 0, 999_999, 424_242, 1_500_000, 31], device=DEV)
    for j, r in enumerate(rows.tolist()):
        cat[r] = q[0] + 1e-3 * (j + 1)           # squared error grows with j
    s, i = ComplementarityScorer(cat, E, "sqerr", k).topk_local(q)
    assert i[0].tolist() == rows.tolist()
    assert torch.all(s[0, 1:] > s[0, :-1])
    sb, ib = ComplementarityScorer(cat.to(torch.bfloat16), E, "sqerr", k).topk_local(q)
    assert sorted(ib[0].tolist()) == sorted(rows.tolist())
    # permutation invariance of the selected set: reversed catalog
    s2, i2 = ComplementarityScorer(cat.flip(0).contiguous(), E, "sqerr", k).topk_local(q)
    assert (n - 1 - i2[0]).tolist() == rows.tolist()


@pytest.mark.parametrize("bf16", [False, True])
def test_full_size_catalog_properties(bf16):
    """BASELINE's inference size: 10 M x 512 (20.5 GB fp32 / 10.2 GB bf16).  Size-independent properties: planted
    near-copies of the query are returned in order of their planted distance; the sharded sweep (8 contiguous shards +
    merge) returns the same list as the single sweep; a second sweep returns the same bits."""
    from codae import _C
    from codae.tool.inference import ComplementarityScorer, shard_rows
    n, E, k = 10_000_000, 512, 10
    g = torch.Generator(device=DEV).manual_seed(11)
    cat = torch.rand((n, E), generator=g, device=DEV)
    if bf16:
        cat = cat.to(torch.bfloat16)
    q = torch.rand((1, E), generator=g, device=DEV)
    rows = torch.tensor([9_999_999, 3, 5_000_001, 7_777_777, 1_250_000, 0, 2_500_000, 8_750_001, 6_000_000, 42], device=DEV)
    for j, r in enumerate(rows.tolist()):
        cat[r] = (q[0] + 2e-2 * (j + 1)).to(cat.dtype)       # squared error grows with j (well above bf16 rounding)
    sc = ComplementarityScorer(cat, E, "sqerr", k)
    s, i = sc.topk_local(q)
    s, i = s.clone(), i.clone()
    assert i[0].tolist() == rows.tolist() and torch.all(s[0, 1:] > s[0, :-1])
    s2, i2 = sc.topk_local(q)
    assert torch.equal(s2, s) and torch.equal(i2, i)
    ps, pi = [], []
    for r in range(8):
        lo, c = shard_rows(n, 8, r)
        a, b = ComplementarityScorer(cat[lo:lo + c], E, "sqerr", k, row_offset=lo).topk_local(q)
        ps.append(a.clone()); pi.append(b.clone())
    ms, mi = torch.empty_like(s), torch.empty_like(i)
    _C.topk_merge(torch.stack(ps), torch.stack(pi), _C.METRIC_SQERR, ms, mi)
    assert torch.equal(mi, i) and torch.equal(ms, s)


def _swap_model(S, E, z, nin, nout, dtype, seed):
    from codae.model import EmbeddingDenoisingAutoencoder
    torch.manual_seed(seed)
    model = EmbeddingDenoisingAutoencoder(S * E, z, E, nin, nout, False)
    for lin in model.linears():
        torch.nn.init.uniform_(lin.bias, -0.05, 0.05)
    W = [lin.weight.detach().clone() for lin in model.linears()]
    b = [lin.bias.detach().clone() for lin in model.linears()]
    model = model.to(DEV).set_compute_dtype(dtype)
    return model, W, b


@pytest.mark.parametrize("n,S,E,slot,k,chunk,cat_bf16", [(3000, 4, 64, 2, 10, 1024, False), (777, 3, 128, 0, 5, 8192, False),
                                                         (5, 4, 64, 3, 10, 256, False), (2500, 4, 64, 1, 16, 512, True)])
def test_swap_scores_match_oracle_fp32(n, S, E, slot, k, chunk, cat_bf16):
    """Full-reconstruction swap scoring on the exact-fp32 engine: the ranking equals the oracle's, errors to 1e-5."""
    from oracle import codae_oracle as O
    from codae.tool.inference import SwapScorer
    model, W, b = _swap_model(S, E, 96, 2, 2, "fp32", n + slot)
    cat = torch.rand(n, E) * 2
    if cat_bf16:
        cat = cat.to(torch.bfloat16)
    if n > 10:
        cat[n - 1] = cat[4]                         # duplicate candidate: tie -> lower index first
    outfit = torch.rand(S * E)
    sc = SwapScorer(model, cat.to(DEV), E, k=k, inv_scale=0.5, row_offset=100, chunk=chunk)
    s, i = sc.topk_local(outfit.to(DEV), slot)
    s, i = s.cpu(), i.cpu()
    err = O.score_swaps(W, b, model.relu, outfit, slot, E, cat.float(), inv_scale=0.5)
    ws, wi = O.topk(err, k, "sqerr", row_offset=100)
    m = min(k, n)
    assert i[:m].tolist() == wi.tolist(), (i, wi)
    assert torch.all(i[m:] == -1)
    assert np.allclose(s[:m].numpy(), ws.numpy(), rtol=1e-5, atol=1e-6)
    if n > 10 and (n - 1 + 100) in i.tolist():
        pos = i.tolist()
        assert pos.index(104) + 1 == pos.index(n - 1 + 100)


def test_swap_scores_bf16_engine_and_chunk_invariance():
    """Tensor-core engine: errors within the bf16 tolerance (1e-2 relative) of the fp32 oracle at the returned indices,
    the returned set is the oracle's top-k up to near-ties, and the ranking does not depend on the chunk size."""
    from oracle import codae_oracle as O
    from codae.tool.inference import SwapScorer
    n, S, E, slot, k = 6000, 4, 128, 1, 10
    model, W, b = _swap_model(S, E, 256, 2, 2, "bf16", 3)
    from codae import _C
    assert model.engine_dtype() == _C.BF16
    cat = torch.rand(n, E)
    outfit = torch.rand(S * E)
    res = []
    for chunk in (1024, 4096, 8192):
        s, i = SwapScorer(model, cat.to(DEV), E, k=k, chunk=chunk).topk_local(outfit.to(DEV), slot)
        res.append((s.cpu(), i.cpu()))
    for s, i in res[1:]:
        # the chunk size changes the GEMM schedule (split-K factor -> fp32 summation order), not the ranking
        assert torch.equal(i, res[0][1]) and np.allclose(s.numpy(), res[0][0].numpy(), rtol=1e-4)
    s, i = res[0]
    err = O.score_swaps(W, b, model.relu, outfit, slot, E, cat)
    assert np.allclose(s.numpy(), err[i].numpy(), rtol=1e-2)
    kth = float(O.topk(err, k)[0][-1])
    assert float(err[i].max()) <= kth * 1.01


def test_ranking_loss_gemm_path_matches_row_sweep_and_fp64():
    """Q >= 64 queries of one category: the dot products of RankingLoss come from ONE contraction on the fp32-parity tensor-core
    engine (codae_linear_fwd_x3 + codae_rank_count) instead of one warp sweep per query (metering.py:46-79 runs for every
    validation batch, train_dae_on_embedding.py:241-261).  Ranks are counts of strict fp32 comparisons: they must equal an fp64
    ranking wherever no pair of cosines is within fp32 noise, and agree with the row-sweep kernel."""
    from codae.dataset import ConcatenatedEmbeddingDataset
    from codae.tool import RankingLoss
    torch.manual_seed(21)
    N, E, S, B = 6000, 128, 3, 384
    cats = [torch.randn(N, E) for _ in range(S)]
    ds = ConcatenatedEmbeddingDataset.from_tensors(cats)
    ds.to(DEV)
    val = list(range(1000, 5000))
    idx = torch.randperm(4000)[:B] + 1000                       # true items inside the validation subset
    idx[::7] = torch.randint(0, 1000, (len(idx[::7]),))         # ... and some outside it
    cat_of = torch.arange(B) % S
    cat_of[:6] = 0                                              # category 0: 132 queries (>= 128: GEMM path), the others 126 (row sweep)
    pred = torch.randn(B, S * E)
    for i in range(B):                                          # prediction near the true item so ranks are spread, not all ~n/2
        c = int(cat_of[i])
        pred[i, c * E:(c + 1) * E] = ds.data_per_category[c][idx[i]].cpu() * (0.3 * torch.rand(1)) + torch.randn(E)
    fmask = torch.ones(B, S * E)
    for i in range(B):
        fmask[i, int(cat_of[i]) * E:(int(cat_of[i]) + 1) * E] = 0
    pred_d, fmask_d = pred.to(DEV), fmask.to(DEV)
    rl_gemm = RankingLoss(ds, val, device=DEV)
    rl_gemm.GEMM_MIN_Q = 128
    rl_sweep = RankingLoss(ds, val, device=DEV)
    rl_sweep.GEMM_MIN_Q = 1 << 30
    r_gemm = rl_gemm.ranks(pred_d, fmask_d, idx).cpu()
    r_sweep = rl_sweep.ranks(pred_d, fmask_d, idx).cpu()
    assert 0 in rl_gemm._gemm and 1 not in rl_gemm._gemm        # category 0 went through the contraction, the others did not
    # fp64 ranking with a margin: pairs closer than 1e-6 may fall either way in fp32
    lo, hi = torch.zeros(B, dtype=torch.int64), torch.zeros(B, dtype=torch.int64)
    v = torch.tensor(val)
    for i in range(B):
        c = int(cat_of[i])
        cat = ds.data_per_category[c].cpu().double()
        qv = pred[i, c * E:(c + 1) * E].double()
        cos = (cat @ qv) / (cat.norm(dim=1) * qv.norm()).clamp_min(1e-8)
        st = cos[idx[i]]
        lo[i] = int((st > cos[v] + 1e-6).sum())
        hi[i] = int((st > cos[v] - 1e-6).sum())
    assert bool(((r_gemm >= lo) & (r_gemm <= hi)).all()) and bool(((r_sweep >= lo) & (r_sweep <= hi)).all())
    assert float((r_gemm == r_sweep).float().mean()) > 0.98 and int((r_gemm - r_sweep).abs().max()) <= 2
    assert r_gemm.float().std() > 100                           # the ranks are spread out: the test is not vacuous
