"""GPU parity tests of the whole training step against the golden vectors minted from the reference
(tests/golden, oracle/gen_golden.py) and against the oracle at the shipped configs' full layer sizes.
fp32 engine: loss / reconstructions / gradients / post-Adam weights within 1e-5 relative;
bf16 tensor-core engine: within 1e-2 (BASELINE.json north_star tolerances)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def load_params(model, flat, shapes):
    o = 0
    for p, sh in zip(model.parameters(), shapes):
        sh = [int(s) for s in sh if s > 0]
        n = int(np.prod(sh))
        p.data.copy_(torch.from_numpy(flat[o:o + n].copy()).reshape(sh))
        o += n


def flat_params(model):
    return np.concatenate([p.detach().float().cpu().numpy().ravel() for p in model.parameters()])


def flat_grads(model):
    return np.concatenate([p.grad.detach().float().cpu().numpy().ravel() for p in model.parameters()])


def build_embedding(g, dtype="fp32"):
    from codae.dataset import ConcatenatedEmbeddingDataset
    from codae.model import EmbeddingDenoisingAutoencoder
    from codae.tool import Corrupter
    io, e = int(g["io"]), int(g["e"])
    ncat = io // e
    ds = ConcatenatedEmbeddingDataset.from_tensors([torch.from_numpy(g["cat%d" % c]) for c in range(ncat)])
    assert np.array_equal(ds.data.numpy(), g["data"])          # same scaling rule as the reference dataset
    model = EmbeddingDenoisingAutoencoder(io, int(g["z"]), e, int(g["nin"]), int(g["nout"]), False)
    load_params(model, g["init"], g["shapes"])
    model.set_compute_dtype(dtype)
    model.to(DEV)
    ds.to(DEV)
    cor = Corrupter(ds.nb_observation, ds.arch, int(g["k_max"]), DEV)
    cor.mask_to_use = torch.from_numpy(g["mask_to_use"])        # inject the reference's table (plain attribute)
    return ds, model, cor


# "fp32" = the reference's precision on the tensor cores (bf16 triples, CODAE_F32X3) wherever every layer is at least 32 wide
# (emb_small, emb_k2, emb_mid), the FFMA engine otherwise (emb_bottleneck); "fp32_simt" forces the FFMA engine.
@pytest.mark.parametrize("dtype", ["fp32", "fp32_simt"])
@pytest.mark.parametrize("name", ["emb_small", "emb_bottleneck", "emb_k2", "emb_mid"])
def test_legacy_api_matches_reference(name, dtype):
    """The reference's own call pattern: get_masks -> corrupt -> model() -> MSELoss -> backward -> clip -> Adam.step."""
    from codae import _C
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    ds, model, cor = build_embedding(g, dtype)
    assert model.engine_dtype() == (_C.F32X3 if dtype == "fp32" and name != "emb_bottleneck" else _C.F32)
    opt = torch.optim.Adam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["wd"]))
    crit = torch.nn.MSELoss(reduction="mean")
    s = 0
    while "idx%d" % s in g:
        bi = tuple(int(i) for i in g["idx%d" % s])
        x = torch.stack([ds[i][0] for i in bi])
        masks, fmask = cor.get_masks(bi, 0)
        assert np.array_equal(fmask.cpu().numpy(), g["fmask%d" % s])
        cx = model.corrupt(input_data=x, mask=fmask)
        assert np.array_equal(cx.cpu().numpy(), g["cx%d" % s])
        y = model(cx)
        loss = crit(x, y)
        opt.zero_grad()
        loss.backward()
        assert rel(y.detach().cpu().numpy(), g["y%d" % s]) < 1e-5
        assert abs(float(loss) - float(g["loss%d" % s])) <= 1e-5 * float(g["loss%d" % s])
        assert rel(flat_grads(model), g["grads%d" % s]) < 1e-5
        if bool(g["clip"]):
            gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
            assert abs(float(gn) - float(g["gnorm%d" % s])) <= 1e-5 * float(g["gnorm%d" % s])
        opt.step()
        assert rel(flat_params(model), g["post%d" % s]) < 1e-5
        s += 1
    assert s >= 2


@pytest.mark.parametrize("dtype", ["fp32", "fp32_simt"])
@pytest.mark.parametrize("name,graph", [("emb_small", False), ("emb_bottleneck", False), ("emb_k2", False), ("emb_mid", False),
                                        ("emb_mid", True)])
def test_fused_step_matches_reference(name, graph, dtype):
    """FusedStep (the scripts' fast path): same observables, no autograd, flat-buffer clip + Adam kernels."""
    from codae.tool import FusedStep
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    ds, model, cor = build_embedding(g, dtype)
    fs = FusedStep(model, cor, ds.data, lr=float(g["lr"]), weight_decay=float(g["wd"]), clip=bool(g["clip"]), use_graph=graph)
    B = int(g["B"])
    s = 0
    prev = g["init"]
    while "idx%d" % s in g:
        idx = torch.from_numpy(g["idx%d" % s]).to(DEV)
        fs.step(idx, run=0)
        assert abs(fs.last_loss(B) - float(g["loss%d" % s])) <= 1e-5 * float(g["loss%d" % s])
        post = flat_params(model)
        assert rel(post, g["post%d" % s]) < 1e-5
        assert rel(post - prev, g["post%d" % s] - prev) < 5e-3       # the update itself
        if not graph:
            assert rel(flat_grads(model), g["grads%d" % s]) < 1e-5    # .grad are views of the flat gradient buffer
            y = fs._bufs[B]["acts"][-1][:, :int(g["io"])]
            assert rel(y.cpu().numpy(), g["y%d" % s]) < 1e-5
            assert torch.equal(fs.last_mask_ids(B).cpu().long(), torch.from_numpy(g["mask_to_use"])[idx.cpu(), 0])
        prev = g["post%d" % s]
        s += 1
    mon = fs.read_monitors()
    ftl = sum(float(g["ftl%d" % i]) for i in range(s))
    ptl = sum(float(g["ptl%d" % i]) for i in range(s))
    assert abs(mon["full"] - ftl) <= 1e-5 * ftl and abs(mon["partial"] - ptl) <= 1e-5 * ptl and mon["rows"] == s * B


def test_fused_step_bf16_within_1e2():
    """bf16 tensor-core engine on the golden run: loss within 1e-2 relative, weights track the fp32 reference."""
    from codae import _C
    from codae.tool import FusedStep
    g = np.load(os.path.join(GOLDEN, "emb_mid.npz"))
    ds, model, cor = build_embedding(g, dtype="bf16")
    assert model.engine_dtype() == _C.BF16
    fs = FusedStep(model, cor, ds.data, lr=float(g["lr"]), weight_decay=float(g["wd"]), clip=True)
    B = int(g["B"])
    for s in range(2):
        fs.step(torch.from_numpy(g["idx%d" % s]).to(DEV), run=0)
        assert abs(fs.last_loss(B) - float(g["loss%d" % s])) <= 1e-2 * float(g["loss%d" % s])
        y = fs._bufs[B]["acts"][-1][:, :int(g["io"])]
        assert rel(y.cpu().numpy(), g["y%d" % s]) < 1e-2
    assert rel(flat_params(model), g["post1"]) < 1e-2
    assert torch.equal(model.flat_bf16.view(torch.int16), model.flat.to(torch.bfloat16).view(torch.int16))


@pytest.mark.parametrize("name", ["abalone_k1", "abalone_k3"])
def test_abalone_fused_and_legacy(name):
    from oracle.gen_golden import abalone_arch
    from codae.dataset import MixedVariableDataset
    from codae.model import MixedVariableDenoisingAutoencoder
    from codae.tool import CombinedCriterion, Corrupter, FusedStep
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    arch = abalone_arch()
    k_max, B = int(g["k_max"]), int(g["B"])

    def build():
        ds = MixedVariableDataset.from_arch(arch, torch.from_numpy(g["data"]))
        m = MixedVariableDenoisingAutoencoder(arch, 11, int(g["z"]), DEV, 2, 2, bool(g["steep"]))
        load_params(m, g["init"], g["shapes"])
        m.to(DEV)
        ds.to(DEV)
        cor = Corrupter(ds.nb_observation, arch, k_max, DEV)
        cor.mask_to_use = torch.from_numpy(g["mask_to_use"])
        return ds, m, cor

    # fused path
    ds, model, cor = build()
    fs = FusedStep(model, cor, ds.data, lr=float(g["lr"]), weight_decay=float(g["wd"]), clip=True,
                   mixed=dict(arch=arch, weight=list(g["weight"]), norm_scale=torch.from_numpy(g["norm_scale"]),
                              norm_min=torch.from_numpy(g["norm_min"]), norm_first=3))
    tot = dict(ftl=0.0, ptl=0.0, fk=0.0, pk=0.0)
    for s in range(3):
        fs.step(torch.from_numpy(g["idx%d" % s]).to(DEV), run=int(g["run%d" % s]))
        assert abs(fs.last_loss(B) - float(g["loss%d" % s])) <= 1e-5 * float(g["loss%d" % s])
        assert rel(flat_grads(model), g["grads%d" % s]) < 1e-5
        assert rel(flat_params(model), g["post%d" % s]) < 1e-5
        tot["ftl"] += g["mon%d" % s].sum(); tot["ptl"] += g["mon_partial%d" % s].sum()
        tot["fk"] = tot["fk"] + g["mon_per_k%d" % s]; tot["pk"] = tot["pk"] + g["mon_partial_per_k%d" % s]
    mon = fs.read_monitors()
    assert abs(mon["ftl"] - tot["ftl"]) <= 1e-5 * tot["ftl"] and abs(mon["ptl"] - tot["ptl"]) <= 1e-5 * tot["ptl"]
    assert rel(mon["ftl_per_k"], tot["fk"]) < 1e-5 and rel(mon["ptl_per_k"], tot["pk"]) < 1e-5

    # legacy path: CombinedCriterion + autograd + torch Adam, as train_dae_on_abalone.py:206-236 calls them
    ds, model, cor = build()
    opt = torch.optim.Adam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["wd"]))
    crit = CombinedCriterion(arch=arch, k_max=k_max, device=DEV, observation_mask=torch.tensor([0, 0, 0] + [1] * 8),
                             weight=list(g["weight"]), reduction="mean")
    mon_c = CombinedCriterion(arch=arch, k_max=k_max, device=DEV, observation_mask=torch.tensor([0, 0, 0] + [1] * 8), reduction="none")
    for s in range(3):
        bi = tuple(int(i) for i in g["idx%d" % s])
        x = torch.stack([ds[i][0] for i in bi])
        masks, fmask = cor.get_masks(bi, int(g["run%d" % s]))
        y = model(model.corrupt(input_data=x, mask=fmask))
        loss = crit(x=x, y=y)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        opt.step()
        assert abs(float(loss) - float(g["loss%d" % s])) <= 1e-5 * float(g["loss%d" % s])
        assert rel(flat_params(model), g["post%d" % s]) < 1e-5
        xd, yd = x.clone(), y.detach().clone()
        sc, mn = torch.from_numpy(g["norm_scale"]).to(DEV), torch.from_numpy(g["norm_min"]).to(DEV)
        xd[:, 3:] = xd[:, 3:] * sc + mn
        yd[:, 3:] = yd[:, 3:] * sc + mn
        ml = mon_c(xd, yd, as_numpy=True)
        assert rel(ml, g["mon%d" % s]) < 1e-5
        assert rel(mon_c.get_per_k(ml, masks), g["mon_per_k%d" % s]) < 1e-5


@pytest.mark.parametrize("cfg,B,dtype,tol,graph", [("embedding", 128, "fp32", 1e-5, False), ("modanet", 32, "fp32", 1e-5, False),
                                                   ("embedding", 128, "fp32", 1e-5, True), ("embedding", 128, "fp32_simt", 1e-5, False),
                                                   ("modanet", 32, "bf16", 1e-2, False), ("bottleneck", 64, "fp32", 1e-5, False),
                                                   ("embedding", 128, "bf16", 1e-2, False), ("embedding", 128, "bf16", 1e-2, True),
                                                   ("bottleneck", 64, "bf16", 1e-2, True)])
def test_full_size_step_vs_oracle(cfg, B, dtype, tol, graph):
    """BASELINE configs at their real layer sizes (10 x 1536^2 / 8 x 1536^2 / the 1067-598-... bottleneck):
    fused steps vs the oracle on the same seeded inputs.  ("embedding", 128, "bf16", graph=True) is exactly what bench.py's
    embedding.yaml block runs: tcgen05 engine, CUDA-graph replay, sum(dW^2) from the weight-gradient epilogue (the first step
    of a shape runs eagerly, the second is captured and replayed, the third is a pure replay)."""
    from oracle import codae_oracle as O
    from oracle.philox import philox_mask_table
    from codae.dataset import ConcatenatedEmbeddingDataset
    from codae.model import EmbeddingDenoisingAutoencoder
    from codae.tool import Corrupter, FusedStep
    torch.manual_seed(11)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    z, nin, nout, lr, wd, clip = {"embedding": (1536, 4, 4, 1e-5, 1e-4, True), "modanet": (1536, 3, 3, 1e-4, 1e-2, False),
                                  "bottleneck": (128, 3, 3, 1e-4, 1e-4, True)}[cfg]
    N, E, S = 512, 512, 3
    cats = [torch.randn(N, E).abs() * (torch.rand(N, E) < 0.7) for _ in range(S)]
    ds = ConcatenatedEmbeddingDataset.from_tensors(cats)
    model = EmbeddingDenoisingAutoencoder(S * E, z, E, nin, nout, False)
    W = [l.weight.detach().clone() for l in model.linears()]
    b = [l.bias.detach().clone() for l in model.linears()]
    data_cpu = ds.data.clone()
    model.set_compute_dtype(dtype)
    model.to(DEV)
    ds.to(DEV)
    cor = Corrupter(N, ds.arch, 1, DEV, seed=2024)
    tbl = torch.from_numpy(philox_mask_table(2024, N, 3).astype(np.int64))
    assert torch.equal(tbl, cor.mask_to_use)
    fs = FusedStep(model, cor, ds.data, lr=lr, weight_decay=wd, clip=clip, use_graph=graph)
    if dtype == "bf16":
        assert fs.eng == 1 and (fs.wgrad_sqnorm or not clip)            # tensor-core engine, norm-free clipped step
    if dtype == "fp32":
        # the reference's precision on the tensor cores: bf16 triples (CODAE_F32X3), same norm-free clipped step
        assert fs.eng == 2 and (fs.wgrad_sqnorm or not clip)
    if dtype == "fp32_simt":
        assert fs.eng == 0
    dae = O.OracleDAE(W, b, model.relu, lr, wd, clip)
    bm, nm, _ = O.binary_masks(ds.arch, 1)
    perm = torch.randperm(N)
    init = np.concatenate([t.numpy().ravel() for pair in zip(W, b) for t in pair])
    for s in range(3 if graph else 2):
        idx = perm[s * B:(s + 1) * B]
        before = flat_params(model)
        fs.step(idx.to(DEV), run=0)
        _, fmask = O.get_masks(bm, nm, tbl, idx.tolist(), 0, 1)
        r = dae.step_embedding(data_cpu[idx], fmask)
        assert abs(fs.last_loss(B) - r["loss"]) <= tol * abs(r["loss"]), (s, fs.last_loss(B), r["loss"])
        y = fs._bufs[B]["acts"][-1][:, :S * E]
        assert rel(y.cpu().numpy(), r["y"].numpy()) < tol, s
        if dtype.startswith("fp32"):
            # gradients.  ReLU's derivative is discontinuous: a unit whose pre-activation is within fp32 summation error
            # (~4e-8 here) of 0 passes gradient on one side and not on the other, while its forward value is ~0 either
            # way (loss and reconstruction still agree to 1e-5).  With 128 x 1536 x 8 ReLU evaluations ~2 such flips are
            # expected per step; each removes one of the ~2e5 (sample, unit) rank-1 terms of the gradient, i.e. an L2
            # error of ~1/sqrt(2e5) = 2e-3 (observed: 1e-3..2.5e-3 per layer).  So at full size the gradient is checked
            # in L2 (1e-2) and entry-wise for 98 % of entries; the strict 1e-5 entry-wise gradient checks live in the
            # golden-vector tests and in the full-size GEMM kernel tests above.
            gg, gw = flat_grads(model), np.concatenate([t.numpy().ravel() for t in r["grads"]])
            err = np.abs(gg - gw) / np.abs(gw).max()
            # per-layer diagnostics (weight, bias interleaved) in the failure message
            sizes = [p.numel() for p in model.parameters()]
            offs = np.cumsum([0] + sizes)
            stats = [(j, float(err[offs[j]:offs[j + 1]].max()),
                      float(np.linalg.norm(gg[offs[j]:offs[j + 1]] - gw[offs[j]:offs[j + 1]]) /
                            max(np.linalg.norm(gw[offs[j]:offs[j + 1]]), 1e-30))) for j in range(len(sizes))]
            l2 = float(np.linalg.norm(gg - gw) / np.linalg.norm(gw))
            assert l2 < 1e-2 and err.max() < 5e-2 and (err < tol).mean() > 0.98, (s, l2, float(err.max()), int((err >= tol).sum()), stats)
        # post-Adam weights.  BASELINE's gate is on loss and reconstructions (asserted above).  Adam's first steps are
        # g/(|g|+eps)-shaped: an entry whose gradient changed (ReLU flips above) or nearly cancels moves by up to
        # 2*lr/(1-b1) regardless of how small the change in g was.  So: 98 % of entries within tol of max|w|, and no
        # entry further than the largest possible Adam step.
        got = flat_params(model)
        want = np.concatenate([t.numpy().ravel() for t in dae.params()])
        dw = np.abs(got - want)
        assert (dw <= max(tol, 1e-5) * np.abs(want).max()).mean() > 0.98, s
        assert dw.max() <= 2.2 * lr / (1 - 0.9) + tol * np.abs(want).max(), (s, float(dw.max()))
        # per-step parity: restart the oracle from the device state so that step s+1 compares like with like
        # (otherwise the ill-conditioned elements above feed a chaotic 1e-5-level drift into the next reconstruction)
        off = 0
        with torch.no_grad():
            for l in range(len(model.dims)):
                dae.W[l].copy_(model.weight_view(model.flat, l).cpu())
                dae.b[l].copy_(model.bias_view(model.flat, l).cpu())
            for j, (mt, vt) in enumerate(zip(dae.m, dae.v)):
                l, is_bias = j // 2, j % 2
                view = model.bias_view if is_bias else model.weight_view
                mt.copy_(view(fs.m, l).cpu())
                vt.copy_(view(fs.v, l).cpu())


def test_pdl_and_splitk_switches_do_not_change_results():
    """Programmatic dependent launch and cluster split-K are scheduling choices: with either switched off the captured
    step gives the same loss (split-K changes fp32 summation order only; PDL and the early weight-tile requests must be
    bit-identical)."""
    from codae import _C
    from codae.tool import FusedStep
    g = np.load(os.path.join(GOLDEN, "emb_mid.npz"))
    results = {}
    for name, (pdl, splitk, prefetch) in {"on": (1, 1, 1), "no_pdl": (0, 1, 1), "no_splitk": (1, 0, 1),
                                          "no_prefetch": (1, 1, 0)}.items():
        _C.set_option(DEV, _C.OPT_PDL, pdl)
        _C.set_option(DEV, _C.OPT_SPLITK, splitk)
        _C.set_option(DEV, _C.OPT_WEIGHT_PREFETCH, prefetch)
        try:
            ds, model, cor = build_embedding(g, dtype="bf16")
            fs = FusedStep(model, cor, ds.data, lr=float(g["lr"]), weight_decay=float(g["wd"]), clip=True, use_graph=True)
            for rep in range(3):
                for s in range(2):
                    fs.step(torch.from_numpy(g["idx%d" % s]).to(DEV), run=0)
            results[name] = (fs.last_loss(int(g["B"])), flat_params(model))
        finally:
            _C.set_option(DEV, _C.OPT_PDL, 1)
            _C.set_option(DEV, _C.OPT_SPLITK, 1)
            _C.set_option(DEV, _C.OPT_WEIGHT_PREFETCH, 1)
    assert results["on"][0] == results["no_pdl"][0] and np.array_equal(results["on"][1], results["no_pdl"][1])
    # weight tiles requested ahead of the stream dependency must never see weights of the previous step
    assert results["on"][0] == results["no_prefetch"][0] and np.array_equal(results["on"][1], results["no_prefetch"][1])
    assert abs(results["on"][0] - results["no_splitk"][0]) <= 1e-3 * abs(results["on"][0])


def test_polyvore_shape_step_bf16_vs_oracle():
    """The largest config's layer shape (4096 wide, k_max=2, bf16) at a batch that takes the persistent, TMEM-double-
    buffered kernel for all three contractions (B=4096: 512 output tiles): one fused step vs the oracle, 1e-2."""
    from oracle import codae_oracle as O
    from oracle.philox import philox_mask_table
    from codae import _C
    from codae.dataset import ConcatenatedEmbeddingDataset
    from codae.model import EmbeddingDenoisingAutoencoder
    from codae.tool import Corrupter, FusedStep
    torch.manual_seed(21)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    S, E, N, B = 8, 512, 4096, 4096
    cats = [torch.randn(N, E).abs() * (torch.rand(N, E) < 0.7) for _ in range(S)]
    ds = ConcatenatedEmbeddingDataset.from_tensors(cats)
    model = EmbeddingDenoisingAutoencoder(S * E, S * E, E, 2, 2, False)          # 6 x Linear(4096, 4096)
    W = [l.weight.detach().clone() for l in model.linears()]
    b = [l.bias.detach().clone() for l in model.linears()]
    data_cpu = ds.data.clone()
    model.set_compute_dtype("bf16")
    model.to(DEV)
    ds.to(DEV)
    cor = Corrupter(N, ds.arch, 2, DEV, seed=7)
    assert cor.nb_run == 36
    tbl = torch.from_numpy(philox_mask_table(7, N, 36).astype(np.int64))
    assert torch.equal(tbl, cor.mask_to_use)
    fs = FusedStep(model, cor, ds.data, lr=1e-4, weight_decay=1e-2, clip=True)
    assert fs.eng == _C.BF16
    idx = torch.randperm(N)[:B]
    fs.step(idx.to(DEV), run=3)
    bm, nm, _ = O.binary_masks(ds.arch, 2)
    _, fmask = O.get_masks(bm, nm, tbl, idx.tolist(), 3, 2)
    dae = O.OracleDAE(W, b, model.relu, 1e-4, 1e-2, True)
    r = dae.step_embedding(data_cpu[idx], fmask)
    assert abs(fs.last_loss(B) - r["loss"]) <= 1e-2 * abs(r["loss"])
    y = fs._bufs[B]["acts"][-1][:, :S * E]
    assert rel(y.cpu().numpy(), r["y"].numpy()) < 1e-2
    assert torch.equal(fs.last_mask_ids(B).cpu().long(), tbl[idx, 3])
    gg, gw = flat_grads(model), np.concatenate([t.numpy().ravel() for t in r["grads"]])
    assert float(np.linalg.norm(gg - gw) / np.linalg.norm(gw)) < 2e-2
    got, want = flat_params(model), np.concatenate([t.numpy().ravel() for t in dae.params()])
    assert (np.abs(got - want) <= 1e-2 * np.abs(want).max()).mean() > 0.999


def test_full_size_polyvore_gradient_linearity():
    """BASELINE's largest training config (10 x Linear(4096, 4096), B = 8192 per GPU, k_max = 2, bf16) is too big for the
    oracle in seconds; size-independent property instead: with the GLOBAL batch size in the loss scale, the gradient of
    the whole batch equals the sum of the gradients of its two halves (what data parallelism relies on).  lr = 0 keeps
    the weights fixed between the three steps."""
    from codae import _C
    from codae.model import EmbeddingDenoisingAutoencoder
    from codae.tool import Corrupter, FusedStep
    torch.manual_seed(31)
    S, E, N, B = 8, 512, 16384, 8192
    io = S * E
    g = torch.Generator(device=DEV).manual_seed(31)
    data = torch.rand((N, io), generator=g, device=DEV) * (torch.rand((N, io), generator=g, device=DEV) < 0.7)
    arch = [dict(name=str(i), size=E, type="regression", position=i * E) for i in range(S)]
    model = EmbeddingDenoisingAutoencoder(io, io, E, 4, 4, False)
    assert len(model.dims) == 10 and model.nb_parameters() == 167_813_120
    model.set_compute_dtype("bf16")
    model.to(DEV)
    cor = Corrupter(N, arch, 2, DEV, seed=5)
    fs = FusedStep(model, cor, data, lr=0.0, weight_decay=0.0, clip=True)
    assert fs.eng == _C.BF16
    w0 = model.flat.clone()
    idx = torch.randperm(N, device=DEV)[:B]
    fs.step(idx, run=1, global_batch=B)
    g_full, loss_full = fs.gflat.clone(), fs.last_loss(B)
    fs.step(idx[:B // 2], run=1, global_batch=B)
    g_a, sum_a = fs.gflat.clone(), float(fs.acc[3].item())
    fs.step(idx[B // 2:], run=1, global_batch=B)
    g_b, sum_b = fs.gflat.clone(), float(fs.acc[3].item())
    assert torch.equal(model.flat, w0)                                   # lr = 0: nothing moved
    assert abs((sum_a + sum_b) / (B * io) - loss_full) <= 1e-4 * loss_full
    err = float((g_a + g_b - g_full).norm() / g_full.norm())
    assert err < 1e-2, err
    assert bool(torch.isfinite(g_full).all()) and float(g_full.abs().max()) > 0


def test_ragged_last_batch_and_validation_pass():
    """No drop_last in the reference: the last batch is smaller; evaluate() = forward + monitors only."""
    from codae.tool import FusedStep
    g = np.load(os.path.join(GOLDEN, "emb_small.npz"))
    ds, model, cor = build_embedding(g)
    fs = FusedStep(model, cor, ds.data, lr=1e-3, weight_decay=0.0, clip=False)
    before = flat_params(model)
    out = fs.evaluate(torch.arange(5, device=DEV), run=0)
    assert np.array_equal(flat_params(model), before) and out.shape == (5, 48)
    x = ds.data[:5]
    m = fs.read_monitors()
    assert abs(m["full"] - float(((x - out) ** 2).sum())) <= 1e-5 * m["full"] and m["rows"] == 5
    fs.step(torch.arange(8, device=DEV))
    fs.step(torch.arange(8, 11, device=DEV))     # ragged: B = 3
    assert not np.array_equal(flat_params(model), before)
    assert np.isfinite(fs.last_loss(3))


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_train_epoch_device_sampler_matches_step_loop(dtype):
    """FusedStep.train_epoch (permutation drawn on the device, K steps per CUDA-graph launch, ragged last batch through the
    per-step path) must leave exactly the weights and monitors of the same steps issued one by one from the host
    (train_dae_on_embedding.py:118-128,194-223: SubsetRandomSampler + DataLoader(batch_size) + the loop body)."""
    from codae.tool import FusedStep
    g = np.load(os.path.join(GOLDEN, "emb_mid.npz"))
    B, n = 16, 150                                  # 9 full batches + a ragged one of 6 rows; 3 epochs
    train_idx = torch.arange(5, 5 + n, dtype=torch.int64, device=DEV)
    out = {}
    for mode in ("epoch", "loop"):
        ds, model, cor = build_embedding(g, dtype)
        fs = FusedStep(model, cor, ds.data, lr=float(g["lr"]), weight_decay=float(g["wd"]), clip=bool(g["clip"]), use_graph=False)
        gen = torch.Generator(device=DEV)
        gen.manual_seed(77)
        steps = 0
        for epoch in range(3):
            if mode == "epoch":
                steps += fs.train_epoch(train_idx, B, generator=gen, graph_steps=4)
            else:
                perm = train_idx[torch.randperm(n, device=DEV, generator=gen)]
                for lo in range(0, n, B):
                    fs.step(perm[lo:lo + B], run=0)
                    steps += 1
        torch.cuda.synchronize()
        out[mode] = (model.flat.clone(), fs.read_monitors(), steps, fs.m.clone())
    assert out["epoch"][2] == out["loop"][2] == 30
    assert torch.equal(out["epoch"][0], out["loop"][0])            # same kernels, same order, same inputs: bit-identical weights
    assert torch.equal(out["epoch"][3], out["loop"][3])
    for k in ("full", "partial", "rows"):
        assert out["epoch"][1][k] == out["loop"][1][k]


@pytest.mark.parametrize("dtype,B", [("bf16", 128), ("fp32", 128), ("bf16", 2048)])
def test_multi_step_graphs_are_bit_identical_to_the_step_loop(dtype, B):
    """FusedStep.train_steps: chunks of graph_steps consecutive steps as ONE CUDA-graph launch each, at embedding.yaml's layer
    width.  Same kernels, same values, same order per buffer => weights, moments, the weight copy the GEMMs read and the monitors
    must equal the one-step-at-a-time loop exactly, replay after replay."""
    from codae.dataset import ConcatenatedEmbeddingDataset
    from codae.model import EmbeddingDenoisingAutoencoder
    from codae.tool import Corrupter, FusedStep
    S, E, N = 3, 512, 4096
    torch.manual_seed(5)
    cats = [torch.randn(N, E).abs() * (torch.rand(N, E) < 0.7) for _ in range(S)]
    rows = torch.stack([torch.randperm(N)[:B] for _ in range(13)]).to(DEV)          # 13 steps: 1 eager + 3 graphs of 4
    out = {}
    for mode in ("graphs", "loop"):
        torch.manual_seed(6)
        ds = ConcatenatedEmbeddingDataset.from_tensors(cats)
        model = EmbeddingDenoisingAutoencoder(S * E, S * E, E, 2, 2, False)           # 6 x Linear(1536, 1536)
        model.set_compute_dtype(dtype)
        model.to(DEV)
        ds.to(DEV)
        cor = Corrupter(N, ds.arch, 1, DEV, seed=9)
        fs = FusedStep(model, cor, ds.data, lr=1e-4, weight_decay=1e-4, clip=True, use_graph=False)
        if mode == "graphs":
            assert fs.train_steps(rows, graph_steps=4) == 13
            assert any(k[0] == "steps" for k in fs._graphs)
        else:
            for j in range(13):
                fs.step(rows[j])
        torch.cuda.synchronize()
        out[mode] = (model.flat.clone(), fs.m.clone(), fs.v.clone(), fs.read_monitors(), model.gemm_weights().clone())
    a, b = out["graphs"], out["loop"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert torch.equal(a[4].view(torch.int16), b[4].view(torch.int16))
    for k in ("full", "partial", "rows", "last_sum"):
        assert a[3][k] == b[3][k], k


def test_polyvore_sized_step_vs_oracle():
    """The configuration bench.py's headline line runs -- polyvore-shaped DAE, 10 x Linear(4096, 4096), B = 8192, k_max = 2, bf16
    tensor-core engine (persistent TMEM-double-buffered GEMMs, bulk-store epilogues, sum(dW^2) from the weight-gradient launches,
    CUDA-graph replay) -- against the oracle on the same seeded inputs: loss and reconstruction within 1e-2 (BASELINE.json's bf16
    tolerance), post-Adam weights within the largest possible Adam step.  Two steps: eager, then captured + replayed."""
    from oracle import codae_oracle as O
    from oracle.philox import philox_mask_table
    from codae.dataset import ConcatenatedEmbeddingDataset
    from codae.model import EmbeddingDenoisingAutoencoder
    from codae.tool import Corrupter, FusedStep
    torch.manual_seed(13)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    S, E, N, B, lr, wd = 8, 512, 3 * 8192, 8192, 1e-5, 1e-4
    cats = [torch.randn(N, E).abs() * (torch.rand(N, E) < 0.7) for _ in range(S)]
    ds = ConcatenatedEmbeddingDataset.from_tensors(cats)
    model = EmbeddingDenoisingAutoencoder(S * E, S * E, E, 4, 4, False)
    assert len(model.dims) == 10 and all(d == (4096, 4096) for d in model.dims)
    W = [l.weight.detach().clone() for l in model.linears()]
    b = [l.bias.detach().clone() for l in model.linears()]
    data_cpu = ds.data.clone()
    model.set_compute_dtype("bf16")
    model.to(DEV)
    ds.to(DEV)
    cor = Corrupter(N, ds.arch, 2, DEV, seed=31)
    nb_run = S + S * (S - 1) // 2
    tbl = torch.from_numpy(philox_mask_table(31, N, nb_run).astype(np.int64))
    assert torch.equal(tbl, cor.mask_to_use)
    fs = FusedStep(model, cor, ds.data, lr=lr, weight_decay=wd, clip=True, use_graph=True)
    assert fs.eng == 1 and fs.wgrad_sqnorm
    dae = O.OracleDAE(W, b, model.relu, lr, wd, True)
    bm, nm, _ = O.binary_masks(ds.arch, 2)
    perm = torch.randperm(N)
    for s in range(3):                     # eager, captured, replayed
        idx = perm[s * B:(s + 1) * B]
        fs.step(idx.to(DEV), run=0)
        _, fmask = O.get_masks(bm, nm, tbl, idx.tolist(), 0, 2)
        r = dae.step_embedding(data_cpu[idx], fmask)
        assert abs(fs.last_loss(B) - r["loss"]) <= 1e-2 * abs(r["loss"]), (s, fs.last_loss(B), r["loss"])
        y = fs._bufs[B]["acts"][-1][:, :S * E]
        assert rel(y.cpu().numpy(), r["y"].numpy()) < 1e-2, s
        got = flat_params(model)
        want = np.concatenate([t.numpy().ravel() for t in dae.params()])
        dw = np.abs(got - want)
        assert (dw <= 1e-2 * np.abs(want).max()).mean() > 0.98 and dw.max() <= 2.2 * lr / (1 - 0.9) + 1e-2 * np.abs(want).max(), s
        # per-step parity: restart the oracle from the device state (as test_full_size_step_vs_oracle does)
        with torch.no_grad():
            for l in range(len(model.dims)):
                dae.W[l].copy_(model.weight_view(model.flat, l).cpu())
                dae.b[l].copy_(model.bias_view(model.flat, l).cpu())
            for j, (mt, vt) in enumerate(zip(dae.m, dae.v)):
                l, is_bias = j // 2, j % 2
                view = model.bias_view if is_bias else model.weight_view
                mt.copy_(view(fs.m, l).cpu())
                vt.copy_(view(fs.v, l).cpu())
