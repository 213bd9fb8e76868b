"""Pins oracle/codae_oracle.py against the golden vectors minted from the unmodified reference
(oracle/gen_golden.py).  CPU only.  Tolerances: bit-exact for masks / ids / layer tables;
fp32 arithmetic restated with the same torch ops is compared at 1e-6 relative (thread-count
differences move last bits)."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import codae_oracle as O
from oracle import philox

from conftest import GOLDEN


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def split_params(flat, shapes):
    out, o = [], 0
    for sh in shapes:
        sh = [int(s) for s in sh if s > 0]
        n = int(np.prod(sh))
        out.append(torch.from_numpy(flat[o:o + n].copy()).reshape(sh))
        o += n
    assert o == len(flat)
    return out


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def test_layer_tables():
    with open(os.path.join(GOLDEN, "layer_tables.json")) as f:
        tab = json.load(f)
    for row in tab["rows"]:
        dims, nenc = O.layer_dims(row["io"], row["z"], row["nin"], row["nout"], bool(row["steep"]),
                                  mixed=row["kind"] == "mixed")
        assert [(a, b) for a, b, _ in dims] == [tuple(d) for d in row["dims"]], row
        assert [int(r) for _, _, r in dims] == row["relu"], row
        assert nenc == row["nenc"]
    assert tab["errors"]["embedding_steep"].startswith("UnboundLocalError")
    with pytest.raises(UnboundLocalError):
        O.layer_dims(48, 48, 2, 2, True, mixed=False)
    with pytest.raises(UnboundLocalError):
        O.layer_dims(48, 8, 1, 2, False, mixed=False)


def test_corrupter_tables():
    g = load("corrupter_tables")
    for tag, k in [("v3k1", 1), ("v3k2", 2), ("v9k1", 1), ("v9k3", 3), ("v8k2", 2)]:
        sizes = g[tag + "_sizes"]
        arch, pos = [], 0
        for s in sizes:
            arch.append(dict(size=int(s), position=pos))
            pos += int(s)
        bm, nmiss, per_k = O.binary_masks(arch, k)
        assert np.array_equal(bm.numpy(), g[tag + "_binary_masks"])
        assert nmiss == list(g[tag + "_nb_missing_per_run"])
        assert per_k == list(g[tag + "_nb_corruption_per_k"])
        assert len(nmiss) == int(g[tag + "_nb_run"])
        random.seed(1234)
        tbl = O.mask_table_compat(17, len(nmiss))
        assert np.array_equal(tbl.numpy(), g[tag + "_mask_to_use"])
    msg = str(g["raises_3"])
    with pytest.raises(Exception) as e:
        O.check_k_max([0, 1, 2], 3)
    assert str(e.value) == msg
    with pytest.raises(Exception):
        O.check_k_max([0, 1, 2], -1)
    O.check_k_max([0, 1, 2], 2)


def _embedding_arch(io, e):
    return [dict(size=e, position=p, type="regression") for p in range(0, io, e)]


@pytest.mark.parametrize("name", ["emb_small", "emb_bottleneck", "emb_k2", "emb_mid"])
def test_embedding_steps(name):
    torch.set_num_threads(1)
    g = load(name)
    io, e, z = int(g["io"]), int(g["e"]), int(g["z"])
    dims, _ = O.layer_dims(io, z, int(g["nin"]), int(g["nout"]), False)
    params = split_params(g["init"], g["shapes"])
    W, b = params[0::2], params[1::2]
    assert [tuple(w.shape) for w in W] == [(o, i) for i, o, _ in dims]
    arch = _embedding_arch(io, e)
    k_max = int(g["k_max"])
    bm, nmiss, _ = O.binary_masks(arch, k_max)
    assert np.array_equal(bm.numpy(), g["binary_masks"])
    dae = O.OracleDAE(W, b, [r for _, _, r in dims], float(g["lr"]), float(g["wd"]), bool(g["clip"]))
    data = torch.from_numpy(g["data"])
    tbl = torch.from_numpy(g["mask_to_use"])
    s = 0
    while "idx%d" % s in g:
        idx = [int(i) for i in g["idx%d" % s]]
        x = data[idx]
        _, fmask = O.get_masks(bm, nmiss, tbl, idx, 0, k_max)
        assert np.array_equal(fmask.numpy(), g["fmask%d" % s])
        r = dae.step_embedding(x, fmask)
        assert np.array_equal(r["cx"].numpy(), g["cx%d" % s])
        assert rel(r["y"].numpy(), g["y%d" % s]) < 1e-6
        assert abs(r["loss"] - float(g["loss%d" % s])) <= 1e-6 * abs(float(g["loss%d" % s]))
        gr = np.concatenate([t.numpy().ravel() for t in r["grads"]])
        assert rel(gr, g["grads%d" % s]) < 1e-5
        if bool(g["clip"]):
            assert abs(r["grad_norm"] - float(g["gnorm%d" % s])) <= 1e-5 * float(g["gnorm%d" % s])
        post = np.concatenate([t.numpy().ravel() for t in dae.params()])
        assert rel(post, g["post%d" % s]) < 1e-6
        # the update itself (not just the weights) must match: compare the step taken
        prev = g["init"] if s == 0 else g["post%d" % (s - 1)]
        assert rel(post - prev_oracle(s, dae, g), g["post%d" % s] - prev) < 2e-3
        assert abs(r["full"] - float(g["ftl%d" % s])) <= 1e-5 * float(g["ftl%d" % s])
        assert abs(r["partial"] - float(g["ptl%d" % s])) <= 1e-5 * float(g["ptl%d" % s])
        s += 1
    assert s >= 2


_prev_cache = {}


def prev_oracle(s, dae, g):
    """weights the oracle held before step s (recomputed from the golden chain: the oracle matched
    it to 1e-6 at s-1, so the reference's own previous weights are the right base)."""
    return g["init"] if s == 0 else g["post%d" % (s - 1)]


@pytest.mark.parametrize("name", ["emb_small", "emb_k2"])
def test_ranking_loss(name):
    g = load(name)
    e = int(g["e"])
    ncat = int(g["io"]) // e
    cats = [torch.from_numpy(g["cat%d" % c]) for c in range(ncat)]
    pred = torch.from_numpy(g["rank_pred"])
    fmask = torch.from_numpy(g["rank_fmask"])
    if int(g["k_max"]) > 1:
        # RankingLoss is only defined for one masked slot (metering.py:56); restrict to such rows
        keep = [(1 - fmask[i]).sum().item() == e for i in range(len(fmask))]
        if not all(keep):
            pytest.skip("k>1 rows present: reference semantics undefined")
    total, ranks = O.ranking_loss(cats, e, [int(v) for v in g["rank_val"]], pred, fmask,
                                  [int(i) for i in g["rank_idx"]])
    assert abs(total - float(g["rank_loss"])) < 1e-9
    # the same ranks through the generic scoring restatement (fp64)
    getter_rows = (1 - fmask).reshape(len(fmask), ncat, e)[:, :, 0]
    for i, idx in enumerate(g["rank_idx"]):
        c = int(torch.argmax(getter_rows[i]))
        s = O.score_candidates(cats[c], pred[i, c * e:(c + 1) * e], metric="cosine")
        assert O.rank_of(s, int(idx), [int(v) for v in g["rank_val"]], "cosine") == ranks[i]


@pytest.mark.parametrize("name", ["abalone_k1", "abalone_k3"])
def test_abalone_steps(name):
    torch.set_num_threads(1)
    g = load(name)
    from oracle.gen_golden import abalone_arch
    arch = abalone_arch()
    dims, _ = O.layer_dims(11, int(g["z"]), 2, 2, bool(g["steep"]), mixed=True)
    params = split_params(g["init"], g["shapes"])
    W, b = params[0::2], params[1::2]
    assert [tuple(w.shape) for w in W] == [(o, i) for i, o, _ in dims]
    k_max = int(g["k_max"])
    bm, nmiss, per_k = O.binary_masks(arch, k_max)
    assert np.array_equal(bm.numpy(), g["binary_masks"])
    assert per_k == list(g["nb_corruption_per_k"])
    type_mask = [0, 0, 0] + [1] * 8
    T = O.mask_transformation(type_mask, [0] * 9)
    assert np.array_equal(T.numpy(), g["mask_transformation"])
    dae = O.OracleDAE(W, b, [r for _, _, r in dims], float(g["lr"]), float(g["wd"]), True)
    data = torch.from_numpy(g["data"])
    tbl = torch.from_numpy(g["mask_to_use"])
    scale, mn = torch.from_numpy(g["norm_scale"]), torch.from_numpy(g["norm_min"])
    s = 0
    while "idx%d" % s in g:
        idx = [int(i) for i in g["idx%d" % s]]
        x = data[idx]
        masks, fmask = O.get_masks(bm, nmiss, tbl, idx, int(g["run%d" % s]), k_max)
        assert np.array_equal(fmask.numpy(), g["fmask%d" % s])
        for k in range(k_max):
            assert np.array_equal(masks[k].numpy(), g["mask%d_k%d" % (s, k)])
        r = dae.step_mixed(arch, g["weight"], x, fmask)
        assert np.array_equal(r["cx"].numpy(), g["cx%d" % s])
        assert rel(r["y"].numpy(), g["y%d" % s]) < 1e-6
        assert abs(r["loss"] - float(g["loss%d" % s])) <= 2e-6 * abs(float(g["loss%d" % s]))
        gr = np.concatenate([t.numpy().ravel() for t in r["grads"]])
        assert rel(gr, g["grads%d" % s]) < 1e-5
        assert abs(r["grad_norm"] - float(g["gnorm%d" % s])) <= 1e-5 * float(g["gnorm%d" % s])
        post = np.concatenate([t.numpy().ravel() for t in dae.params()])
        assert rel(post, g["post%d" % s]) < 1e-6
        # monitors on de-normalised values (train_dae_on_abalone.py:227-236)
        xd, yd = x.clone(), r["y"].clone()
        xd[:, 3:] = O.normalizer_undo(xd[:, 3:], scale, mn)
        yd[:, 3:] = O.normalizer_undo(yd[:, 3:], scale, mn)
        ml = O.combined_full_loss(arch, xd, yd).numpy()
        assert rel(ml, g["mon%d" % s]) < 1e-5
        assert rel(O.per_k(ml, masks, T), g["mon_per_k%d" % s]) < 1e-5
        pl = O.partial(ml, fmask, T)
        assert rel(pl, g["mon_partial%d" % s]) < 1e-5
        assert rel(O.per_k(pl, masks, T), g["mon_partial_per_k%d" % s]) < 1e-5
        s += 1
    assert s == 3


def test_normalizer_roundtrip():
    d = torch.rand(5, 8)
    sc, mn = torch.rand(8) + 0.5, torch.rand(8)
    assert torch.allclose(O.normalizer_do(O.normalizer_undo(d, sc, mn), sc, mn), d, atol=1e-6)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10 (published algorithm, Salmon et al. SC'11)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = philox.philox4x32_10(*c, *k)
        assert tuple(int(x) for x in got) == want


def test_philox_table_is_permutation_and_shardable():
    t = philox.philox_mask_table(27493045, 500, 9)
    assert t.dtype == np.int16
    assert (np.sort(t, axis=1) == np.arange(9)).all()
    # ids depend only on (seed, observation): a shard starting at row 123 reproduces rows 123..
    assert np.array_equal(philox.philox_mask_table(27493045, 50, 9, first_observation=123), t[123:173])
    assert not np.array_equal(philox.philox_mask_table(1, 500, 9), t)
    counts = np.bincount(t[:, 0].astype(int), minlength=9)
    assert counts.min() > 25  # roughly uniform first draw


def test_topk_restatement_ties_and_merge():
    cat = torch.tensor([[0.0, 0], [1, 0], [1, 0], [0, 1], [3, 3]])
    q = torch.tensor([0.0, 0])
    s = O.score_candidates(cat, q, "sqerr")
    sc, idx = O.topk(s, 3, "sqerr")
    assert idx.tolist() == [0, 1, 2]  # tie between 1,2,3 -> lower index first
    s1, i1 = O.topk(s[:2], 2, "sqerr", row_offset=0)
    s2, i2 = O.topk(s[2:], 2, "sqerr", row_offset=2)
    ms, mi = O.topk_merge([s1, s2], [i1, i2], 3, "sqerr")
    assert mi.tolist() == [0, 1, 2]
    c = O.score_candidates(cat, torch.tensor([1.0, 0]), "cosine")
    _, ci = O.topk(c, 2, "cosine")
    assert ci.tolist() == [1, 2]
