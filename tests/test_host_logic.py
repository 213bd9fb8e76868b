"""CPU: host-side mirror of the reference interface (constructors, attributes, exceptions, samplers, shards)."""
import json
import os
import random
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, PKG

from codae.dataset import ConcatenatedEmbeddingDataset, MixedVariableDataset
from codae.model import EmbeddingDenoisingAutoencoder, MixedVariableDenoisingAutoencoder
from codae.tool import Corrupter, CombinedCriterion, Normalizer, collate_embedding, get_mask_transformation, get_rmse
from codae.tool.inference import shard_rows


def test_layer_rule_state_dict_and_errors():
    tab = json.load(open(os.path.join(GOLDEN, "layer_tables.json")))
    for r in tab["rows"]:
        if r["kind"] == "embedding":
            e = 512 if r["io"] % 512 == 0 else (16 if r["io"] == 48 else 64)
            m = EmbeddingDenoisingAutoencoder(r["io"], r["z"], e, r["nin"], r["nout"], False)
        else:
            m = MixedVariableDenoisingAutoencoder([], r["io"], r["z"], "cpu", r["nin"], r["nout"], bool(r["steep"]))
        assert [list(d) for d in m.dims] == r["dims"]
        assert [int(x) for x in m.relu] == r["relu"]
        assert m.nb_encoder == r["nenc"]
        assert list(m.state_dict().keys()) == r["keys"]
    with pytest.raises(Exception, match="io_size must be a multiple of embedding_size"):
        EmbeddingDenoisingAutoencoder(50, 8, 16, 2, 2, False)
    with pytest.raises(UnboundLocalError):
        EmbeddingDenoisingAutoencoder(48, 48, 16, 2, 2, True)
    with pytest.raises(UnboundLocalError):
        EmbeddingDenoisingAutoencoder(48, 8, 16, 1, 2, False)
    m = MixedVariableDenoisingAutoencoder([], 11, 11, "cpu")
    with pytest.raises(Exception, match="invalid corruption type"):
        m.corrupt(torch.zeros(1, 11), torch.ones(1, 11), corruption_type="gaussian")


def test_seeded_init_matches_reference_weights():
    """Same module construction order => same torch RNG consumption => the golden run's initial weights."""
    g = np.load(os.path.join(GOLDEN, "emb_small.npz"))
    # the generator seeds, builds the synthetic embeddings with a private Generator, then Corrupter (python random),
    # then the model: torch's global RNG is untouched before the model is built.
    torch.manual_seed(int(g["seed"]))
    m = EmbeddingDenoisingAutoencoder(int(g["io"]), int(g["z"]), int(g["e"]), int(g["nin"]), int(g["nout"]), False)
    flat = np.concatenate([p.detach().numpy().ravel() for p in m.parameters()])
    assert np.array_equal(flat, g["init"])


def test_flat_layout_is_16_byte_aligned_and_padded():
    m = EmbeddingDenoisingAutoencoder(48, 8, 16, 3, 3, False)
    lay, total = m.layout()
    assert total % 8 == 0
    for (w_off, ld, bcol), (i, o) in zip(lay, m.dims):
        assert w_off % 8 == 0 and ld % 8 == 0 and i <= bcol < ld        # bias column after the (padded) weights
    assert m.nb_parameters() == sum(p.numel() for p in m.parameters())
    flat = torch.arange(total, dtype=torch.float32)
    for l, (i, o) in enumerate(m.dims):
        assert m.weight_view(flat, l).shape == (o, i) and m.bias_view(flat, l).shape == (o,)
        assert m.aug_view(flat, l).shape == (o, lay[l][2] + 1)
        assert m.bias_view(flat, l)[0] == m.aug_view(flat, l)[0, -1]


def test_corrupter_attributes_match_reference():
    g = np.load(os.path.join(GOLDEN, "corrupter_tables.npz"))
    for tag, k in [("v3k1", 1), ("v3k2", 2), ("v9k1", 1), ("v9k3", 3), ("v8k2", 2)]:
        arch, pos = [], 0
        for s in g[tag + "_sizes"]:
            arch.append(dict(size=int(s), position=pos))
            pos += int(s)
        random.seed(1234)
        c = Corrupter(17, arch, k, torch.device("cpu"))
        assert c.nb_run == int(g[tag + "_nb_run"])
        assert np.array_equal(c.binary_masks.numpy(), g[tag + "_binary_masks"])
        assert c.nb_missing_per_run == list(g[tag + "_nb_missing_per_run"])
        assert c.nb_corruption_per_k == list(g[tag + "_nb_corruption_per_k"])
        assert np.array_equal(c.mask_to_use.numpy(), g[tag + "_mask_to_use"])
    for bad in (-1, 3):
        with pytest.raises(Exception) as e:
            Corrupter(2, [dict(size=1, position=i) for i in range(3)], bad, torch.device("cpu"))
        assert str(e.value) == str(g["raises_%d" % bad])


def test_concatenated_dataset_matches_reference():
    g = np.load(os.path.join(GOLDEN, "emb_small.npz"))
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "oracle"))
    from oracle.gen_golden import synth_embeddings
    gen = torch.Generator().manual_seed(int(g["seed"]))
    emb = synth_embeddings(96, ["top", "bottom", "shoe"], 16, gen)
    emb["incomplete"] = {"top": [0.0] * 16}            # dropped: lacks a used category
    ds = ConcatenatedEmbeddingDataset(embeddings=emb, used_category=["top", "bottom", "shoe"])
    assert ds.nb_observation == 96 and ds.embedding_size == 16 and ds.nb_predictor == 48
    assert np.array_equal(ds.data.numpy(), g["data"])
    assert abs(ds.scale - float(g["scale"])) == 0
    for c in range(3):
        assert np.array_equal(ds.data_per_category[c].numpy(), g["cat%d" % c])
    assert [a["position"] for a in ds.arch] == [0, 16, 32] and all(a["type"] == "regression" for a in ds.arch)
    row, idx = ds[5]
    assert idx == 5 and torch.equal(row, ds.data[5])
    rows, ids = collate_embedding([ds[1], ds[7]])
    assert rows.shape == (2, 48) and ids == (1, 7)


def test_mixed_dataset_arch():
    import pandas as pd
    df = pd.DataFrame({"Sex": ["M", "F", "M", "I"], "Length": [0.1, 0.2, 0.3, 0.4], "Rings": [1.0, 2.0, 3.0, 4.0]})
    ds = MixedVariableDataset(df)
    assert [(a["size"], a["type"], a["position"]) for a in ds.arch] == [(3, "classification", 0), (1, "regression", 3), (1, "regression", 4)]
    assert ds.data[:, :3].tolist() == [[1, 0, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1]]   # first-appearance label order
    assert ds.type_mask.tolist() == [0, 0, 0, 1, 1]


def test_mask_transformation_and_normalizer_and_rmse():
    g = np.load(os.path.join(GOLDEN, "abalone_k1.npz"))
    T = get_mask_transformation([0, 0, 0] + [1] * 8, [0] * 9)
    assert np.array_equal(T.numpy(), g["mask_transformation"])

    class S:
        data_min_, data_max_, data_range_ = np.array([1.0, 2.0]), np.array([3.0, 6.0]), np.array([2.0, 4.0])
    n = Normalizer(S, "cpu")
    d = torch.tensor([[0.5, 0.25]])
    assert torch.allclose(n.undo(d), torch.tensor([[2.0, 3.0]]))
    assert torch.allclose(n.do(n.undo(d)), d)
    assert abs(get_rmse(np.array([1.0, 2.0]), np.array([1.0, 4.0])) - np.sqrt(2.0)) < 1e-12


def test_combined_criterion_host_monitors():
    g = np.load(os.path.join(GOLDEN, "abalone_k3.npz"))
    from oracle.gen_golden import abalone_arch
    crit = CombinedCriterion(abalone_arch(), 3, "cpu", torch.tensor([0, 0, 0] + [1] * 8), reduction="none")
    masks = [torch.from_numpy(g["mask0_k%d" % k]) for k in range(3)]
    assert np.allclose(crit.get_per_k(g["mon0"], masks), g["mon_per_k0"], rtol=1e-6)
    pl = crit.get_partial(g["mon0"], torch.from_numpy(g["fmask0"]))
    assert np.allclose(pl, g["mon_partial0"], rtol=1e-6)
    with pytest.raises(Exception, match="Unknown reduction type."):
        CombinedCriterion(abalone_arch(), 1, "cpu", torch.tensor([0, 0, 0] + [1] * 8), reduction="sum")(torch.zeros(1, 11), torch.zeros(1, 11))


def test_shard_rows_and_epoch_batches():
    for n, g in [(10, 4), (10_000_000, 8), (7, 8), (0, 2)]:
        spans = [shard_rows(n, g, r) for r in range(g)]
        assert sum(c for _, c in spans) == n
        for (lo, c), (lo2, _) in zip(spans, spans[1:]):
            assert lo + c == lo2 or c == 0
    sys.path.insert(0, os.path.join(PKG, "script"))
    from _common import epoch_batches
    idx = list(range(100, 145))
    seen = []
    for r in range(2):
        rng = np.random.RandomState(3)
        for local, gb in epoch_batches(idx, 16, rng, r, 2):
            seen += list(local)
            assert gb in (16, 13)
    assert sorted(seen) == idx          # ranks partition every global batch, ragged last batch included


def test_yaml_configs_keep_the_reference_schema():
    """The three shipped configs carry the reference's keys and values (checked against a snapshot of the parsed
    reference files taken in the build container), plus the new polyvore-shaped config with the same schema."""
    import yaml
    cfg = os.path.join(PKG, "config")
    emb = yaml.safe_load(open(os.path.join(cfg, "embedding.yaml")))
    assert emb["MODEL"] == {"Z_SIZE": 1536, "BATCH_SIZE": 128, "NB_INPUT_LAYER": 4, "NB_OUTPUT_LAYER": 4, "STEEP_LAYER_SIZE": False,
                            "EPOCH": 50, "LEARNING_RATE": 1e-05, "WEIGHT_DECAY": 0.0001, "NB_CORRUPTED": 1, "TRUNK_GRAD": True}
    assert emb["DATASET"] == {"NAME": "EMBEDDING", "USED_CATEGORY": ["top", "bottom", "shoe"], "EMBEDDING_SIZE": 512,
                              "SHUFFLE": True, "SPLIT": [0.7, 0.3]} and emb["SEED"] == 27493045
    mod = yaml.safe_load(open(os.path.join(cfg, "modanet_merge_top_bottom_shoe.yaml")))
    assert mod["MODEL"]["BATCH_SIZE"] == 32 and mod["MODEL"]["LEARNING_RATE"] == 1e-4 and mod["MODEL"]["WEIGHT_DECAY"] == 0.01
    assert "TRUNK_GRAD" not in mod["MODEL"] and mod["SEED"] == 50493213 and mod["DATASET"]["SPLIT"] == [0.8, 0.2]
    aba = yaml.safe_load(open(os.path.join(cfg, "abalone.yaml")))
    assert aba["MODEL"]["Z_SIZE"] == 11 and aba["MODEL"]["STEEP_LAYER_SIZE"] is True and aba["MODEL"]["TRUNK_GRAD"] is True
    assert aba["MODEL"]["LEARNING_RATE"] == 5e-05 and aba["MODEL"]["WEIGHT_DECAY"] == 1e-06 and aba["SEED"] == 27123045
    assert set(aba) == {"MODEL", "DATASET", "SEED", "EVALUATION", "PLOT"}
    poly = yaml.safe_load(open(os.path.join(cfg, "polyvore_multislot.yaml")))
    assert set(poly["MODEL"]) >= set(emb["MODEL"]) and len(poly["DATASET"]["USED_CATEGORY"]) == 8


def test_corrupter_refuses_what_the_device_tables_cannot_hold():
    """The device mask-id table is int16 and a mask is one 64-bit word of variables: sizes beyond that must fail loudly
    (the reference has no such limits; silently wrapping ids would corrupt rows out of bounds)."""
    arch = [dict(name=str(i), size=1, type="regression", position=i) for i in range(65)]
    with pytest.raises(Exception, match="at most 64"):
        Corrupter(4, arch, 1, torch.device("cpu"))
    arch = [dict(name=str(i), size=1, type="regression", position=i) for i in range(60)]
    with pytest.raises(Exception, match="32767"):
        Corrupter(4, arch, 3, torch.device("cpu"))          # 60 + 1770 + 34220 subsets
    Corrupter(4, arch, 2, torch.device("cpu"))              # 1830 subsets: fine


def test_split3_arithmetic_matches_fp32_to_2_pow_minus_24():
    """The three-plane format of the fp32-parity engine, in numpy: hi + mid + lo reproduces x to 2^-24 |x|, and the six products
    the kernel keeps (hh | hm, mh, mm, hl, lh) give an fp32-accurate contraction where plain bf16 is 2e-3 off."""
    rng = np.random.RandomState(0)

    def bf16(a):
        u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
        u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16                  # round to nearest even
        return u.astype(np.uint32).view(np.float32)

    def split(a):
        h = bf16(a)
        m = bf16(a - h)
        return h, m, bf16(a - h - m)
    x = (rng.randn(4096) * 10.0 ** rng.randint(-6, 6, 4096)).astype(np.float32)
    h, m, l = split(x)
    assert np.all(np.abs(h.astype(np.float64) + m + l - x) <= np.abs(x) * 2.0 ** -24)
    A, B = rng.randn(64, 512).astype(np.float32), (rng.randn(48, 512) / 22).astype(np.float32)
    (ah, am, al), (bh, bm, bl) = split(A), split(B)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    x3 = ah @ bh.T + (ah @ bm.T + am @ bh.T + am @ bm.T + ah @ bl.T + al @ bh.T)
    err = lambda y: np.abs(y - ref).max() / np.abs(ref).max()
    assert err(x3) < 1e-6 and err(A @ B.T) < 1e-6 and err(ah @ bh.T) > 1e-4


@pytest.mark.parametrize("n,B,W", [(45, 16, 2), (64, 16, 4), (45, 15, 2), (7, 16, 2), (128, 32, 1)])
def test_device_sampler_splits_an_epoch_like_the_host_sampler(n, B, W):
    """FusedStep.epoch_rows (the slicing behind train_epoch's device-side sampler) == script/_common.epoch_batches on the same
    permutation: consecutive global batches, rank r takes elements r::W, ragged / indivisible batches through the per-step path."""
    from codae.tool import FusedStep
    sys.path.insert(0, os.path.join(PKG, "script"))
    from _common import epoch_batches
    perm = torch.randperm(n) + 100

    class FixedPerm:                      # epoch_batches draws rng.permutation(len(indices)): hand it the identity
        @staticmethod
        def permutation(k):
            return np.arange(k)
    for r in range(W):
        want = [(list(loc), gb) for loc, gb in epoch_batches(perm.tolist(), B, FixedPerm, r, W)]
        local, tail = FusedStep.epoch_rows(perm, B, W, r)
        got = [] if local is None else [(row.tolist(), B) for row in local]
        got += [(rows.tolist(), gb) for rows, gb in tail]
        assert got == want
