"""ORACLE tooling: mint golden vectors by running the UNMODIFIED reference (imported from
/root/reference) in the build container.  The reference has no golden vectors of its own
(SURVEY.md section 4), so these fixtures are what pins oracle/codae_oracle.py and, through it,
the CUDA path.  Run:  python oracle/gen_golden.py   (writes tests/golden/*.npz; CPU, ~10 s).

The loop bodies replayed here are script/train_dae_on_embedding.py:198-223 and
script/train_dae_on_abalone.py:206-236, driven around the reference's own classes
(the scripts themselves need matplotlib and real data files -- SURVEY.md section 8c).
Every fixture records (seed, threads): seeded runs are bit-reproducible at a fixed thread count.
"""
import os
import random
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
THREADS = 1


def import_reference():
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.path.insert(0, REF)
    import codae.model as M
    import codae.tool as T
    import codae.dataset as D
    assert M.__file__.startswith(REF)
    return M, T, D


def seed_all(s):
    random.seed(s)
    np.random.seed(s % (2 ** 32))
    torch.manual_seed(s)


def flat(tensors):
    return np.concatenate([t.detach().numpy().ravel() for t in tensors])


def synth_embeddings(n, cats, e, gen):
    """{obs_id: {category: [e floats]}} -- the on-disk wire format (script/encode_coco.py:65-78),
    post-ReLU-like values."""
    out = {}
    for i in range(n):
        out[str(i)] = {}
        for c in cats:
            v = torch.randn(e, generator=gen).abs() * (torch.rand(e, generator=gen) < 0.7)
            out[str(i)][c] = v.tolist()
    return out


def embedding_case(M, T, D, name, seed, n, cats, e, z, nin, nout, B, steps, lr, wd, clip, k_max=1):
    seed_all(seed)
    gen = torch.Generator().manual_seed(seed)
    emb = synth_embeddings(n, cats, e, gen)
    _stdout = sys.stdout
    sys.stdout = open(os.devnull, "w")
    ds = D.ConcatenatedEmbeddingDataset(embeddings=emb, used_category=cats)
    sys.stdout = _stdout
    io = e * len(cats)
    corr = T.Corrupter(nb_observation=ds.nb_observation, arch=ds.arch, k_max=k_max, device=torch.device("cpu"))
    model = M.EmbeddingDenoisingAutoencoder(io_size=io, z_size=z, embedding_size=e, nb_input_layer=nin,
                                           nb_output_layer=nout, steep_layer_size=False)
    params = list(model.parameters())
    init = [p.detach().clone() for p in params]
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    mean_c = torch.nn.MSELoss(reduction="mean")
    full_c = torch.nn.MSELoss(reduction="none")
    perm = np.random.permutation(n)
    rec = dict(seed=seed, threads=THREADS, io=io, e=e, z=z, nin=nin, nout=nout, B=B, lr=lr, wd=wd,
               clip=int(clip), k_max=k_max, scale=ds.scale,
               data=ds.data.numpy().copy(), mask_to_use=corr.mask_to_use.numpy().copy(),
               binary_masks=corr.binary_masks.numpy().copy(),
               nb_missing_per_run=np.array(corr.nb_missing_per_run),
               shapes=np.array([list(p.shape) + [0] * (2 - p.dim()) for p in params]),
               init=flat(init))
    for c in range(len(cats)):
        rec["cat%d" % c] = ds.data_per_category[c].numpy().copy()
    for s in range(steps):
        bi = tuple(int(i) for i in perm[s * B:(s + 1) * B])
        input_data = torch.stack([ds[i][0] for i in bi])
        masks, fmask = corr.get_masks(bi, 0)
        c_in = model.corrupt(input_data=input_data, mask=fmask)
        out = model(c_in)
        loss = mean_c(input_data, out)
        opt.zero_grad()
        loss.backward()
        grads = flat([p.grad for p in params])
        if clip:
            gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
            rec["gnorm%d" % s] = float(gn)
        opt.step()
        fl = full_c(input_data, out).cpu().detach().numpy()
        rec["idx%d" % s] = np.array(bi)
        rec["fmask%d" % s] = fmask.numpy().copy()
        rec["cx%d" % s] = c_in.detach().numpy().copy()
        rec["y%d" % s] = out.detach().numpy().copy()
        rec["loss%d" % s] = float(loss)
        rec["grads%d" % s] = grads
        rec["post%d" % s] = flat(params)
        rec["ftl%d" % s] = float(np.sum(fl))
        rec["ptl%d" % s] = float(np.sum((1 - fmask.cpu().numpy()) * fl))
    # validation-style ranking (train_dae_on_embedding.py:241-261, metering.py:46-79)
    val = [int(i) for i in perm[-min(24, n // 2):]]
    rl = T.RankingLoss(ds, val, device=torch.device("cpu"))
    bi = tuple(val[:8])
    input_data = torch.stack([ds[i][0] for i in bi])
    masks, fmask = corr.get_masks(bi, 0)
    with torch.no_grad():
        out = model(model.corrupt(input_data=input_data, mask=fmask))
    rec["rank_val"] = np.array(val)
    rec["rank_idx"] = np.array(bi)
    rec["rank_fmask"] = fmask.numpy().copy()
    rec["rank_pred"] = out.numpy().copy()
    rec["rank_loss"] = float(rl.get(out, fmask, bi))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "loss", [rec["loss%d" % s] for s in range(steps)], "rank_loss", rec["rank_loss"])


def abalone_arch():
    """arch of the UCI abalone frame as MixedVariableDataset builds it
    (codae/dataset/mixed_variable_dataset.py:28-49): C3 (Sex) + 8 x R1.  Built directly because the
    pandas-3 incompatibilities at train_dae_on_abalone.py:90 / mixed_variable_dataset.py:36 stop the
    reference's own data prep here."""
    arch = [dict(name="Sex", size=3, type="classification", position=0)]
    arch[0]["lambda"] = 1
    for i, n in enumerate(["Length", "Diameter", "Height", "Whole", "Shucked", "Viscera", "Shell", "Rings"]):
        arch.append({"name": n, "lambda": 1, "size": 1, "type": "regression", "position": 3 + i})
    return arch


def abalone_case(M, T, name, seed, n, z, steep, B, steps, lr, wd, k_max):
    seed_all(seed)
    gen = torch.Generator().manual_seed(seed)
    arch = abalone_arch()
    io = 11
    data = torch.zeros(n, io)
    lab = torch.randint(0, 3, (n,), generator=gen)
    data[torch.arange(n), lab] = 1
    data[:, 3:] = torch.rand(n, 8, generator=gen)
    type_mask = torch.zeros(io)
    type_mask[3:] = 1

    class _Scaler:  # stands in for sklearn MinMaxScaler attributes read by Normalizer (data_tool.py:62-64)
        data_min_ = np.linspace(0.1, 0.8, 8)
        data_max_ = np.linspace(1.5, 30.0, 8)
        data_range_ = data_max_ - data_min_

    norm = T.Normalizer(normalizer=_Scaler, device=torch.device("cpu"))
    corr = T.Corrupter(nb_observation=n, arch=arch, k_max=k_max, device=torch.device("cpu"))
    model = M.MixedVariableDenoisingAutoencoder(arch=arch, io_size=io, z_size=z, device=torch.device("cpu"),
                                                nb_input_layer=2, nb_output_layer=2, steep_layer_size=steep)
    params = list(model.parameters())
    init = [p.detach().clone() for p in params]
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    w = [0.4, 1, 1, 1, 1, 1, 1, 1, 1]
    full_criterion = T.CombinedCriterion(arch=arch, k_max=k_max, device=torch.device("cpu"),
                                         observation_mask=type_mask, weight=w, reduction="mean")
    monitor = T.CombinedCriterion(arch=arch, k_max=k_max, device=torch.device("cpu"),
                                  observation_mask=type_mask, reduction="none")
    perm = np.random.permutation(n)
    rec = dict(seed=seed, threads=THREADS, io=io, z=z, steep=int(steep), B=B, lr=lr, wd=wd, k_max=k_max,
               data=data.numpy().copy(), mask_to_use=corr.mask_to_use.numpy().copy(),
               binary_masks=corr.binary_masks.numpy().copy(),
               nb_missing_per_run=np.array(corr.nb_missing_per_run),
               nb_corruption_per_k=np.array(corr.nb_corruption_per_k),
               shapes=np.array([list(p.shape) + [0] * (2 - p.dim()) for p in params]),
               init=flat(init), weight=np.array(w), norm_min=norm.min.numpy(), norm_scale=norm.scale.numpy(),
               mask_transformation=monitor.mask_transformation.copy())
    for s in range(steps):
        run = s % corr.nb_run
        bi = tuple(int(i) for i in perm[s * B:(s + 1) * B])
        input_data = torch.stack([data[i] for i in bi])
        masks, fmask = corr.get_masks(bi, run)
        c_in = model.corrupt(input_data=input_data, mask=fmask)
        out = model(c_in)
        loss = full_criterion(x=input_data, y=out)
        opt.zero_grad()
        loss.backward()
        grads = flat([p.grad for p in params])
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        opt.step()
        rec["y%d" % s] = out.detach().numpy().copy()
        rec["x%d" % s] = input_data.numpy().copy()
        input_data[:, 3:] = norm.undo(input_data[:, 3:])
        out = out.detach()
        out[:, 3:] = norm.undo(out[:, 3:])
        ml = monitor(input_data, out, as_numpy=True)
        rec["run%d" % s] = run
        rec["idx%d" % s] = np.array(bi)
        rec["fmask%d" % s] = fmask.numpy().copy()
        for k in range(k_max):
            rec["mask%d_k%d" % (s, k)] = masks[k].numpy().copy()
        rec["cx%d" % s] = c_in.detach().numpy().copy()
        rec["loss%d" % s] = float(loss)
        rec["grads%d" % s] = grads
        rec["gnorm%d" % s] = float(gn)
        rec["post%d" % s] = flat(params)
        rec["mon%d" % s] = ml.copy()
        rec["mon_per_k%d" % s] = monitor.get_per_k(ml, masks)
        pl = monitor.get_partial(ml, fmask)
        rec["mon_partial%d" % s] = pl.copy()
        rec["mon_partial_per_k%d" % s] = monitor.get_per_k(pl, masks)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "loss", [rec["loss%d" % s] for s in range(steps)])


def corrupter_tables(T):
    rec = {}
    for tag, V, size, k in [("v3k1", 3, 4, 1), ("v3k2", 3, 4, 2), ("v9k1", 9, 1, 1), ("v9k3", 9, 1, 3), ("v8k2", 8, 2, 2)]:
        arch, pos = [], 0
        for i in range(V):
            s = 3 if (tag.startswith("v9") and i == 0) else size
            arch.append(dict(name=str(i), size=s, type="regression", position=pos))
            pos += s
        random.seed(1234)
        c = T.Corrupter(nb_observation=17, arch=arch, k_max=k, device=torch.device("cpu"))
        rec[tag + "_sizes"] = np.array([a["size"] for a in arch])
        rec[tag + "_binary_masks"] = c.binary_masks.numpy()
        rec[tag + "_nb_missing_per_run"] = np.array(c.nb_missing_per_run)
        rec[tag + "_nb_corruption_per_k"] = np.array(c.nb_corruption_per_k)
        rec[tag + "_mask_to_use"] = c.mask_to_use.numpy()
        rec[tag + "_nb_run"] = c.nb_run
    for bad in (-1, 3):
        try:
            T.Corrupter(nb_observation=2, arch=[dict(size=1, position=i) for i in range(3)], k_max=bad,
                        device=torch.device("cpu"))
            rec["raises_%d" % bad] = ""
        except Exception as e:  # noqa
            rec["raises_%d" % bad] = str(e)
    np.savez_compressed(os.path.join(OUT, "corrupter_tables.npz"), **rec)
    print("corrupter_tables", {k: v for k, v in rec.items() if k.startswith("raises")})


def layer_tables(M):
    """Known-answer table of the layer-size rule."""
    rows = []
    for (io, z, nin, nout, e) in [(1536, 1536, 4, 4, 512), (1536, 1536, 3, 3, 512), (1536, 128, 2, 2, 512),
                                  (1536, 128, 3, 3, 512), (1536, 100, 4, 2, 512), (48, 8, 3, 3, 16),
                                  (4096, 4096, 4, 4, 512), (192, 64, 2, 3, 64)]:
        m = M.EmbeddingDenoisingAutoencoder(io_size=io, z_size=z, embedding_size=e, nb_input_layer=nin,
                                            nb_output_layer=nout, steep_layer_size=False)
        dims = [(l.in_features, l.out_features) for l in list(m.input_layer) + list(m.output_layer)
                if isinstance(l, torch.nn.Linear)]
        relu = []
        seq = list(m.input_layer) + list(m.output_layer)
        for i, l in enumerate(seq):
            if isinstance(l, torch.nn.Linear):
                relu.append(int(i + 1 < len(seq) and isinstance(seq[i + 1], torch.nn.ReLU)))
        nenc = sum(isinstance(l, torch.nn.Linear) for l in m.input_layer)
        rows.append(dict(kind="embedding", io=io, z=z, nin=nin, nout=nout, steep=0, dims=dims, relu=relu, nenc=nenc,
                         keys=list(m.state_dict().keys())))
    for (io, z, nin, nout, steep) in [(11, 11, 2, 2, 1), (11, 4, 2, 2, 0), (11, 5, 3, 2, 0), (11, 3, 2, 3, 1)]:
        m = M.MixedVariableDenoisingAutoencoder(arch=[], io_size=io, z_size=z, device=torch.device("cpu"),
                                                nb_input_layer=nin, nb_output_layer=nout, steep_layer_size=bool(steep))
        seq = list(m.input_layer) + list(m.output_layer)
        dims = [(l.in_features, l.out_features) for l in seq if isinstance(l, torch.nn.Linear)]
        relu = [int(i + 1 < len(seq) and isinstance(seq[i + 1], torch.nn.ReLU)) for i, l in enumerate(seq)
                if isinstance(l, torch.nn.Linear)]
        nenc = sum(isinstance(l, torch.nn.Linear) for l in m.input_layer)
        rows.append(dict(kind="mixed", io=io, z=z, nin=nin, nout=nout, steep=steep, dims=dims, relu=relu, nenc=nenc,
                         keys=list(m.state_dict().keys())))
    errs = {}
    for label, fn in [("embedding_steep", lambda: M.EmbeddingDenoisingAutoencoder(48, 48, 16, 2, 2, True)),
                      ("embedding_nin1", lambda: M.EmbeddingDenoisingAutoencoder(48, 8, 16, 1, 2, False)),
                      ("embedding_io", lambda: M.EmbeddingDenoisingAutoencoder(50, 8, 16, 2, 2, False))]:
        try:
            fn()
            errs[label] = ""
        except Exception as e:  # noqa
            errs[label] = type(e).__name__ + ":" + str(e)
    import json
    with open(os.path.join(OUT, "layer_tables.json"), "w") as f:
        json.dump(dict(rows=rows, errors=errs), f, indent=1)
    print("layer_tables", len(rows), errs)


if __name__ == "__main__":
    torch.set_num_threads(THREADS)
    os.makedirs(OUT, exist_ok=True)
    M, T, D = import_reference()
    layer_tables(M)
    corrupter_tables(T)
    cats = ["top", "bottom", "shoe"]
    embedding_case(M, T, D, "emb_small", 27493045, n=96, cats=cats, e=16, z=48, nin=2, nout=2, B=8, steps=3,
                   lr=1e-5, wd=1e-4, clip=True)
    embedding_case(M, T, D, "emb_bottleneck", 50493213, n=96, cats=cats, e=16, z=8, nin=3, nout=3, B=12, steps=3,
                   lr=1e-4, wd=1e-2, clip=False)
    embedding_case(M, T, D, "emb_k2", 777, n=96, cats=["a", "b", "c", "d"], e=8, z=32, nin=2, nout=2, B=16, steps=2,
                   lr=1e-3, wd=0.0, clip=True, k_max=2)
    embedding_case(M, T, D, "emb_mid", 4242, n=160, cats=cats, e=64, z=192, nin=3, nout=3, B=32, steps=2,
                   lr=1e-4, wd=1e-2, clip=True)
    abalone_case(M, T, "abalone_k1", 27123045, n=256, z=11, steep=True, B=64, steps=3, lr=5e-5, wd=1e-6, k_max=1)
    abalone_case(M, T, "abalone_k3", 99, n=200, z=4, steep=False, B=50, steps=3, lr=5e-3, wd=1e-6, k_max=3)
