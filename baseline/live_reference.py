"""Live-reference CPU arm: times the UNMODIFIED reference classes (installed into baseline/_ref by baseline/install_ref.py)
on the host cores, replaying the reference's own loop body (script/train_dae_on_embedding.py:194-223) around them:

    corrupter.get_masks -> model.corrupt -> model() -> MSELoss(mean) -> zero_grad / backward -> [clip_grad_norm_] -> Adam.step
    -> MSELoss(none).cpu().numpy() sums (full / partial monitors)

and the reference's scoring op (codae/tool/metering.py:67-69: cosine_similarity of the prediction against the catalog) + topk.
Test / benchmark infrastructure only: imported by bench.py (`--impl reference`, `cpu_baseline`) and never by the product.
The scripts themselves cannot run here (they import matplotlib and need real data files), hence the
replay; matplotlib is stubbed exactly like oracle/gen_golden.py does.
"""
import os
import random
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "codae", "model", "embedding_denoising_autoencoder.py"))


def import_reference():
    """(codae.model, codae.tool, codae.dataset) of the reference, from baseline/_ref.  The repo's own package is also called
    `codae`: the reference is imported under a private sys.modules view and the previous entries are restored afterwards."""
    saved = {k: v for k, v in sys.modules.items() if k == "codae" or k.startswith("codae.")}
    for k in saved:
        del sys.modules[k]
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    had_mpl = "matplotlib" in sys.modules
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.path.insert(0, REF)
    try:
        import codae.model as M
        import codae.tool as T
        import codae.dataset as D
        assert os.path.abspath(M.__file__).startswith(REF), M.__file__
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "codae" or k.startswith("codae.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        if not had_mpl:
            sys.modules.pop("matplotlib", None)
            sys.modules.pop("matplotlib.pyplot", None)
    return M, T, D


def train_steps(w, rows, B, steps, warmup, seed=0, threads=None):
    """w: bench.py workload dict; rows: [n, io] fp32 CPU tensor of synthetic (already scaled) dataset rows.
    Returns dict(seconds, steps, B, cores, loss)."""
    M, T, D = import_reference()
    cores = threads or (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    io, E = w["S"] * w["E"], w["E"]
    n = rows.shape[0]
    arch = [dict(name=str(i), size=E, type="regression", position=i * E) for i in range(w["S"])]
    dev = torch.device("cpu")
    corrupter = T.Corrupter(nb_observation=n, arch=arch, k_max=w["k_max"], device=dev)
    model = M.EmbeddingDenoisingAutoencoder(io_size=io, z_size=w["z"], embedding_size=E, nb_input_layer=w["nin"],
                                           nb_output_layer=w["nout"], steep_layer_size=False)
    model.to(dev)
    optimizer = torch.optim.Adam(model.parameters(), lr=w["lr"], weight_decay=w["wd"])
    mean_criterion = torch.nn.MSELoss(reduction="mean")
    full_criterion = torch.nn.MSELoss(reduction="none")
    rng = np.random.RandomState(seed)
    ftl = ptl = 0.0
    t0 = None
    loss = None
    for s in range(warmup + steps):
        if s == warmup:
            t0 = time.perf_counter()
        batch_indices = tuple(int(i) for i in rng.randint(0, n, size=B))
        input_data = torch.stack([rows[i] for i in batch_indices])         # collate_embedding (data_tool.py:96-103)
        masks, fmask = corrupter.get_masks(batch_indices, 0)
        c_input_data = model.corrupt(input_data=input_data, mask=fmask)
        output_data = model(c_input_data)
        loss = mean_criterion(input_data, output_data)
        optimizer.zero_grad()
        loss.backward()
        if w["clip"]:
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        optimizer.step()
        full = full_criterion(input_data, output_data).cpu().detach().numpy()
        ftl += np.sum(full)
        ptl += np.sum((1 - fmask.cpu().numpy()) * full)
    dt = time.perf_counter() - t0
    return dict(seconds=dt, steps=steps, B=B, cores=cores, loss=float(loss), ftl=float(ftl), ptl=float(ptl))


def scoring_sweeps(E, rows, reps=3, k=10, seed=0, threads=None):
    """The reference's scoring op over a [rows, E] catalog: cosine_similarity(catalog, prediction) (metering.py:67-69) + topk."""
    cores = threads or (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(seed)
    cat = torch.rand(rows, E, generator=g)
    q = torch.rand(E, generator=g)
    t0 = time.perf_counter()
    for _ in range(reps):
        s = torch.nn.functional.cosine_similarity(cat, q.reshape(1, -1))
        torch.topk(s, k)
    return dict(seconds=time.perf_counter() - t0, scores=rows * reps, cores=cores)
