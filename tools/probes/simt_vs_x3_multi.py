"""Probe: five training steps of GB rows -- FFMA engine and fp32-parity tensor-core engine against torch fp64 (autograd + clip +
torch.optim.Adam) on the same batches: per step, the gradient error of each engine (its own weights differ after step 1, so the
reference is re-run from the engine's weights each step: like-with-like)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import numpy as np, torch
from codae.dataset import ConcatenatedEmbeddingDataset
from codae.model import EmbeddingDenoisingAutoencoder
from codae.tool import Corrupter, FusedStep
dev = torch.device("cuda", 0)

def run(dtype, GB):
    torch.manual_seed(3)
    S, E, N = 3, 128, 1024
    cats = [torch.randn(N, E).abs() for _ in range(S)]
    ds = ConcatenatedEmbeddingDataset.from_tensors(cats)
    m = EmbeddingDenoisingAutoencoder(S * E, S * E, E, 2, 2, False)
    m.set_compute_dtype(dtype); m.to(dev); ds.to(dev)
    cor = Corrupter(N, ds.arch, 1, dev, seed=77)
    fs = FusedStep(m, cor, ds.data, lr=1e-3, weight_decay=1e-4, clip=True)
    rng = np.random.RandomState(5)
    batches = [rng.permutation(1024)[:GB] for _ in range(4)]
    L = len(m.dims)
    for s, g in enumerate(batches):
        idx = torch.as_tensor(g, dtype=torch.int64, device=dev)
        Ws = [m.weight_view(m.flat, l).double().clone().requires_grad_(True) for l in range(L)]
        bs = [m.bias_view(m.flat, l).double().clone().requires_grad_(True) for l in range(L)]
        fs.step(idx, global_batch=GB)
        torch.cuda.synchronize()
        _, fmask = cor.get_masks(idx, 0)
        x = ds.data[idx].double()
        a = x * fmask.double()
        pre_min = 1e9
        for l in range(L):
            a = a @ Ws[l].t() + bs[l]
            if m.relu[l]:
                pre_min = min(pre_min, float(a.abs().min()))
                a = torch.relu(a)
        loss = ((x - a) ** 2).mean()
        loss.backward()
        errs = []
        for l in range(L):
            gg = m.weight_view(fs.gflat, l).double()
            errs.append(float((gg - Ws[l].grad).abs().max() / Ws[l].grad.abs().max()))
        print("  GB=%d %-9s step %d: loss %.8f vs fp64 %.8f ; per-layer |g - g_ref|max/|g|max: %s ; smallest |pre-activation| %.1e" % (
            GB, dtype, s, fs.last_loss(GB), float(loss.detach()), " ".join("%.1e" % e for e in errs), pre_min))

for GB in (128, 512):
    for dtype in ("fp32_simt", "fp32"):
        run(dtype, GB)
