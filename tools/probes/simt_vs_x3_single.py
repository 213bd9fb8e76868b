"""Probe: ONE training step of GB rows with the FFMA engine vs the fp32-parity tensor-core engine vs torch (fp64 on the GPU):
which engine's post-step weights are off at GB = 512, and where?"""
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
    W0 = [l.weight.detach().clone().double().to(dev) for l in m.linears()]
    b0 = [l.bias.detach().clone().double().to(dev) for l in m.linears()]
    m.set_compute_dtype(dtype); m.to(dev); ds.to(dev)
    cor = Corrupter(N, ds.arch, 1, dev, seed=77)
    fs = FusedStep(m, cor, ds.data, lr=1e-3, weight_decay=1e-4, clip=True)
    rng = np.random.RandomState(5)
    g = rng.permutation(1024)[:GB]
    idx = torch.as_tensor(g, dtype=torch.int64, device=dev)
    fs.step(idx, global_batch=GB)
    torch.cuda.synchronize()
    # torch fp64 reference of the same step
    _, fmask = cor.get_masks(idx, 0)
    x = ds.data[idx].double()
    cx = x * fmask.double()
    Ws = [w.clone().requires_grad_(True) for w in W0]
    bs = [b.clone().requires_grad_(True) for b in b0]
    a = cx
    for l, (w, b) in enumerate(zip(Ws, bs)):
        a = a @ w.t() + b
        if m.relu[l]:
            a = torch.relu(a)
    loss = ((x - a) ** 2).mean()
    loss.backward()
    out = []
    for l in range(len(Ws)):
        gw = Ws[l].grad + 1e-4 * W0[l]
        # Adam step 1: m = (1-b1) g, v = (1-b2) g^2 ; update = lr * m/(1-b1) / (sqrt(v/(1-b2)) + eps) = lr * g / (|g| + eps)
        upd = 1e-3 * gw / (gw.abs() + 1e-8)
        want = W0[l] - upd
        got = m.weight_view(m.flat, l).double()
        gg = m.weight_view(fs.gflat, l).double()
        out.append((float((got - want).abs().max()), int(((got - want).abs() > 1e-6).sum()), got.numel(),
                    float((gg - Ws[l].grad).abs().max() / Ws[l].grad.abs().max())))
    return out, float(loss), fs.last_loss(GB)

for GB in (128, 512):
    for dtype in ("fp32_simt", "fp32"):
        out, l64, lk = run(dtype, GB)
        print("GB=%d %-9s loss %.8f (fp64 %.8f)" % (GB, dtype, lk, l64))
        for l, (e, n, tot, ge) in enumerate(out):
            print("    layer %d: |w - w_ref|max %.3e, %d of %d elements off by > 1e-6 ; |g - g_ref|max/|g|max %.2e" % (l, e, n, tot, ge))
