"""Probe: one training step, fp32-parity engine vs FFMA engine vs fp64 replay -- error of every activation and every weight
gradient, layer by layer (relative to the max of the fp64 tensor), to see where the x3 engine's 2e-6 gradient error comes from."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import numpy as np, torch
from codae import _C
from codae.dataset import ConcatenatedEmbeddingDataset
from codae.model import EmbeddingDenoisingAutoencoder
from codae.tool import Corrupter, FusedStep
dev = torch.device("cuda", 0)

def rel(a, b): return float((a.double() - b).abs().max() / b.abs().max())

def run(dtype, GB):
    torch.manual_seed(3)
    S, E, N = 3, 128, 1024
    cats = [torch.randn(N, E).abs() for _ in range(S)]
    ds = ConcatenatedEmbeddingDataset.from_tensors(cats)
    m = EmbeddingDenoisingAutoencoder(S * E, S * E, E, 2, 2, False)
    m.set_compute_dtype(dtype); m.to(dev); ds.to(dev)
    cor = Corrupter(N, ds.arch, 1, dev, seed=77)
    fs = FusedStep(m, cor, ds.data, lr=0.0, weight_decay=0.0, clip=False)      # lr 0: weights stay, buffers can be read after the step
    L = len(m.dims)
    idx = torch.as_tensor(np.random.RandomState(5).permutation(1024)[:GB], dtype=torch.int64, device=dev)
    fs.step(idx, global_batch=GB)
    torch.cuda.synchronize()
    Ws = [m.weight_view(m.flat, l).double().clone().requires_grad_(True) for l in range(L)]
    bs = [m.bias_view(m.flat, l).double().clone().requires_grad_(True) for l in range(L)]
    _, fmask = cor.get_masks(idx, 0)
    x = ds.data[idx].double()
    a = x * fmask.double()
    acts64 = [a]
    for l in range(L):
        a = a @ Ws[l].t() + bs[l]
        if m.relu[l]:
            a = torch.relu(a)
        a.retain_grad()
        acts64.append(a)
    loss = ((x - a) ** 2).mean()
    loss.backward()
    b = fs._bufs[GB]
    out = []
    for l in range(L + 1):
        t = b["acts"][l]
        w = m.dims[l][0] if l < L else m.dims[L - 1][1]
        t = _C.x3_to_f32(t) if t.dim() == 3 else t
        out.append("a%d %.1e" % (l, rel(t[:, :w], acts64[l].detach())))
    gout = []
    for l in range(L):
        gout.append("dW%d %.1e db%d %.1e" % (l, rel(m.weight_view(fs.gflat, l), Ws[l].grad), l, rel(m.bias_view(fs.gflat, l), bs[l].grad)))
    # dL/d(out_l) for the last three layers still sit in the rotating buffers
    gb = [b["g0"], b["g1"], b["g2"]]
    dout = []
    for l in range(min(3, L)):
        t = gb[l % 3]
        t = _C.x3_to_f32(t) if t.dim() == 3 else t
        want = acts64[l + 1].grad
        if m.relu[l]:
            want = want * (acts64[l + 1] > 0)
        dout.append("g%d %.1e" % (l, rel(t[:, :m.dims[l][1]], want)))
    print("GB=%d %-9s | %s\n        %s\n        %s" % (GB, dtype, "  ".join(out), "  ".join(gout), "  ".join(dout)))

for GB in (128,):
    for dtype in ("fp32_simt", "fp32"):
        run(dtype, GB)
