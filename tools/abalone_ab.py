"""A/B of the tabular step: per-layer FFMA kernels (22 launches) vs the whole-network kernels (7 launches), abalone.yaml shapes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import torch
from codae.dataset import MixedVariableDataset
from codae.model import MixedVariableDenoisingAutoencoder
from codae.tool import Corrupter, FusedStep
dev = torch.device("cuda", 0)
arch = [dict(name="Sex", size=3, type="classification", position=0)] + [dict(name=str(i), size=1, type="regression", position=3 + i) for i in range(8)]
torch.manual_seed(1)
data = torch.rand(4177, 11); data[:, :3] = torch.nn.functional.one_hot(torch.randint(0, 3, (4177,)), 3).float()
for graph in (False, True):
    for tiny in (False, True):
        ds = MixedVariableDataset.from_arch(arch, data.clone()); ds.to(dev)
        m = MixedVariableDenoisingAutoencoder(arch, 11, 11, dev, 2, 2, True); m.to(dev)
        cor = Corrupter(4177, arch, 1, dev, seed=3)
        fs = FusedStep(m, cor, ds.data, 5e-5, 1e-6, clip=True, tiny_mlp=tiny, use_graph=graph,
                       mixed=dict(arch=arch, weight=[0.4] + [1] * 8, norm_scale=torch.rand(8), norm_min=torch.rand(8), norm_first=3))
        idx = [torch.randint(0, 4177, (64,), device=dev) for _ in range(8)]
        for s in range(20):
            fs.step(idx[s % 8], run=s % 9)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(500):
            fs.step(idx[s % 8], run=0)
        e1.record(); torch.cuda.synchronize()
        print("abalone step graph=%s tiny_mlp=%s: %d launches, %.1f us/step device, %.1f us/step wall"
              % (graph, tiny, fs.kernel_launches, e0.elapsed_time(e1) * 2, (time.perf_counter() - t0) * 2000), flush=True)
