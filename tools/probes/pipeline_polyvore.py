"""Probe: polyvore-shaped steps (10 x 4096^2, B = 8192) through FusedStep.train_steps (multi-step CUDA graphs).  This was the harness
of the pipelining experiment of visit 8 (optimizer update of step s beside the forward pass of step s + 1; CODAE_PIPELINE selected
it): 7.31 ms/step without it, a trap in the bounded gate wait with it -- the variant is removed, the harness still times the graphs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import torch
from codae.dataset import ConcatenatedEmbeddingDataset
from codae.model import EmbeddingDenoisingAutoencoder
from codae.tool import Corrupter, FusedStep
DEV = torch.device("cuda", 0)
S, E, N, B = 8, 512, 65536, 8192
torch.manual_seed(5)
data = torch.rand(N, S * E, device=DEV)
ds = ConcatenatedEmbeddingDataset.__new__(ConcatenatedEmbeddingDataset)
ds.data, ds.nb_observation, ds.embedding_size, ds.nb_used_category = data, N, E, S
ds.arch = [dict(name=str(i), size=E, type="regression", position=i * E) for i in range(S)]
model = EmbeddingDenoisingAutoencoder(S * E, S * E, E, 4, 4, False)
model.set_compute_dtype("bf16"); model.to(DEV)
cor = Corrupter(N, ds.arch, 2, DEV, seed=1)
fs = FusedStep(model, cor, data, lr=1e-5, weight_decay=1e-4, clip=True, use_graph=False)
rows = torch.randint(0, N, (17, B), device=DEV)
fs.train_steps(rows, graph_steps=8); torch.cuda.synchronize()          # eager + 2 graphs of 8
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fs.train_steps(rows[:16], graph_steps=8); e1.record(); torch.cuda.synchronize()
print("CODAE_PIPELINE=%s: %.3f ms/step over 16 steps (2 graphs of 8), loss %.6f" % (os.environ.get("CODAE_PIPELINE", "1"), e0.elapsed_time(e1) / 16, fs.last_loss(B)))
