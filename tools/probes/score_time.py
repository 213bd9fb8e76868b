"""Probe: time of one 10 M x 512 sweep (f32 and bf16 catalog) through the public scorer; GB/s against the catalog bytes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import torch
from codae.tool.inference import ComplementarityScorer
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
n = 10_000_000
cat = torch.rand((n, 512), generator=g, device=dev)
q = torch.rand((1, 512), generator=g, device=dev)
for c in (cat, cat.to(torch.bfloat16)):
    sc = ComplementarityScorer(c, 512, metric="sqerr", k=10)
    for _ in range(3):
        sc.topk(q)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        s, i = sc.topk(q)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%s: %.3f ms per sweep, %.0f GB/s, top-3 %s" % (c.dtype, ms, c.numel() * c.element_size() / ms / 1e6, i[0, :3].tolist()))
