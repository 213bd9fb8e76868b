"""Scoring-only driver for ncu captures: one fp32 and one bf16 sweep of a synthetic catalog through the public scorer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import torch
from codae.tool.inference import ComplementarityScorer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
cat = torch.rand((n, 512), generator=g, device=dev)
q = torch.rand((1, 512), generator=g, device=dev)
for c in (cat, cat.to(torch.bfloat16)):
    sc = ComplementarityScorer(c, 512, metric="sqerr", k=10)
    for _ in range(3):
        s, i = sc.topk(q)
    torch.cuda.synchronize()
    print(c.dtype, i[0].tolist())
