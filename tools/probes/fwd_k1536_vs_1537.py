"""Probe: what the bias column of the augmented contraction costs a small-batch forward launch -- K = 1537 is 25 k-blocks (5 splits of
5), K = 1536 is 24 (6 splits of 4).  50 back-to-back launches in a CUDA graph, both tensor-core engines."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import torch
from codae import _C
dev = torch.device("cuda", 0)
M, N, ld = 128, 1536, 1600
def x3(rows):
    t = _C.new_x3((rows, ld), dev); _C.split_x3(torch.randn(rows, ld, device=dev) / 40, t); return t
X, W, Y = x3(M), x3(N), x3(M)
Xb, Wb = X[0].clone(), W[0].clone(); Yb = torch.zeros(M, ld, device=dev, dtype=torch.bfloat16)
for K in (1537, 1536):
    for tag, fn in (("x3", lambda: _C.linear_fwd(X[:, :, :K], W[:, :, :K], None, Y[:, :, :N], M, N, K, _C.ACT_RELU, _C.F32X3)),
                    ("bf16", lambda: _C.linear_fwd(Xb[:, :K], Wb[:, :K], None, Yb[:, :N], M, N, K, _C.ACT_RELU, _C.BF16))):
        fn(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(50):
                fn()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        print("K=%d %-4s %.2f us per launch" % (K, tag, e0.elapsed_time(e1) * 1e3 / 50), flush=True)
