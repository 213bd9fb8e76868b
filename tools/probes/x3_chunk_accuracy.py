"""Probe: accuracy of the fp32-parity contraction vs the TMEM accumulation chunk (CODAE_X3_CHUNK_KB) -- error against an fp64
product relative to max |result| for the embedding.yaml shapes, next to the FFMA engine and torch's own fp32 GEMM on the GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import torch
from codae import _C
dev = torch.device("cuda", 0)
torch.manual_seed(1)
def planes(t, pitch):
    r, c = t.shape
    src = torch.zeros(r, pitch, device=dev); src[:, :c] = t
    out = _C.new_x3((r, pitch), dev); _C.split_x3(src, out); return out[:, :, :c]
def rel(a, b): return float((a.double() - b).abs().max() / b.abs().max())
for (M, N, K) in ((128, 1536, 1536), (128, 1536, 4096), (2048, 1536, 1536)):
    X = torch.randn(M, K, device=dev).abs(); W = torch.randn(N, K, device=dev) / K ** 0.5; dY = torch.randn(M, N, device=dev) * 1e-3
    Y = torch.zeros(M, N, device=dev); dW = torch.zeros(N, K, device=dev); dX = torch.zeros(M, K, device=dev)
    Xp, Wp, dYp = planes(X, K + 8), planes(W, K + 8), planes(dY, N + 8)
    _C.linear_fwd(Xp, Wp, None, Y, M, N, K, _C.ACT_NONE, _C.F32X3)
    _C.linear_wgrad(dYp, Xp, dW, None, M, N, K, _C.F32X3)
    _C.linear_dgrad(dYp, Wp, None, dX, M, N, K, _C.F32X3)
    r64 = (X.double() @ W.double().t(), dY.double().t() @ X.double(), dY.double() @ W.double())
    Ys = torch.zeros(M, N, device=dev); dWs = torch.zeros(N, K, device=dev); dXs = torch.zeros(M, K, device=dev)
    _C.linear_fwd(X, W, None, Ys, M, N, K, _C.ACT_NONE, _C.F32)
    _C.linear_wgrad(dY, X, dWs, None, M, N, K, _C.F32)
    _C.linear_dgrad(dY, W, None, dXs, M, N, K, _C.F32)
    torch.backends.cuda.matmul.allow_tf32 = False
    print("chunk_kb=%s M=%d N=%d K=%d  x3 fwd/wgrad/dgrad %.2e %.2e %.2e | FFMA %.2e %.2e %.2e | torch fp32 %.2e %.2e %.2e" % (
        os.environ.get("CODAE_X3_CHUNK_KB", "8"), M, N, K, rel(Y, r64[0]), rel(dW, r64[1]), rel(dX, r64[2]),
        rel(Ys, r64[0]), rel(dWs, r64[1]), rel(dXs, r64[2]), rel(X @ W.t(), r64[0]), rel(dY.t() @ X, r64[1]), rel(dY @ W, r64[2])))
