"""Diagnostic probe (not a pytest file): runs the tcgen05 engine on structured inputs and prints error maps, so
that one GPU call tells which of {descriptor, swizzle, majorness, epilogue} is wrong if the parity test fails."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import torch
from codae import _C

dev = torch.device("cuda", 0)
bf = torch.bfloat16


def report(tag, got, want):
    d = (got.double() - want.double()).abs()
    scale = float(want.double().abs().max()) + 1e-30
    bad = d > 1e-3 * scale
    print("%-34s max_rel=%.3e bad=%d/%d" % (tag, float(d.max()) / scale, int(bad.sum()), bad.numel()), flush=True)
    if bad.any():
        r, c = torch.nonzero(bad)[0].tolist()
        rows_bad = torch.nonzero(bad.any(1)).flatten()[:8].tolist()
        cols_bad = torch.nonzero(bad.any(0)).flatten()[:8].tolist()
        print("   first bad (%d,%d): got %.5f want %.5f; bad rows %s cols %s" % (r, c, float(got[r, c]), float(want[r, c]), rows_bad, cols_bad), flush=True)


for (M, N, K) in [(128, 64, 64), (128, 128, 256), (256, 256, 512), (128, 1536, 1536), (100, 200, 136)]:
    torch.manual_seed(0)
    X = torch.randint(-2, 3, (M, K)).float()
    W = torch.randint(-2, 3, (N, K)).float()
    dY = torch.randint(-2, 3, (M, N)).float()
    Xd, Wd, dYd = X.to(dev, bf), W.to(dev, bf), dY.to(dev, bf)
    print("== M=%d N=%d K=%d" % (M, N, K), flush=True)
    Y = torch.zeros(M, N, device=dev)
    _C.linear_fwd(Xd, Wd, None, Y, M, N, K, _C.ACT_NONE, _C.BF16); torch.cuda.synchronize()
    report("fwd  (A K-major, B K-major)", Y.cpu(), X @ W.t())
    dX = torch.zeros(M, K, device=dev)
    _C.linear_dgrad(dYd, Wd, None, dX, M, N, K, _C.BF16); torch.cuda.synchronize()
    report("dgrad(A K-major, B MN-major)", dX.cpu(), dY @ W)
    dW = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
    _C.linear_wgrad(dYd, Xd, dW, db, M, N, K, _C.BF16); torch.cuda.synchronize()
    report("wgrad(A MN-major, B MN-major)", dW.cpu(), dY.t() @ X)
print("probe done", flush=True)
