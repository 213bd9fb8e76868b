# Diagnostic (not a pytest): where does the TMA bulk-store epilogue differ from the per-thread store epilogue?
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mui-deepautoencoder_b200"))
import torch
from codae import _C as C
DEV = torch.device("cuda", 0)
ru = lambda x, m: (x + m - 1) // m * m
for B, o, i in [(128, 192, 328), (128, 1536, 1536)]:
    torch.manual_seed(23)
    bf = torch.bfloat16
    K = ru(i, 8) + 1; ld = ru(K, 64)
    dY = torch.randn(B, ru(o, 8)).to(DEV, bf)
    X = torch.zeros(B, ld, device=DEV, dtype=bf); X[:, :i] = torch.randn(B, i).to(DEV, bf); X[:, ru(i, 8)] = 1
    out = {}
    for on in (0, 1):
        C.set_option(DEV, C.OPT_TMA_STORE, on)
        dW = torch.full((o + 3, ld), 5.0, device=DEV)
        C.linear_wgrad(dY[:, :o], X[:, :K], dW[:o, :K], None, B, o, K, C.BF16)
        torch.cuda.synchronize()
        out[on] = dW.cpu()
    C.set_option(DEV, C.OPT_TMA_STORE, 0)
    want = dY[:, :o].double().cpu().t().mm(X[:, :K].double().cpu())
    for on in (0, 1):
        e = (out[on][:o, :K].double() - want).abs().max() / want.abs().max()
        print("shape", (B, o, i), "tma" if on else "ref", "max rel err vs fp64 %.3e" % float(e))
    d = (out[0] != out[1]).nonzero()
    print("  differing elements:", d.shape[0], "of", out[0].numel())
    if d.shape[0]:
        r, c = d[:, 0], d[:, 1]
        print("  rows min/max", int(r.min()), int(r.max()), "cols min/max", int(c.min()), int(c.max()))
        print("  row%8 hist", torch.bincount(r % 8, minlength=8).tolist())
        print("  (col%32)//4 hist", torch.bincount((c % 32) // 4, minlength=8).tolist())
        print("  col//32 hist (first 12)", torch.bincount(c // 32)[:12].tolist())
        print("  in padding (col>=K):", int((c >= K).sum()), " in guard rows:", int((r >= o).sum()))
        for t in range(min(8, d.shape[0])):
            rr, cc = int(r[t]), int(c[t])
            print("   [%d,%d] tma %.6f ref %.6f want %.6f" % (rr, cc, float(out[1][rr, cc]), float(out[0][rr, cc]), float(want[rr, cc]) if rr < o and cc < K else float("nan")))
        # is the TMA result a permutation of 16-byte chunks inside 128-byte rows?
        rr = int(r[0]); base = (int(c[0]) // 32) * 32
        print("   ref row", rr, "cols", base, "..", out[0][rr, base:base + 32].tolist())
        print("   tma row", rr, "cols", base, "..", out[1][rr, base:base + 32].tolist())
