"""CTA-pair (cta_group::2) persistent kernel against the single-CTA persistent kernel: same bits for every role and fused
epilogue (ragged row tiles, a peer CTA whose rows are entirely out of range, ragged column tiles, the bias-gradient column),
then -- with --time -- the three contractions of one polyvore-shaped layer (M = 8192, 4096 x 4096) timed with CUDA events.

    python tools/probes/cta_pair_check.py [--time | --time-only]

Prints one line per case and `PAIR OK` when every case is bit-identical.  tests/test_gpu_variants.py runs it in a child process
under a timeout (a protocol bug in a persistent kernel traps or times out there instead of taking the session with it)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "mui-deepautoencoder_b200"))
import torch  # noqa: E402

from codae import _C as C  # noqa: E402

DEV = torch.device("cuda", 0)
bf = torch.bfloat16


def run_cases(pair):
    C.set_option(DEV, C.OPT_CTA_PAIR, pair)
    out = {}
    torch.manual_seed(57)
    for (M, N, K) in [(4000, 4096, 264), (3900, 4096, 200), (8192, 1067, 264), (8192, 1040, 136)]:
        ldn = (N + 7) // 8 * 8 + 8
        X = torch.randn(M, K).to(DEV, bf)
        W = (torch.randn(N, K) / 16).to(DEV, bf)
        Yb = torch.full((M + 2, ldn), 3.0, device=DEV, dtype=bf)
        Yf = torch.full((M + 2, ldn), 3.0, device=DEV)
        C.linear_fwd(X, W, None, Yb[:M, :N], M, N, K, C.ACT_RELU, C.BF16)
        C.linear_fwd(X, W, None, Yf[:M, :N], M, N, K, C.ACT_NONE, C.BF16)
        out["fwd relu bf16 %dx%dx%d" % (M, N, K)] = Yb
        out["fwd f32 %dx%dx%d" % (M, N, K)] = Yf
    for (M, N, K) in [(4000, 4096, 264), (3900, 4096, 264)]:
        Wd = (torch.randn(K, N) / 64).to(DEV, bf)             # dgrad: dX[M, N] = dY[M, K] . Wd[K, N], masked by A_prev
        dY = torch.randn(M, K).to(DEV, bf)
        A_prev = torch.randn(M, N).to(DEV, bf)
        dXb = torch.full((M + 2, N + 8), 3.0, device=DEV, dtype=bf)
        C.linear_dgrad(dY, Wd, A_prev, dXb[:M, :N], M, K, N, C.BF16)
        out["dgrad %dx%dx%d" % (M, N, K)] = dXb
    for (Mb, Nf, Kf, ld) in [(256, 4096, 4097, 4160), (200, 4000, 4097, 4160), (328, 4096, 2400, 2432)]:
        dYw = torch.randn(Mb, Nf).to(DEV, bf)
        Xa = torch.zeros(Mb, ld, device=DEV, dtype=bf)
        Xa[:, :Kf - 1] = torch.randn(Mb, Kf - 1).to(DEV, bf)
        Xa[:, Kf - 1] = 1
        dW = torch.full((Nf + 2, ld), 7.0, device=DEV)
        slots = C.linear_wgrad_sq_slots(DEV, Mb, Nf, Kf, C.BF16)
        part = torch.zeros(slots, dtype=torch.float64, device=DEV)
        C.linear_wgrad_sq(dYw, Xa[:, :Kf], dW[:Nf, :Kf], Mb, Nf, Kf, C.BF16, part)
        out["wgrad %dx%dx%d" % (Nf, Kf, Mb)] = dW
        out["wgrad sumsq %dx%dx%d" % (Nf, Kf, Mb)] = part.sum().reshape(1)
        # fp64 check of what was stored (the single-CTA path is itself pinned by the oracle tests)
        want = dYw.double().t().mm(Xa[:, :Kf].double())
        err = float((dW[:Nf, :Kf].double() - want).abs().max() / want.abs().max())
        assert err < 1e-5, ("wgrad vs fp64", pair, err)
    torch.cuda.synchronize()
    return out


def time_layer(pair, reps=20):
    C.set_option(DEV, C.OPT_CTA_PAIR, pair)
    torch.manual_seed(5)
    B, n = 8192, 4096
    ld = 4160
    X = torch.zeros(B, ld, device=DEV, dtype=bf)
    X[:, :n] = torch.randn(B, n).to(DEV, bf)
    X[:, n] = 1
    W = torch.zeros(n, ld, device=DEV, dtype=bf)
    W[:, :n + 1] = (torch.randn(n, n + 1) / 64).to(DEV, bf)
    Y = torch.zeros(B, ld, device=DEV, dtype=bf)
    dY = torch.randn(B, n).to(DEV, bf)
    dX = torch.zeros(B, ld, device=DEV, dtype=bf)
    dW = torch.zeros(n, ld, device=DEV)
    slots = C.linear_wgrad_sq_slots(DEV, B, n, n + 1, C.BF16)
    part = torch.zeros(slots, dtype=torch.float64, device=DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    res = {}
    calls = {
        "fwd": lambda: C.linear_fwd(X[:, :n + 1], W[:, :n + 1], None, Y[:, :n], B, n, n + 1, C.ACT_RELU, C.BF16),
        "dgrad": lambda: C.linear_dgrad(dY, W[:, :n], X[:, :n], dX[:, :n], B, n, n, C.BF16),
        "wgrad": lambda: C.linear_wgrad_sq(dY, X[:, :n + 1], dW[:, :n + 1], B, n, n + 1, C.BF16, part),
    }
    for name, fn in calls.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        res[name] = ts[len(ts) // 2]
    return res


def main():
    C.ctx(DEV)
    saved = C.get_option(DEV, C.OPT_CTA_PAIR)
    try:
        bad = 0
        base = pair = {}
        if "--time-only" not in sys.argv:
            base = run_cases(0)
            pair = run_cases(1)
        for k in base:
            a, b = base[k], pair[k]
            if "sumsq" in k:
                same = bool(torch.allclose(a, b, rtol=1e-12, atol=0))
            else:
                same = torch.equal(a.view(torch.int16), b.view(torch.int16)) if a.dtype == bf else torch.equal(a, b)
            d = float((a.double() - b.double()).abs().max())
            print("%-34s %s  max |diff| %.3e" % (k, "same" if same else "DIFFERENT", d), flush=True)
            bad += 0 if same else 1
        if base:
            print("PAIR OK" if bad == 0 else "PAIR MISMATCH (%d cases)" % bad, flush=True)
        if "--time" in sys.argv or "--time-only" in sys.argv:
            flop = 2.0 * 8192 * 4096 * 4096
            for rnd in range(2):
                for on in (0, 1):
                    t = time_layer(on)
                    print("round %d pair=%d  " % (rnd, on) + "  ".join("%s %.4f ms (%.0f TFLOP/s)" % (k, v, flop / v / 1e9) for k, v in t.items()),
                          flush=True)
        return 0 if bad == 0 else 1
    finally:
        C.set_option(DEV, C.OPT_CTA_PAIR, saved)


if __name__ == "__main__":
    sys.exit(main())
