"""Diagnostic (not a pytest file): codae_linear_chain (one persistent launch for a chain of Linear layers) against the
per-layer kernels on the same operands -- forward chains and input-gradient chains at the shipped layer sizes -- and its
time per layer inside a CUDA graph.  Prints one line per case and "CHAIN OK" at the end; exits non-zero on a mismatch.
Run under a timeout: the kernel spins on global counters (bounded) and on cluster barriers (not bounded)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import torch
from codae import _C as C

dev = torch.device("cuda", 0)
bf = torch.bfloat16
ru = lambda x, m: (x + m - 1) // m * m
torch.manual_seed(5)
ws = C.linear_chain_workspace(dev)
bad = 0


def act_buf(B, width, dtype):
    t = torch.zeros(B, ru(ru(width, 8) + 1, 64), device=dev, dtype=dtype)
    t[:, ru(width, 8)] = 1                      # the constant-1 column of the augmented layout
    return t


def forward_case(B, widths, time_it=False):
    global bad
    L = len(widths) - 1
    Ws = []
    for l in range(L):
        i, o = widths[l], widths[l + 1]
        w = torch.zeros(o, ru(ru(i, 8) + 1, 64), device=dev, dtype=bf)
        w[:, :i] = (torch.randn(o, i, device=dev) / i ** 0.5).to(bf)
        w[:, ru(i, 8)] = (torch.randn(o, device=dev) * 0.1).to(bf)      # bias column
        Ws.append(w)
    x = act_buf(B, widths[0], bf)
    x[:, :widths[0]] = torch.randn(B, widths[0], device=dev).to(bf)
    ref = [x] + [act_buf(B, widths[l + 1], torch.float32 if l == L - 1 else bf) for l in range(L)]
    got = [x] + [act_buf(B, widths[l + 1], torch.float32 if l == L - 1 else bf) for l in range(L)]
    relu = [l != L - 1 for l in range(L)]
    for l in range(L):
        C.linear_fwd(ref[l], Ws[l][:, :ru(widths[l], 8) + 1], None, ref[l + 1], B, widths[l + 1], ru(widths[l], 8) + 1,
                     C.ACT_RELU if relu[l] else C.ACT_NONE, C.BF16)
    layers = [C.chain_layer(got[l], Ws[l][:, :ru(widths[l], 8) + 1], True, got[l + 1], widths[l + 1], ru(widths[l], 8) + 1,
                            C.ACT_RELU if relu[l] else C.ACT_NONE) for l in range(L)]
    C.linear_chain(layers, B, ws)
    torch.cuda.synchronize()
    worst, equal = 0.0, True
    for l in range(1, L + 1):
        a, b = got[l].float(), ref[l].float()
        worst = max(worst, float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)))
        equal = equal and torch.equal(a, b)
    ok = worst < 1e-2
    bad += 0 if ok else 1
    msg = "fwd  B=%3d widths=%s: max rel diff vs per-layer %.3e, bit-identical=%s %s" % (B, widths, worst, equal, "ok" if ok else "MISMATCH")
    if time_it and ok:
        def per_layer():
            for l in range(L):
                C.linear_fwd(ref[l], Ws[l][:, :ru(widths[l], 8) + 1], None, ref[l + 1], B, widths[l + 1], ru(widths[l], 8) + 1,
                             C.ACT_RELU if relu[l] else C.ACT_NONE, C.BF16)
        def chain():
            C.linear_chain(layers, B, ws)
        times = []
        for fn in (per_layer, chain):
            fn(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(10):
                    fn()
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 100.0)           # us per pass
        msg += " | us per forward pass: per-layer %.1f, chain %.1f" % (times[0], times[1])
    print(msg, flush=True)


def dgrad_case(B, widths):
    """Input-gradient chain: g_{l-1} = (g_l . W_l) * (a_{l-1} > 0), from the last layer down to layer 1."""
    global bad
    L = len(widths) - 1
    Ws = [(torch.randn(widths[l + 1], ru(widths[l], 8), device=dev) / widths[l + 1] ** 0.5).to(bf) for l in range(L)]
    acts = [torch.randn(B, ru(widths[l], 8), device=dev).to(bf) for l in range(L + 1)]           # stored activations (mask source)
    g_top = torch.randn(B, ru(widths[L], 8), device=dev).to(bf)
    ref = {L: g_top}
    got = {L: g_top}
    for l in range(L - 1, 0, -1):
        ref[l] = torch.zeros(B, ru(widths[l], 8), device=dev, dtype=bf)
        got[l] = torch.zeros(B, ru(widths[l], 8), device=dev, dtype=bf)
    for l in range(L - 1, 0, -1):                 # layer index l+1 in FusedStep's numbering: dX of Linear l (in widths[l], out widths[l+1])
        C.linear_dgrad(ref[l + 1], Ws[l][:, :widths[l]], acts[l], ref[l], B, widths[l + 1], widths[l], C.BF16)
    layers = [C.chain_layer(got[l + 1], Ws[l][:, :widths[l]], False, got[l], widths[l], widths[l + 1], C.ACT_NONE, acts[l])
              for l in range(L - 1, 0, -1)]
    C.linear_chain(layers, B, ws)
    torch.cuda.synchronize()
    worst, equal = 0.0, True
    for l in range(1, L):
        a, b = got[l].float(), ref[l].float()
        worst = max(worst, float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)))
        equal = equal and torch.equal(a, b)
    ok = worst < 1e-2
    bad += 0 if ok else 1
    print("dgrad B=%3d widths=%s: max rel diff vs per-layer %.3e, bit-identical=%s %s" % (B, widths, worst, equal, "ok" if ok else "MISMATCH"), flush=True)


forward_case(128, [1536] * 3)
forward_case(128, [1536] * 11, time_it=True)
forward_case(32, [1536] * 9, time_it=True)
forward_case(100, [1536, 1536, 1067, 598, 128, 597, 1066, 1535, 1536])
forward_case(128, [192, 328, 64, 192])
dgrad_case(128, [1536] * 11)
dgrad_case(64, [1536, 1536, 1064, 600, 128, 600, 1064, 1536, 1536])
print("CHAIN OK" if bad == 0 else "CHAIN MISMATCHES: %d" % bad)
sys.exit(0 if bad == 0 else 1)
