"""Diagnostic (not a pytest file): phase timeline of the fp32-parity (CODAE_F32X3) tcgen05 GEMM at the embedding.yaml shapes,
from %globaltimer stamps of CTA (0,0,0) (debug hook codae_debug_set_trace), next to the bf16 engine on the same shapes."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import torch
from codae import _C

dev = torch.device("cuda", 0)
bf = torch.bfloat16
lib = _C.lib()
lib.codae_debug_set_trace.argtypes = [ctypes.c_void_p]
names = ["entry", "prologue", "pdl_wait", "first_operands", "mma_issued", "acc_complete", "staged", "cluster_bar1", "stored", "cluster_bar2", "exit_cta0", "exit_last"]
M, N, K = 128, 1536, 1537
ld = 1600
R = 12
buf = torch.zeros(R * 16, dtype=torch.int64, device=dev)


def x3(rows, fill=True):
    t = _C.new_x3((rows, ld), dev)
    if fill:
        src = torch.randn(rows, ld, device=dev) / 40
        _C.split_x3(src, t)
    return t


X, W, dY = x3(M), x3(N), x3(M)
Y, dX = x3(M, False), x3(M, False)
Yf = torch.zeros(M, ld, device=dev)
dW = torch.zeros(N, ld, device=dev)
Xb, Wb, dYb = X[0].clone(), W[0].clone(), dY[0].clone()
Yb, dXb = torch.zeros(M, ld, device=dev, dtype=bf), torch.zeros(M, ld, device=dev, dtype=bf)


def run(tag, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    buf.zero_()
    for i in range(R):
        lib.codae_debug_set_trace(ctypes.c_void_p(buf.data_ptr() + 128 * i))
        fn()
    lib.codae_debug_set_trace(None)
    torch.cuda.synchronize()
    t = buf.cpu().view(R, 16)
    last = t[R - 1]
    rel = [(names[j], int(last[j] - last[0])) for j in range(12) if int(last[j]) != 0]
    gaps = [int(t[i + 1][0] - t[i][0]) for i in range(R - 1)]
    print("%-12s entry-to-entry ns: %s" % (tag, gaps[-5:]), flush=True)
    print("        phases (ns since entry): " + "  ".join("%s=%d" % (n, v) for n, v in rel), flush=True)


run("x3 fwd", lambda: _C.linear_fwd(X[:, :, :K], W[:, :, :K], None, Y[:, :, :N], M, N, K, _C.ACT_RELU, _C.F32X3))
run("x3 fwd f32", lambda: _C.linear_fwd(X[:, :, :K], W[:, :, :K], None, Yf[:, :N], M, N, K, _C.ACT_NONE, _C.F32X3))
run("x3 dgrad", lambda: _C.linear_dgrad(dY[:, :, :N], W[:, :, :1536], X[:, :, :1536], dX[:, :, :1536], M, N, 1536, _C.F32X3))
run("x3 dgrad nomask", lambda: _C.linear_dgrad(dY[:, :, :N], W[:, :, :1536], None, dX[:, :, :1536], M, N, 1536, _C.F32X3))
run("x3 wgrad", lambda: _C.linear_wgrad(dY[:, :, :N], X[:, :, :K], dW[:, :K], None, M, N, K, _C.F32X3))
run("bf16 fwd", lambda: _C.linear_fwd(Xb[:, :K], Wb[:, :K], None, Yb[:, :N], M, N, K, _C.ACT_RELU, _C.BF16))
run("bf16 dgrad", lambda: _C.linear_dgrad(dYb[:, :N], Wb[:, :1536], Xb[:, :1536], dXb[:, :1536], M, N, 1536, _C.BF16))
run("bf16 wgrad", lambda: _C.linear_wgrad(dYb[:, :N], Xb[:, :K], dW[:, :K], None, M, N, K, _C.BF16))

for cta in (0, 143, 144, 147, 148, 155):
    lib.codae_debug_set_trace_cta(cta)
    run("x3 wgrad cta %d" % cta, lambda: _C.linear_wgrad(dY[:, :, :N], X[:, :, :K], dW[:, :K], None, M, N, K, _C.F32X3))
lib.codae_debug_set_trace_cta(0)

# graph-captured repetitions: GPU time per launch without host gaps
for tag, fn in (("x3 fwd", lambda: _C.linear_fwd(X[:, :, :K], W[:, :, :K], None, Y[:, :, :N], M, N, K, _C.ACT_RELU, _C.F32X3)),
                ("x3 dgrad", lambda: _C.linear_dgrad(dY[:, :, :N], W[:, :, :1536], X[:, :, :1536], dX[:, :, :1536], M, N, 1536, _C.F32X3)),
                ("x3 wgrad", lambda: _C.linear_wgrad(dY[:, :, :N], X[:, :, :K], dW[:, :K], None, M, N, K, _C.F32X3)),
                ("bf16 fwd", lambda: _C.linear_fwd(Xb[:, :K], Wb[:, :K], None, Yb[:, :N], M, N, K, _C.ACT_RELU, _C.BF16)),
                ("bf16 dgrad", lambda: _C.linear_dgrad(dYb[:, :N], Wb[:, :1536], Xb[:, :1536], dXb[:, :1536], M, N, 1536, _C.BF16))):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(50):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print("%-12s %.2f us per launch (50 back-to-back in a graph, same weights: L2-warm)" % (tag, e0.elapsed_time(e1) * 1e3 / 50), flush=True)
print("trace done")
