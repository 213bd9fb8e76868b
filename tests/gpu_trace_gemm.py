"""Diagnostic (not a pytest file): phase timeline of the tcgen05 GEMM at the embedding.yaml shapes, from %globaltimer
stamps written by CTA (0,0,0) (debug hook codae_debug_set_trace).  Prints ns since kernel entry per phase and the
entry-to-entry interval of back-to-back launches."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mui-deepautoencoder_b200"))
import torch
from codae import _C

dev = torch.device("cuda", 0)
bf = torch.bfloat16
lib = _C.lib()
lib.codae_debug_set_trace.argtypes = [ctypes.c_void_p]
names = ["entry", "prologue", "pdl_wait", "first_operands", "mma_issued", "acc_complete", "staged", "cluster_bar1", "stored", "cluster_bar2"]
M, N, K = 128, 1536, 1537
ld = 1600
X = torch.zeros(M, ld, device=dev, dtype=bf); X[:, :1536] = torch.randn(M, 1536, device=dev).to(bf); X[:, 1536] = 1
W = torch.zeros(N, ld, device=dev, dtype=bf); W[:, :1537] = (torch.randn(N, 1537, device=dev) / 40).to(bf)
Y = torch.zeros(M, ld, device=dev, dtype=bf)
dY = torch.randn(M, ld, device=dev).to(bf)
dX = torch.zeros(M, ld, device=dev, dtype=bf)
dW = torch.zeros(N, ld, device=dev)
R = 12
buf = torch.zeros(R * 16, dtype=torch.int64, device=dev)


def run(tag, fn):
    for split in (1, 0):
        _C.set_splitk(dev, bool(split))
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
        rel = [(names[j], int(last[j] - last[0])) for j in range(10) if int(last[j]) != 0]
        gaps = [int(t[i + 1][0] - t[i][0]) for i in range(R - 1)]
        print("%-6s splitk=%d  entry-to-entry ns: %s" % (tag, split, gaps[-5:]), flush=True)
        print("        phases (ns since entry): " + "  ".join("%s=%d" % (n, v) for n, v in rel), flush=True)
    _C.set_splitk(dev, True)


run("fwd", lambda: _C.linear_fwd(X[:, :K], W[:, :K], None, Y, M, N, K, _C.ACT_RELU, _C.BF16))
run("dgrad", lambda: _C.linear_dgrad(dY[:, :N], W[:, :1536], X[:, :1536], dX, M, N, 1536, _C.BF16))
run("wgrad", lambda: _C.linear_wgrad(dY[:, :N], X[:, :K], dW[:, :K], None, M, N, K, _C.BF16))

# ---- inside a CUDA graph: a chain of dependent launches (fwd -> fwd -> ...), as in the training step ----
names2 = names + ["exit_cta0", "exit_last"]
for split in (1, 0):
    _C.set_splitk(dev, bool(split))
    Ys = [torch.zeros(M, ld, device=dev, dtype=bf) for _ in range(2)]
    for y in Ys:
        y[:, 1536] = 1
    def chain():
        src = X
        for i in range(R):
            lib.codae_debug_set_trace(ctypes.c_void_p(buf.data_ptr() + 128 * i))
            dst = Ys[i % 2]
            _C.linear_fwd(src[:, :K], W[:, :K], None, dst, M, N, K, _C.ACT_RELU, _C.BF16)
            src = dst
        lib.codae_debug_set_trace(None)
    chain(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        chain()
    for _ in range(3):
        buf.zero_(); gr.replay()
    torch.cuda.synchronize()
    t = buf.cpu().view(R, 16)
    print("graph chain splitk=%d" % split)
    for i in range(R - 4, R):
        e = int(t[i][0])
        print("   launch %2d: entry-to-entry %6d ns | in-kernel: %s | gap to next entry after last exit: %s" % (
            i, e - int(t[i - 1][0]), " ".join("%s=%d" % (names2[j], int(t[i][j]) - e) for j in (2, 3, 5, 8, 10, 11) if int(t[i][j])),
            (int(t[i + 1][0]) - int(t[i][11])) if i + 1 < R else "-"), flush=True)
_C.set_splitk(dev, True)

# ---- two streams inside a graph: dgrad chain on the capture stream, wgrad on a side stream (the backward pass) ----
side = torch.cuda.Stream(device=dev)
gb = [torch.randn(M, ld, device=dev).to(bf) for _ in range(3)]
dWs = [torch.zeros(N, ld, device=dev) for _ in range(R)]
buf2 = torch.zeros(2 * R * 16, dtype=torch.int64, device=dev)


def backward_like():
    main = torch.cuda.current_stream()
    for i in range(R):
        ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
        with torch.cuda.stream(side):
            lib.codae_debug_set_trace(ctypes.c_void_p(buf2.data_ptr() + 128 * (2 * i)))
            _C.linear_wgrad(gb[i % 3][:, :N], X[:, :K], dWs[i][:, :K], None, M, N, K, _C.BF16)
        lib.codae_debug_set_trace(ctypes.c_void_p(buf2.data_ptr() + 128 * (2 * i + 1)))
        _C.linear_dgrad(gb[i % 3][:, :N], W[:, :1536], X[:, :1536], gb[(i + 1) % 3], M, N, 1536, _C.BF16)
    lib.codae_debug_set_trace(None)
    main.wait_stream(side)


backward_like(); torch.cuda.synchronize()
gr2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr2):
    backward_like()
for _ in range(3):
    buf2.zero_(); gr2.replay()
torch.cuda.synchronize()
t = buf2.cpu().view(2 * R, 16)
t0 = int(t[1][0])
print("two-stream backward inside a graph (ns relative to the first dgrad entry): entry / pdl_wait / last exit")
for i in range(4, 9):
    w, d = t[2 * i], t[2 * i + 1]
    print("   layer %d: wgrad %7d %7d %7d | dgrad %7d %7d %7d" % (i, int(w[0]) - t0, int(w[2]) - t0, int(w[11]) - t0,
                                                               int(d[0]) - t0, int(d[2]) - t0, int(d[11]) - t0), flush=True)
print("trace done")
