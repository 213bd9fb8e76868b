"""compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer_closed.txt), so out-of-bounds WRITES are hunted the manual
way: every output buffer of every hot-path kernel is embedded in a larger allocation filled with a sentinel -- guard rows before
and after, guard columns inside the pitch padding -- and after the launch every byte the kernel does not own must still hold the
sentinel.  TMA bulk stores (clipped at 16-byte granularity), DSMEM split-K epilogues, ragged tiles (N % 4 != 0, the bias column)
and the three-plane outputs of the fp32-parity engine are the interesting cases."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
GUARD = 3          # guard rows on either side


@pytest.fixture(scope="module")
def C():
    from codae import _C
    return _C


def guarded(rows, pitch, dtype, fill, planes=None):
    """(view [rows, pitch] (or [3, rows, pitch]) inside a sentinel-filled allocation, checker(cols_written))."""
    shape = (rows + 2 * GUARD, pitch) if planes is None else (planes, rows + 2 * GUARD, pitch)
    fill = float("nan")                   # no kernel output is ever NaN; any finite sentinel is a legitimate bf16 output value
    base = torch.full(shape, fill, dtype=dtype, device=DEV)
    view = base[GUARD:GUARD + rows] if planes is None else base[:, GUARD:GUARD + rows]

    def check(cols):
        b = base.float()
        if planes is None:
            assert bool(b[:GUARD].isnan().all()) and bool(b[GUARD + rows:].isnan().all()), "guard rows overwritten"
            assert bool(b[GUARD:GUARD + rows, cols:].isnan().all()), "pitch padding overwritten"
            assert not bool(b[GUARD:GUARD + rows, :cols].isnan().any()), "owned region not fully written"
        else:
            assert bool(b[:, :GUARD].isnan().all()) and bool(b[:, GUARD + rows:].isnan().all()), "guard rows overwritten"
            assert bool(b[:, GUARD:GUARD + rows, cols:].isnan().all()), "pitch padding overwritten"
            assert not bool(b[:, GUARD:GUARD + rows, :cols].isnan().any()), "owned region not fully written"
    return view, check


def x3(C, t, pitch):
    rows, cols = t.shape
    src = torch.zeros(rows, pitch, device=DEV)
    src[:, :cols] = t.to(DEV)
    out = C.new_x3((rows, pitch), DEV)
    C.split_x3(src, out)
    return out[:, :, :cols]


SHAPES = [(128, 1536, 1537), (32, 600, 1064), (200, 192, 193), (1, 128, 832), (128, 64, 65), (1024, 1536, 513), (300, 328, 73)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_bf16_engine_writes_only_what_it_owns(C, M, N, K):
    torch.manual_seed(1)
    bf = torch.bfloat16
    Kp, Np = (K + 7) // 8 * 8, (N + 7) // 8 * 8
    X = torch.zeros(M, Kp + 8, device=DEV, dtype=bf); X[:, :K] = torch.randn(M, K, device=DEV).to(bf) + 3
    W = torch.zeros(N, Kp + 8, device=DEV, dtype=bf); W[:, :K] = (torch.randn(N, K, device=DEV) / 30).to(bf)
    dY = torch.zeros(M, Np + 8, device=DEV, dtype=bf); dY[:, :N] = torch.randn(M, N, device=DEV).to(bf)
    for out_dt in (bf, torch.float32):
        Y, chk = guarded(M, Np + 16, out_dt, -7.0)
        C.linear_fwd(X[:, :K], W[:, :K], None, Y[:, :N], M, N, K, C.ACT_NONE, C.BF16)
        torch.cuda.synchronize(); chk(N)
    Kd = K - 1 if K % 8 else K                    # dgrad output width: a multiple of 8 (layer inputs are)
    dX, chk = guarded(M, Kp + 16, bf, -7.0)
    C.linear_dgrad(dY[:, :N], W[:, :Kd], None, dX[:, :Kd], M, N, Kd, C.BF16)
    torch.cuda.synchronize(); chk(Kd)
    dW, chk = guarded(N, (K + 3) // 4 * 4 + 8, torch.float32, -7.0)
    C.linear_wgrad(dY[:, :N], X[:, :K], dW[:, :K], None, M, N, K, C.BF16)
    torch.cuda.synchronize(); chk(K)
    slots = C.linear_wgrad_sq_slots(DEV, M, N, K, C.BF16)
    sq = torch.full((slots + 2,), -7.0, dtype=torch.float64, device=DEV)
    dW, chk = guarded(N, (K + 3) // 4 * 4 + 8, torch.float32, -7.0)
    C.linear_wgrad_sq(dY[:, :N], X[:, :K], dW[:, :K], M, N, K, C.BF16, sq[1:1 + slots])
    torch.cuda.synchronize(); chk(K)
    assert float(sq[0]) == -7.0 and float(sq[-1]) == -7.0 and bool((sq[1:1 + slots] >= 0).all())


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_fp32_parity_engine_writes_only_what_it_owns(C, M, N, K):
    torch.manual_seed(2)
    Kp, Np = (K + 7) // 8 * 8, (N + 7) // 8 * 8
    X = x3(C, torch.randn(M, K) + 3, Kp + 8)
    W = x3(C, torch.randn(N, K) / 30, Kp + 8)
    dY = x3(C, torch.randn(M, N), Np + 8)
    Y, chk = guarded(M, Np + 16, torch.float32, -7.0)
    C.linear_fwd(X, W, None, Y[:, :N], M, N, K, C.ACT_RELU, C.F32X3)
    torch.cuda.synchronize(); chk(N)
    Yp, chk = guarded(M, Np + 16, torch.bfloat16, -7.0, planes=3)
    C.linear_fwd(X, W, None, Yp[:, :, :N], M, N, K, C.ACT_NONE, C.F32X3)
    torch.cuda.synchronize(); chk(N)
    Kd = K - 1 if K % 8 else K
    dXp, chk = guarded(M, Kp + 16, torch.bfloat16, -7.0, planes=3)
    C.linear_dgrad(dY, W[:, :, :Kd], X[:, :, :Kd], dXp[:, :, :Kd], M, N, Kd, C.F32X3)
    torch.cuda.synchronize(); chk(Kd)        # the ReLU-mask epilogue writes zeros too: every owned element is written
    slots = C.linear_wgrad_sq_slots(DEV, M, N, K, C.F32X3)
    sq = torch.full((slots + 2,), -7.0, dtype=torch.float64, device=DEV)
    dW, chk = guarded(N, (K + 3) // 4 * 4 + 8, torch.float32, -7.0)
    C.linear_wgrad_sq(dY, X, dW[:, :K], M, N, K, C.F32X3, sq[1:1 + slots])
    torch.cuda.synchronize(); chk(K)
    assert float(sq[0]) == -7.0 and float(sq[-1]) == -7.0 and bool((sq[1:1 + slots] >= 0).all())


@pytest.mark.parametrize("out", ["f32", "bf16", "x3"])
def test_corrupt_and_loss_kernels_respect_pitch_padding(C, out):
    from codae.tool import Corrupter
    torch.manual_seed(3)
    S, E, N, B = 3, 64, 300, 77
    io = S * E
    arch = [dict(name=str(i), size=E, type="regression", position=i * E) for i in range(S)]
    data = torch.randn(N, io, device=DEV) + 5
    cor = Corrupter(N, arch, 1, DEV, seed=5)
    table, bits, col_var, nmiss = cor.device_tables()
    idx = torch.randperm(N, device=DEV)[:B]
    pitch = io + 24
    if out == "x3":
        cx, chk = guarded(B, pitch, torch.bfloat16, -7.0, planes=3)
        # the kernel owns a whole [3, B, ld] buffer: hand it one (inside the guard rows the plane stride would be wrong)
        own = C.new_x3((B, pitch), DEV); own.fill_(-7.0)
        mid = torch.full((B,), -1, dtype=torch.int32, device=DEV)
        C.corrupt_fwd(data, idx, B, table, 0, bits, col_var, io, own, None, mid)
        torch.cuda.synchronize()
        assert bool((own[:, :, io:].float() == -7.0).all()) and bool((mid >= 0).all())
        want = data[idx] * (C.x3_to_f32(own)[:, :io] != 0)
        assert bool(((C.x3_to_f32(own)[:, :io] - want).abs() <= want.abs() * 2.0 ** -22).all())
        g = C.new_x3((B, pitch), DEV); g.fill_(-7.0)
    else:
        dt = torch.float32 if out == "f32" else torch.bfloat16
        cx, chk = guarded(B, pitch, dt, -7.0)
        mid = torch.full((B,), -1, dtype=torch.int32, device=DEV)
        C.corrupt_fwd(data, idx, B, table, 0, bits, col_var, io, cx, None, mid)
        torch.cuda.synchronize(); chk(io)
        g, gchk = guarded(B, pitch, dt, -7.0)
    y = torch.randn(B, io + 8, device=DEV)
    acc = torch.zeros(4, dtype=torch.float64, device=DEV)
    C.mse_loss_fwd_bwd(data, idx, y, mid, bits, col_var, B, io, 2.0 / (B * io), g, acc, C.loss_workspace(DEV))
    torch.cuda.synchronize()
    if out == "x3":
        assert bool((g[:, :, io:].float() == -7.0).all())
        want = (2.0 / (B * io)) * (y[:, :io] - data[idx])
        assert bool(((C.x3_to_f32(g)[:, :io] - want).abs() <= want.abs() * 2.0 ** -21 + 1e-30).all())
    else:
        gchk(io)
    want_sum = float(((y[:, :io] - data[idx]).double() ** 2).sum())
    assert abs(float(acc[3]) - want_sum) <= 1e-5 * want_sum


@pytest.mark.parametrize("shadow", [None, "bf16", "x3"])
def test_optimizer_kernels_stop_at_n(C, shadow):
    torch.manual_seed(4)
    n, pad = 4096 * 3 + 8, 64
    def buf(dtype=torch.float32, rows=None):
        shape = (n + 2 * pad,) if rows is None else (rows, n)
        return torch.full(shape, -7.0, dtype=dtype, device=DEV)
    p, g, m, v = buf(), buf(), buf(), buf()
    for t, init in ((p, torch.randn(n)), (g, torch.randn(n) * 0.01), (m, torch.zeros(n)), (v, torch.zeros(n))):
        t[pad:pad + n] = init.to(DEV)
    sh = None
    if shadow == "bf16":
        shb = torch.full((n + 2 * pad,), -7.0, dtype=torch.bfloat16, device=DEV); sh = shb[pad:pad + n]
    elif shadow == "x3":
        shb = torch.full((3 * n + 2 * pad,), -7.0, dtype=torch.bfloat16, device=DEV); sh = shb[pad:pad + 3 * n].view(3, n)
    sqn = torch.zeros(1, device=DEV)
    P, G, Mv, V = (t[pad:pad + n] for t in (p, g, m, v))
    C.clip_adam_step(P, G, Mv, V, sh, 1e-3, 0.9, 0.999, 1e-8, 1e-4, 1, 1.0, sqn, C.sqnorm_workspace(DEV), 1.0)
    part = (G.double() ** 2).sum().view(1)
    C.adam_step_partials(P, G, Mv, V, sh, 1e-3, 0.9, 0.999, 1e-8, 1e-4, 2, 1.0, part, sqn, 1.0)
    C.adam_step(P, G, Mv, V, sh, 1e-3, 0.9, 0.999, 1e-8, 1e-4, 3, -1.0, None, 1.0)
    torch.cuda.synchronize()
    for t in (p, g, m, v):
        assert bool((t[:pad] == -7.0).all()) and bool((t[pad + n:] == -7.0).all())
    if shadow:
        assert bool((shb[:pad].float() == -7.0).all()) and bool((shb[-pad:].float() == -7.0).all())
        assert not bool((sh.float() == -7.0).any())


def test_scoring_outputs_and_workspace_bounds(C):
    from codae.tool.inference import ComplementarityScorer
    torch.manual_seed(6)
    for dt in (torch.float32, torch.bfloat16):
        cat = torch.rand(50_001, 512, device=DEV).to(dt)
        q = torch.rand(5, 512, device=DEV)
        sc = ComplementarityScorer(cat, 512, "sqerr", k=10)
        s, i = sc.topk_local(q)
        torch.cuda.synchronize()
        assert bool((i >= 0).all()) and bool((i < 50_001).all()) and bool((s[:, 1:] >= s[:, :-1]).all())
