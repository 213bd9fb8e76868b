"""GPU parity tests of the norm-free clipped step: codae_linear_wgrad_sq leaves sum(dW^2) per CTA behind and
codae_adam_step_partials derives clip_grad_norm_'s scale from those partials (train_dae_on_embedding.py:212-215),
so the optimizer never makes a pass over the gradients for the norm.
Checked: the gradient is bit-identical to codae_linear_wgrad's on every epilogue path (staged, cluster split-K, direct,
persistent), the partial sums add up to the fp64 norm, and FusedStep(wgrad_sqnorm=True) tracks the cooperative
clip+Adam step and the golden vectors of the reference."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.fixture(scope="module")
def C():
    from codae import _C
    _C.ctx(DEV)
    return _C


def _ru(x, m):
    return (x + m - 1) // m * m


def assert_same_update(got, want):
    """Two runs whose clip scales differ by a few fp32 ulps (different summation order of the norm): all but a handful of
    elements agree to 1e-6 of the weight scale; elements whose effective gradient g*coef + wd*p nearly cancels are
    ill-conditioned in Adam's g/(|g|+eps)-shaped first steps and may move by a fraction of lr."""
    got, want = got.double().cpu(), want.double().cpu()
    d = (got - want).abs() / want.abs().max()
    assert float(d.max()) < 1e-3 and float((d > 2e-6).double().mean()) < 1e-4, (float(d.max()), float((d > 2e-6).double().mean()))


# (batch, out, in, persistent option): staged single pass; cluster split-K; persistent; direct one-tile-per-CTA
@pytest.mark.parametrize("B,o,i,persistent", [(128, 1536, 1536, 1), (32, 1536, 1536, 1), (1024, 128, 832, 1),
                                              (640, 598, 1067, 1), (256, 4096, 4096, 1), (256, 4096, 4096, 0)])
def test_wgrad_sq_equals_wgrad_and_sums_to_the_norm(C, B, o, i, persistent):
    torch.manual_seed(21)
    bf = torch.bfloat16
    K = _ru(i, 8) + 1                      # augmented contraction: bias gradient = the constant-1 column
    ld = _ru(K, 64)
    dY = torch.randn(B, _ru(o, 8)).to(DEV, bf)
    X = torch.zeros(B, ld, device=DEV, dtype=bf)
    X[:, :i] = torch.randn(B, i).to(DEV, bf)
    X[:, _ru(i, 8)] = 1
    C.set_option(DEV, C.OPT_PERSISTENT, persistent)
    try:
        slots = C.linear_wgrad_sq_slots(DEV, B, o, K, C.BF16)
        assert slots >= 1
        want = torch.full((o, ld), 3.0, device=DEV)
        C.linear_wgrad(dY[:, :o], X[:, :K], want[:, :K], None, B, o, K, C.BF16)
        got = torch.full((o, ld), 3.0, device=DEV)
        part = torch.full((slots,), -1.0, dtype=torch.float64, device=DEV)
        C.linear_wgrad_sq(dY[:, :o], X[:, :K], got[:, :K], B, o, K, C.BF16, part)
        assert torch.equal(got, want)                                   # same kernel, same stores
        assert float(part.min()) >= 0.0                                 # every slot was written
        total = float((want[:, :K].double() ** 2).sum())
        assert abs(float(part.sum()) - total) <= 1e-6 * total
        again = torch.zeros_like(part)
        C.linear_wgrad_sq(dY[:, :o], X[:, :K], got[:, :K], B, o, K, C.BF16, again)
        assert torch.equal(again, part)                                 # fixed reduction tree: reproducible
        with pytest.raises(RuntimeError):                               # wrong slot count is refused, nothing launched
            C.linear_wgrad_sq(dY[:, :o], X[:, :K], got[:, :K], B, o, K, C.BF16, torch.zeros(slots + 1, dtype=torch.float64, device=DEV))
    finally:
        C.set_option(DEV, C.OPT_PERSISTENT, 1)


@pytest.mark.parametrize("B,o,i", [(128, 1536, 1536), (32, 1536, 1536), (640, 598, 1067), (128, 192, 328), (100, 832, 128)])
def test_tma_store_epilogue_is_bit_identical(C, B, o, i):
    """CODAE_OPT_TMA_STORE: the staged f32 tile leaves through cp.async.bulk.tensor stores (128-byte swizzle box) instead of
    per-thread stores.  Same gradient bits, same partial sums, ragged edges clipped by the tensor map, padding untouched."""
    torch.manual_seed(23)
    bf = torch.bfloat16
    K = _ru(i, 8) + 1
    ld = _ru(K, 64)
    dY = torch.randn(B, _ru(o, 8)).to(DEV, bf)
    X = torch.zeros(B, ld, device=DEV, dtype=bf)
    X[:, :i] = torch.randn(B, i).to(DEV, bf)
    X[:, _ru(i, 8)] = 1
    slots = C.linear_wgrad_sq_slots(DEV, B, o, K, C.BF16)
    res = {}
    saved = C.get_option(DEV, C.OPT_TMA_STORE)
    for on in (0, 1):
        C.set_option(DEV, C.OPT_TMA_STORE, on)
        try:
            dW = torch.full((o + 3, ld), 5.0, device=DEV)             # 3 guard rows below the matrix
            part = torch.full((slots,), -1.0, dtype=torch.float64, device=DEV)
            C.linear_wgrad_sq(dY[:, :o], X[:, :K], dW[:o, :K], B, o, K, C.BF16, part)
            dW2 = torch.full((o + 3, ld), 5.0, device=DEV)
            C.linear_wgrad(dY[:, :o], X[:, :K], dW2[:o, :K], None, B, o, K, C.BF16)
            torch.cuda.synchronize()
            res[on] = (dW, part, dW2)
        finally:
            C.set_option(DEV, C.OPT_TMA_STORE, saved)
    assert torch.equal(res[1][0], res[0][0]) and torch.equal(res[1][2], res[0][2]) and torch.equal(res[1][0], res[1][2])
    s0, s1 = float(res[0][1].sum()), float(res[1][1].sum())          # threads own different elements on the two paths
    assert float(res[1][1].min()) >= 0 and abs(s1 - s0) <= 1e-6 * s0
    assert float((res[1][0][:, K:] - 5.0).abs().max()) == 0 and float((res[1][0][o:] - 5.0).abs().max()) == 0
    want = dY[:, :o].double().cpu().t().mm(X[:, :K].double().cpu())
    assert rel(res[1][0][:o, :K].cpu().numpy(), want.numpy()) < 1e-4


@pytest.mark.parametrize("n,shadow", [(23_608_320 // 8 + 5, True), (4096, False)])
def test_adam_step_partials_equals_clip_adam(C, n, shadow):
    """Given the same sum of squares, the partials kernel is the cooperative norm+Adam kernel (weights, moments, shadow)."""
    torch.manual_seed(22)
    p0, g = torch.randn(n) * 0.05, torch.randn(n) * 3.0
    A = [p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)]
    Bf = [p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)]
    shA = torch.zeros(n, dtype=torch.bfloat16, device=DEV) if shadow else None
    shB = torch.zeros(n, dtype=torch.bfloat16, device=DEV) if shadow else None
    sqA, sqB = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    ws = C.sqnorm_workspace(DEV)
    G = g.to(DEV)
    # partial sums as the weight-gradient kernels would leave them: 777 chunks
    chunks = torch.tensor_split(G.double() ** 2, 777)
    part = torch.stack([c.sum() for c in chunks]).contiguous()
    for step in range(1, 4):
        C.clip_adam_step(A[0], G, A[1], A[2], shA, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step, 1.0, sqA, ws, 1.0)
        C.adam_step_partials(Bf[0], G, Bf[1], Bf[2], shB, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step, 1.0, part, sqB, 1.0)
        assert float(sqA.item()) > 0 and abs(float(sqA.item()) - float(sqB.item())) <= 1e-6 * float(sqA.item())
        for x, y in zip(A, Bf):
            assert_same_update(y, x)
    if shadow:
        assert torch.equal(shB.view(torch.int16), Bf[0].to(torch.bfloat16).view(torch.int16))


def _build(g, dtype):
    from test_gpu_training import build_embedding
    return build_embedding(g, dtype=dtype)


@pytest.mark.parametrize("graph", [False, True])
def test_fused_step_wgrad_sqnorm_tracks_default_path(graph):
    """Same golden run through both optimizer paths of the bf16 engine: gradients bit-identical, norm within 1e-6, weights
    within fp32 rounding of each other, and within the bf16 tolerance (1e-2) of the reference's fp32 run."""
    from codae.tool import FusedStep
    g = np.load(os.path.join(GOLDEN, "emb_mid.npz"))
    B = int(g["B"])
    runs = {}
    for sq in (False, True):
        ds, model, cor = _build(g, "bf16")
        fs = FusedStep(model, cor, ds.data, lr=float(g["lr"]), weight_decay=float(g["wd"]), clip=True, use_graph=graph,
                       wgrad_sqnorm=sq)
        rec = []
        for rep in range(2):                       # the golden batches twice: the graph is replayed from the third call on
            s = 0
            while "idx%d" % s in g:
                fs.step(torch.from_numpy(g["idx%d" % s]).to(DEV), run=0)
                rec.append((fs.last_loss(B), float(fs.sqnorm.item()), fs.gflat.clone(), model.flat.clone()))
                if rep == 0:
                    assert abs(rec[-1][0] - float(g["loss%d" % s])) <= 1e-2 * float(g["loss%d" % s])
                s += 1
        runs[sq] = rec
    assert runs[False][0][0] == runs[True][0][0]                         # first step: same forward, bit for bit
    for a, b in zip(runs[False], runs[True]):
        assert abs(a[0] - b[0]) <= 1e-5 * abs(a[0])
        assert a[1] > 0 and abs(a[1] - b[1]) <= 1e-6 * a[1]              # same norm (different summation order)
        assert_same_update(b[3], a[3])                                   # same weights
    assert torch.equal(runs[False][0][2], runs[True][0][2])              # first step: bit-identical gradients
