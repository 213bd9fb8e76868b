"""GPU parity tests of OPT-IN switches: code paths that are bit-identical to the default path on a B200 (these tests passed
there) but whose effect on step time has not been measured yet, so they stay off by default:
  FusedStep(layerwise_adam=True) / CODAE_LAYERWISE_ADAM=1      per-layer Adam beside the dgrad chain (un-clipped steps)
  CODAE_OPT_TMA_STORE_PERSISTENT / CODAE_TMA_STORE_PERSISTENT=1 bulk-store epilogue of the persistent GEMM kernel"""
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


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_layerwise_adam_equals_one_update_after_the_backward_pass(graph, dtype):
    """FusedStep(layerwise_adam=True, clip=False): Adam per layer on the weight-gradient stream as soon as wgrad(l) and
    dgrad(l) are done.  Without clipping every element's update only depends on its own gradient, so weights, moments and
    the bf16 shadow must equal the one-launch update bit for bit (same kernel arithmetic, different launch spans)."""
    from test_gpu_training import build_embedding
    from codae.tool import FusedStep
    g = np.load(os.path.join(GOLDEN, "emb_mid.npz"))
    runs = {}
    for lw in (False, True):
        ds, model, cor = build_embedding(g, dtype=dtype)
        fs = FusedStep(model, cor, ds.data, lr=float(g["lr"]), weight_decay=float(g["wd"]), clip=False, use_graph=graph,
                       layerwise_adam=lw, wgrad_sqnorm=False)
        for rep in range(2):
            s = 0
            while "idx%d" % s in g:
                fs.step(torch.from_numpy(g["idx%d" % s]).to(DEV), run=0)
                s += 1
        torch.cuda.synchronize()
        runs[lw] = (model.flat.clone(), fs.m.clone(), fs.v.clone(), None if model.flat_bf16 is None else model.flat_bf16.clone(),
                    fs.last_loss(int(g["B"])), int(fs.step_dev.item()))
    a, b = runs[False], runs[True]
    assert a[5] == b[5] == 2 * s
    assert a[4] == b[4]
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y)
    if a[3] is not None:
        assert torch.equal(a[3].view(torch.int16), b[3].view(torch.int16))


def test_persistent_kernel_bulk_store_epilogue_is_bit_identical():
    """CODAE_OPT_TMA_STORE_PERSISTENT: the persistent kernel's epilogue warps stage 32 x 128-byte boxes and issue
    cp.async.bulk.tensor stores.  Every fused epilogue (ReLU, ReLU mask, plain f32), bf16 and f32 outputs, ragged last
    column tiles (bias-gradient column) and ragged row tiles must give the per-thread-store result bit for bit, with
    padding columns and guard rows untouched."""
    from codae import _C as C
    C.ctx(DEV)
    torch.manual_seed(31)
    bf = torch.bfloat16
    M, N, K = 4000, 4096, 264            # fwd / dgrad: 32 x 16 tiles of 128 x 256 (> 2 per SM), last row tile ragged
    X = torch.randn(M, K).to(DEV, bf)
    W = (torch.randn(N, K) / 16).to(DEV, bf)
    dY = torch.randn(M, N).to(DEV, bf)
    Wd = (torch.randn(264, N) / 64).to(DEV, bf)             # dgrad: dX[M, N] = dY2[M, 264] . Wd[264, N], masked by A_prev
    dY2 = torch.randn(M, 264).to(DEV, bf)
    A_prev = torch.randn(M, N).to(DEV, bf)
    Mb, Nf, Kf, ld = 256, 4096, 4097, 4160                   # wgrad of an augmented 4096-wide layer: 17th column tile = bias column
    dYw = torch.randn(Mb, Nf).to(DEV, bf)
    Xa = torch.zeros(Mb, ld, device=DEV, dtype=bf)
    Xa[:, :4096] = torch.randn(Mb, 4096).to(DEV, bf)
    Xa[:, 4096] = 1
    saved = C.get_option(DEV, C.OPT_TMA_STORE_PERSISTENT)
    res = {}
    try:
        for on in (0, 1):
            C.set_option(DEV, C.OPT_TMA_STORE_PERSISTENT, on)
            Yb = torch.full((M + 2, N + 8), 3.0, device=DEV, dtype=bf)
            Yf = torch.full((M + 2, N + 8), 3.0, device=DEV)
            C.linear_fwd(X, W, None, Yb[:M, :N], M, N, K, C.ACT_RELU, C.BF16)
            C.linear_fwd(X, W, None, Yf[:M, :N], M, N, K, C.ACT_NONE, C.BF16)
            dXb = torch.full((M + 2, N + 8), 3.0, device=DEV, dtype=bf)
            C.linear_dgrad(dY2, Wd, A_prev, dXb[:M, :N], M, 264, N, C.BF16)
            dW = torch.full((Nf + 2, ld), 7.0, device=DEV)
            slots = C.linear_wgrad_sq_slots(DEV, Mb, Nf, Kf, C.BF16)
            part = torch.zeros(slots, dtype=torch.float64, device=DEV)
            C.linear_wgrad_sq(dYw, Xa[:, :Kf], dW[:Nf, :Kf], Mb, Nf, Kf, C.BF16, part)
            torch.cuda.synchronize()
            res[on] = (Yb, Yf, dXb, dW, part)
    finally:
        C.set_option(DEV, C.OPT_TMA_STORE_PERSISTENT, saved)
    for a, b in zip(res[0][:4], res[1][:4]):
        av, bv = (a.view(torch.int16), b.view(torch.int16)) if a.dtype == bf else (a, b)
        assert torch.equal(av, bv), float((a.float() - b.float()).abs().max())
    assert torch.equal(res[0][4], res[1][4])
    want = X.double().cpu().mm(W.double().cpu().t())
    assert rel(res[1][1][:M, :N].cpu().numpy(), want.numpy()) < 1e-4
