"""GPU parity tests of alternative code paths that must agree with the baseline path bit for bit (or with the goldens):
  * the bulk-store (cp.async.bulk.tensor) epilogue of the persistent GEMM kernel vs per-thread stores (default: bulk store,
    measured 7.155 -> 6.970 ms/step on the polyvore-shaped step, profiles/r02_notes.md);
  * the whole-network kernels for tabular widths (codae_tiny_mlp_fwd / _bwd) vs the abalone goldens of the reference.
Each subprocess case runs under a timeout: a protocol bug in a persistent kernel must not take the test session with it."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DEV = torch.device("cuda", 0)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


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


def test_persistent_bulk_store_ragged_bf16_columns():
    """CODAE_OPT_TMA_STORE_PERSISTENT with bf16 outputs whose width is not a multiple of 64 / 8: odd chunk counts in the last
    column tile (half-filled box) and a row tail stored by the threads.  (The even / f32 cases are in test_gpu_optin.py.)"""
    code = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "..")); sys.path.insert(0, os.path.join(%r, "..", "mui-deepautoencoder_b200"))
import torch
from codae import _C as C
DEV = torch.device("cuda", 0)
torch.manual_seed(41)
bf = torch.bfloat16
for M, N, K in [(8192, 1067, 264), (8192, 1040, 136), (16384, 600, 72)]:
    ldn = (N + 7) // 8 * 8 + 8
    X = torch.randn(M, K).to(DEV, bf)
    W = (torch.randn(N, K) / 8).to(DEV, bf)
    res = {}
    for on in (0, 1):
        C.set_option(DEV, C.OPT_TMA_STORE_PERSISTENT, on)
        Y = torch.full((M + 1, ldn), 3.0, device=DEV, dtype=bf)
        C.linear_fwd(X, W, None, Y[:M, :N], M, N, K, C.ACT_RELU, C.BF16)
        torch.cuda.synchronize()
        res[on] = Y
    C.set_option(DEV, C.OPT_TMA_STORE_PERSISTENT, 1)
    assert torch.equal(res[0].view(torch.int16), res[1].view(torch.int16)), (M, N, K, int((res[0] != res[1]).sum()))
print("RAGGED OK")
''' % (HERE, HERE)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0 and "RAGGED OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_cta_pair_persistent_kernel_is_bit_identical():
    """CODAE_OPT_CTA_PAIR (default on): the persistent kernel as CTA pairs (tcgen05.mma.cta_group::2 on 256 x 256 tiles, the
    peer's TMA loads counted on the leader's barriers, multicast commits) against the single-CTA persistent kernel -- forward
    (ReLU bf16 / plain f32), masked input gradient and weight gradient + its sum of squares, ragged row tiles, a peer CTA whose
    rows are all out of range, ragged column tiles (bias-gradient column, n_eff = 32 / 64 / 96): same bits, guard rows and
    padding columns untouched; the weight gradients also against fp64 products (tools/probes/cta_pair_check.py)."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "..", "tools", "probes", "cta_pair_check.py")], capture_output=True,
                       text=True, timeout=240)
    assert r.returncode == 0 and "PAIR OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("name", ["abalone_k1", "abalone_k3"])
def test_tiny_mlp_abalone_matches_reference(name):
    """FusedStep(tiny_mlp=True): the abalone model's forward and backward passes as one launch each (codae_tiny_mlp_fwd / _bwd)
    against the golden vectors of the reference (same checks as test_abalone_fused_and_legacy's fused path)."""
    code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "..")); sys.path.insert(0, os.path.join(%r, "..", "mui-deepautoencoder_b200"))
import numpy as np, torch
from conftest import GOLDEN
from test_gpu_training import load_params, flat_grads, flat_params, rel, DEV
from oracle.gen_golden import abalone_arch
from codae.dataset import MixedVariableDataset
from codae.model import MixedVariableDenoisingAutoencoder
from codae.tool import Corrupter, FusedStep
g = np.load(os.path.join(GOLDEN, %r + ".npz"))
arch = abalone_arch()
k_max, B = int(g["k_max"]), int(g["B"])
ds = MixedVariableDataset.from_arch(arch, torch.from_numpy(g["data"]))
m = MixedVariableDenoisingAutoencoder(arch, 11, int(g["z"]), DEV, 2, 2, bool(g["steep"]))
load_params(m, g["init"], g["shapes"])
m.to(DEV); ds.to(DEV)
cor = Corrupter(ds.nb_observation, arch, k_max, DEV)
cor.mask_to_use = torch.from_numpy(g["mask_to_use"])
fs = FusedStep(m, cor, ds.data, lr=float(g["lr"]), weight_decay=float(g["wd"]), clip=True, tiny_mlp=True,
               mixed=dict(arch=arch, weight=list(g["weight"]), norm_scale=torch.from_numpy(g["norm_scale"]),
                          norm_min=torch.from_numpy(g["norm_min"]), norm_first=3))
assert fs.tiny_mlp
for s in range(3):
    fs.step(torch.from_numpy(g["idx%%d" %% s]).to(DEV), run=int(g["run%%d" %% s]))
    assert abs(fs.last_loss(B) - float(g["loss%%d" %% s])) <= 1e-5 * float(g["loss%%d" %% s])
    assert rel(flat_grads(m), g["grads%%d" %% s]) < 1e-5
    assert rel(flat_params(m), g["post%%d" %% s]) < 1e-5
assert fs.kernel_launches == 7
print("TINY OK")
''' % (HERE, HERE, HERE, name)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0 and "TINY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_load_state_dict_refreshes_the_weight_shadow_and_engine_changes_are_refused(dtype):
    """The tensor-core GEMMs read a copy of the weights (bf16, or the three planes of the fp32-parity engine): loading a
    checkpoint must rewrite it, and switching the engine under a live FusedStep must fail loudly."""
    from codae import _C
    from codae.dataset import ConcatenatedEmbeddingDataset
    from codae.model import EmbeddingDenoisingAutoencoder
    from codae.tool import Corrupter, FusedStep
    torch.manual_seed(2)
    S, E, N, B = 3, 64, 256, 32
    ds = ConcatenatedEmbeddingDataset.from_tensors([torch.randn(N, E).abs() for _ in range(S)])
    donor = EmbeddingDenoisingAutoencoder(S * E, S * E, E, 2, 2, False)
    sd = {k: v.clone() for k, v in donor.state_dict().items()}
    out = []
    for load_after in (False, True):
        torch.manual_seed(3)
        m = EmbeddingDenoisingAutoencoder(S * E, S * E, E, 2, 2, False)
        if not load_after:
            m.load_state_dict(sd)
        m.set_compute_dtype(dtype).to(DEV)
        if load_after:
            ds.to(DEV)
        cor = Corrupter(N, ds.arch, 1, DEV, seed=4)
        fs = FusedStep(m, cor, ds.data.to(DEV), lr=1e-3, weight_decay=0.0, clip=True)
        if load_after:
            m.load_state_dict(sd)                        # after the FusedStep (and its shadow) exist
        shadow = m.gemm_weights()
        if dtype == "bf16":
            assert torch.equal(shadow.view(torch.int16), m.flat.to(torch.bfloat16).view(torch.int16))
        else:
            want = torch.empty_like(shadow)
            _C.split_x3(m.flat, want)
            assert torch.equal(shadow.view(torch.int16), want.view(torch.int16))
        fs.step(torch.arange(B, device=DEV))
        out.append((fs.last_loss(B), m.flat.clone()))
    assert out[0][0] == out[1][0] and torch.equal(out[0][1], out[1][1])
    m.set_compute_dtype("fp32_simt")
    with pytest.raises(RuntimeError, match="build a new FusedStep"):
        fs.step(torch.arange(B, device=DEV))


def test_out_of_range_observation_ids_are_refused_before_launch():
    """The row gather of codae_corrupt_fwd is where a caller-supplied index addresses memory.  train_epoch checks its ids up front
    and raises before anything is launched (the reference raises IndexError); inside the kernel an id outside [0, n_rows) traps
    instead of gathering foreign bytes -- that path is deliberately NOT exercised here (a trapped kernel ends the process's CUDA
    context and shows up in the host's fault log)."""
    from codae.dataset import ConcatenatedEmbeddingDataset
    from codae.model import EmbeddingDenoisingAutoencoder
    from codae.tool import Corrupter, FusedStep
    ds = ConcatenatedEmbeddingDataset.from_tensors([torch.randn(256, 64).abs() for _ in range(3)])
    m = EmbeddingDenoisingAutoencoder(192, 192, 64, 2, 2, False)
    m.to(DEV)
    ds.to(DEV)
    fs = FusedStep(m, Corrupter(256, ds.arch, 1, DEV, seed=1), ds.data, lr=1e-3, weight_decay=0.0)
    before = m.flat.clone()
    with pytest.raises(Exception, match="Observation index out of range"):
        fs.train_epoch(torch.arange(200, 300, device=DEV), 32)
    with pytest.raises(Exception, match="Observation index out of range"):
        fs.train_epoch(torch.tensor([5, -1, 7], device=DEV), 2)
    torch.cuda.synchronize()
    assert torch.equal(before, m.flat)                       # nothing ran
    assert fs.train_epoch(torch.arange(0, 256, device=DEV), 32) == 8
