"""CPU: the arithmetic of the whole-network kernels for tabular widths (csrc/tiny_mlp.h, the code tiny_mlp.cu executes per
thread) compiled with g++ and checked against the oracle's forward / backward (oracle/codae_oracle.py) on the abalone model's
layer sizes: reconstructions, every weight / bias gradient (the bias gradient is the constant-1 column of the augmented
contraction) and the input-gradient chain with its ReLU masks, independent of the emulated thread count."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT

DRIVER = r'''
#include "tiny_mlp.h"
extern "C" void tiny_cpu_fwd(const TinyLayer* L, int nl, const float* flat, float** acts, long long ld_act, int B, int nthreads) {
    for (int l = 0; l < nl; ++l)
        for (int tid = 0; tid < nthreads; ++tid)
            tiny_fwd_layer(L[l], flat + L[l].w_off, acts[l], acts[l + 1], ld_act, 0, B, tid, nthreads);
}
extern "C" void tiny_cpu_bwd(const TinyLayer* L, int nl, const float* flat, float* gflat, float** acts, long long ld_act, float** g3,
                             long long ld_g, int B, int nthreads) {
    for (int l = nl - 1; l >= 0; --l) {
        for (int tid = 0; tid < nthreads; ++tid)
            tiny_wgrad_layer(L[l], g3[l % 3], ld_g, acts[l], ld_act, gflat + L[l].w_off, B, tid, nthreads);
        if (l > 0)
            for (int tid = 0; tid < nthreads; ++tid)
                tiny_dgrad_layer(L[l], g3[l % 3], flat + L[l].w_off, acts[l], ld_act, L[l - 1].relu, g3[(l - 1) % 3], ld_g, B, tid, nthreads);
    }
}
extern "C" int tiny_sizeof_layer(void) { return (int)sizeof(TinyLayer); }
'''


class TinyLayer(ctypes.Structure):
    _fields_ = [("w_off", ctypes.c_int64), ("ld", ctypes.c_int32), ("bcol", ctypes.c_int32), ("in_", ctypes.c_int32),
                ("out", ctypes.c_int32), ("relu", ctypes.c_int32), ("pad", ctypes.c_int32)]


def ru(x, m):
    return (x + m - 1) // m * m


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("tiny")
    (d / "driver.cpp").write_text(DRIVER)
    so = d / "libtiny_cpu.so"
    r = subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "mui-deepautoencoder_b200", "csrc"),
                        str(d / "driver.cpp"), "-o", str(so)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    l = ctypes.CDLL(str(so))
    assert l.tiny_sizeof_layer() == ctypes.sizeof(TinyLayer)
    return l


def ptr_array(arrays):
    return (ctypes.POINTER(ctypes.c_float) * len(arrays))(*[a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) for a in arrays])


@pytest.mark.parametrize("dims,relu", [
    ([(11, 11), (11, 7), (7, 4), (4, 8), (8, 11), (11, 11)], [True, True, False, True, True, False]),      # abalone.yaml, non-steep
    ([(11, 11)] * 6, [True, True, False, True, True, False]),                                               # steep
    ([(24, 24), (24, 9), (9, 24)], [True, False, False]),
])
@pytest.mark.parametrize("B,nthreads", [(64, 1), (64, 256), (5, 96)])
def test_tiny_mlp_matches_the_oracle(lib, dims, relu, B, nthreads):
    from oracle import codae_oracle as O
    torch.manual_seed(17)
    W = [torch.randn(o, i) * 0.4 for i, o in dims]
    b = [torch.randn(o) * 0.1 for _, o in dims]
    x = torch.rand(B, dims[0][0])
    # flat augmented layout (FlatMLP.layout): W'[out, ld], bias in column round_up(in, 8), ld = round_up(that + 1, 64)
    layers, off = [], 0
    for (i, o), r in zip(dims, relu):
        bcol = ru(i, 8)
        ld = ru(bcol + 1, 64)
        layers.append(TinyLayer(off, ld, bcol, i, o, 1 if r else 0, 0))
        off += o * ld
    flat = np.zeros(off, dtype=np.float32)
    for ly, w, bb in zip(layers, W, b):
        v = flat[ly.w_off:ly.w_off + ly.out * ly.ld].reshape(ly.out, ly.ld)
        v[:, :ly.in_] = w.numpy()
        v[:, ly.bcol] = bb.numpy()
    ld_act = max(ly.ld for ly in layers)
    widths = [dims[0][0]] + [o for _, o in dims]
    acts = []
    for w_ in widths:
        a = np.zeros((B, ld_act), dtype=np.float32)
        a[:, ru(w_, 8)] = 1.0                                   # the constant-1 column of the augmented layout
        acts.append(a)
    acts[0][:, :widths[0]] = x.numpy()
    L = (TinyLayer * len(layers))(*layers)
    lib.tiny_cpu_fwd(L, len(layers), flat.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ptr_array(acts), ctypes.c_longlong(ld_act), B,
                     nthreads)
    y_ref, acts_ref = O.forward(W, b, relu, x, keep=True)
    for l in range(1, len(widths)):
        np.testing.assert_allclose(acts[l][:, :widths[l]], acts_ref[l].numpy(), rtol=1e-5, atol=1e-6)
        assert np.all(acts[l][:, ru(widths[l], 8)] == 1.0)       # the ones column survives (only columns < out are written)
    # backward from a random dL/dy
    dy = torch.randn(B, widths[-1]) * 0.05
    g3 = [np.zeros((B, ld_act), dtype=np.float32) for _ in range(3)]
    nl = len(layers)
    g3[(nl - 1) % 3][:, :widths[-1]] = dy.numpy()
    gflat = np.full(off, 7.0, dtype=np.float32)
    lib.tiny_cpu_bwd(L, nl, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), gflat.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                     ptr_array(acts), ctypes.c_longlong(ld_act), ptr_array(g3), ctypes.c_longlong(ld_act), B, nthreads)
    gW, gb = O.backward(W, relu, acts_ref, dy)
    for ly, w_ref, b_ref in zip(layers, gW, gb):
        v = gflat[ly.w_off:ly.w_off + ly.out * ly.ld].reshape(ly.out, ly.ld)
        scale = max(float(w_ref.abs().max()), 1e-12)
        assert float(np.abs(v[:, :ly.in_] - w_ref.numpy()).max()) <= 2e-6 * max(scale, 1.0) + 1e-5 * scale
        assert float(np.abs(v[:, ly.bcol] - b_ref.numpy()).max()) <= 1e-5 * max(float(b_ref.abs().max()), 1e-3)
        assert np.all(v[:, ly.in_:ly.bcol] == 0.0)              # gradient of the zero padding is exactly zero
        assert np.all(v[:, ly.bcol + 1:] == 7.0)                # columns beyond the bias column are never written
