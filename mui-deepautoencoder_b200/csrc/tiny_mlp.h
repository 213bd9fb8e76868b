// Whole-network forward / backward of a TINY multilayer perceptron (tabular models: abalone's 11-wide layers), host + device.
// The arithmetic lives here as per-element phase functions over (tid, nthreads) so that the CUDA kernels (tiny_mlp.cu: one
// launch for all layers) and the CPU unit test (tests/test_tiny_mlp_cpu.py, g++) execute the SAME code.
// Layout: the augmented flat parameter buffer of FlatMLP (codae/model/_flat_mlp.py): layer l is W'[out, ld] at w_off with the
// bias in column bcol = round_up(in, 8); activations [B, ld_act] carry a constant 1 in column bcol of the NEXT layer and zeros
// in the padding, so  y[r, o] = sum_{k <= bcol} a[r, k] * W'[o, k]  adds the bias and  dW'[o, k] = sum_r g[r, o] * a[r, k]
// yields the bias gradient in column bcol.
// Reference: nn.Sequential of Linear / ReLU in codae/model/mixed_variable_denoising_autoencoder.py:45-181 and autograd's
// mm / threshold_backward for loss.backward() (script/train_dae_on_abalone.py:219).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TINY_HD __host__ __device__ __forceinline__
#else
#define TINY_HD static inline
#endif

#define CODAE_TINY_MAX_LAYERS 8

typedef struct TinyLayer {
    int64_t w_off;             // offset of W' in the flat buffers (elements)
    int32_t ld, bcol;          // pitch of W', bias column
    int32_t in, out;           // nn.Linear(in, out)
    int32_t relu, pad;         // ReLU follows this layer
} TinyLayer;

// a_out[r, o] = act(sum_k a_in[r, k] W'[o, k]),  rows [row0, row1)
TINY_HD void tiny_fwd_layer(const TinyLayer ly, const float* W, const float* a_in, float* a_out, int64_t ld_act, int row0,
                            int row1, int tid, int nthreads) {
    const int n = (row1 - row0) * ly.out;
    for (int e = tid; e < n; e += nthreads) {
        const int r = row0 + e / ly.out, o = e % ly.out;
        const float* a = a_in + (int64_t)r * ld_act;
        const float* w = W + (int64_t)o * ly.ld;
        float acc = 0.f;
        for (int k = 0; k <= ly.bcol; ++k) acc = fmaf(a[k], w[k], acc);
        if (ly.relu) acc = fmaxf(acc, 0.f);
        a_out[(int64_t)r * ld_act + o] = acc;
    }
}

// dW'[o, k] = sum_r g[r, o] a_in[r, k]   (k <= bcol: weights, zero padding, bias column), rows summed in order
TINY_HD void tiny_wgrad_layer(const TinyLayer ly, const float* g, int64_t ld_g, const float* a_in, int64_t ld_act, float* dW, int B,
                              int tid, int nthreads) {
    const int kk = ly.bcol + 1, n = ly.out * kk;
    for (int e = tid; e < n; e += nthreads) {
        const int o = e / kk, k = e % kk;
        float acc = 0.f;
        for (int r = 0; r < B; ++r) acc = fmaf(g[(int64_t)r * ld_g + o], a_in[(int64_t)r * ld_act + k], acc);
        dW[(int64_t)o * ly.ld + k] = acc;
    }
}

// g_prev[r, k] = (sum_o g[r, o] W'[o, k]) * (mask ? a_in[r, k] > 0 : 1),  k < in
TINY_HD void tiny_dgrad_layer(const TinyLayer ly, const float* g, const float* W, const float* a_in, int64_t ld_act, int mask,
                              float* g_prev, int64_t ld_g, int B, int tid, int nthreads) {
    const int n = B * ly.in;
    for (int e = tid; e < n; e += nthreads) {
        const int r = e / ly.in, k = e % ly.in;
        float acc = 0.f;
        for (int o = 0; o < ly.out; ++o) acc = fmaf(g[(int64_t)r * ld_g + o], W[(int64_t)o * ly.ld + k], acc);
        if (mask && !(a_in[(int64_t)r * ld_act + k] > 0.f)) acc = 0.f;
        g_prev[(int64_t)r * ld_g + k] = acc;
    }
}
