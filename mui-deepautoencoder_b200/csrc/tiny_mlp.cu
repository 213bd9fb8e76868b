// K2 for tabular widths -- the whole network in one launch (codae_tiny_mlp_step, "io < 64 SIMT whole-net").
// abalone.yaml's model is 6 Linear layers of at most 11 x 11: as per-layer launches its forward and backward passes are
// 17 kernels of a few hundred FMAs each.  Here: codae_tiny_mlp_fwd = every layer of the forward pass (rows are independent:
// one CTA per 32 rows, no inter-CTA dependency), codae_tiny_mlp_bwd = every weight gradient and the input-gradient chain in
// ONE CTA (weight gradients sum over the batch: fixed row order, deterministic).  Exact fp32 FMA arithmetic (tiny_mlp.h is
// shared with the CPU unit test).  OPT-IN (FusedStep(tiny_mlp=True)): not yet run on a B200.
#include "common.cuh"
#include "tiny_mlp.h"

namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerCta = 32;

struct TinyArgs {
    TinyLayer layer[CODAE_TINY_MAX_LAYERS];
    float* act[CODAE_TINY_MAX_LAYERS + 1];
    float* g[3];                  // dL/d(out_l) lives in g[l % 3] (backward only)
    int L, B;
    int64_t ld_act, ld_g;
};

__global__ void __launch_bounds__(kThreads) tiny_mlp_fwd_kernel(const TinyArgs a, const float* __restrict__ flat) {
    const int row0 = blockIdx.x * kRowsPerCta, row1 = min(a.B, row0 + kRowsPerCta);
    for (int l = 0; l < a.L; ++l) {
        tiny_fwd_layer(a.layer[l], flat + a.layer[l].w_off, a.act[l], a.act[l + 1], a.ld_act, row0, row1, threadIdx.x, blockDim.x);
        __syncthreads();          // this CTA's rows of layer l are written (global memory, same CTA) before layer l+1 reads them
    }
}

__global__ void __launch_bounds__(kThreads) tiny_mlp_bwd_kernel(const TinyArgs a, const float* __restrict__ flat,
                                                                float* __restrict__ gflat) {
    for (int l = a.L - 1; l >= 0; --l) {
        const TinyLayer ly = a.layer[l];
        const float* g = a.g[l % 3];
        tiny_wgrad_layer(ly, g, a.ld_g, a.act[l], a.ld_act, gflat + ly.w_off, a.B, threadIdx.x, blockDim.x);
        if (l > 0)
            tiny_dgrad_layer(ly, g, flat + ly.w_off, a.act[l], a.ld_act, a.layer[l - 1].relu, a.g[(l - 1) % 3], a.ld_g, a.B,
                             threadIdx.x, blockDim.x);
        __syncthreads();
    }
}

int fill(codae_ctx* ctx, TinyArgs& a, const codae_tiny_layer* layers, int n_layers, float* const* acts, int64_t ld_act, int B,
         const char* who) {
    CODAE_REQUIRE(ctx, layers && acts && n_layers >= 1 && n_layers <= CODAE_TINY_MAX_LAYERS, "%s: 1..%d layers", who, CODAE_TINY_MAX_LAYERS);
    CODAE_REQUIRE(ctx, B >= 1 && ld_act >= 1, "%s: bad shape", who);
    memset(&a, 0, sizeof(a));
    for (int l = 0; l < n_layers; ++l) {
        const codae_tiny_layer& s = layers[l];
        CODAE_REQUIRE(ctx, s.in >= 1 && s.out >= 1 && s.bcol >= s.in && s.ld > s.bcol && s.w_off >= 0 && ld_act > s.bcol && ld_act >= s.out,
                      "%s: layer %d: bad layout", who, l);
        CODAE_REQUIRE(ctx, acts[l], "%s: activation buffer %d is NULL", who, l);
        a.layer[l].w_off = s.w_off; a.layer[l].ld = s.ld; a.layer[l].bcol = s.bcol;
        a.layer[l].in = s.in; a.layer[l].out = s.out; a.layer[l].relu = s.relu ? 1 : 0;
        a.act[l] = acts[l];
    }
    CODAE_REQUIRE(ctx, acts[n_layers], "%s: activation buffer %d is NULL", who, n_layers);
    a.act[n_layers] = acts[n_layers];
    a.L = n_layers; a.B = B; a.ld_act = ld_act;
    return CODAE_OK;
}

}  // namespace

extern "C" {

int codae_tiny_mlp_fwd(codae_ctx* ctx, const codae_tiny_layer* layers, int n_layers, const float* flat, float* const* acts,
                       int64_t ld_act, int B, void* stream) {
    CODAE_REQUIRE(ctx, ctx && flat, "codae_tiny_mlp_fwd: NULL argument");
    TinyArgs a;
    int rc = fill(ctx, a, layers, n_layers, acts, ld_act, B, "codae_tiny_mlp_fwd");
    if (rc) return rc;
    tiny_mlp_fwd_kernel<<<(B + kRowsPerCta - 1) / kRowsPerCta, kThreads, 0, as_stream(stream)>>>(a, flat);
    return codae_check_launch(ctx, "tiny_mlp_fwd_kernel");
}

int codae_tiny_mlp_bwd(codae_ctx* ctx, const codae_tiny_layer* layers, int n_layers, const float* flat, float* gflat,
                       float* const* acts, int64_t ld_act, float* const* g3, int64_t ld_g, int B, void* stream) {
    CODAE_REQUIRE(ctx, ctx && flat && gflat && g3 && g3[0] && g3[1] && g3[2], "codae_tiny_mlp_bwd: NULL argument");
    TinyArgs a;
    int rc = fill(ctx, a, layers, n_layers, acts, ld_act, B, "codae_tiny_mlp_bwd");
    if (rc) return rc;
    for (int l = 0; l < n_layers; ++l)
        CODAE_REQUIRE(ctx, ld_g >= layers[l].out && ld_g >= layers[l].in, "codae_tiny_mlp_bwd: gradient pitch < layer width");
    a.g[0] = g3[0]; a.g[1] = g3[1]; a.g[2] = g3[2];
    a.ld_g = ld_g;
    tiny_mlp_bwd_kernel<<<1, kThreads, 0, as_stream(stream)>>>(a, flat, gflat);
    return codae_check_launch(ctx, "tiny_mlp_bwd_kernel");
}

}  // extern "C"
