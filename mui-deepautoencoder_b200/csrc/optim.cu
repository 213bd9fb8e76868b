// K4 -- clip_grad_norm_ + Adam over the flat parameter / gradient / moment buffers.
// HBM-bound: 4 B/param for the norm pass, 28 B/param (+2 with a bf16 shadow) for the update.
// The arithmetic follows torch 2.11's single-tensor Adam operation by operation (no re-association,
// explicit _rn intrinsics where ATen does not fuse) so that weights track the reference to fp32 rounding.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxGrid = 148 * 8 * 2;

struct NormWs {
    unsigned int ticket;
    unsigned int pad[3];
    double partial[kMaxGrid];
};

__global__ void __launch_bounds__(kThreads) sqnorm_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out,
                                                          NormWs* __restrict__ ws) {
    pdl_launch_dependents();
    pdl_wait();
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // two independent 128-bit loads in flight per iteration
    for (; e + stride < n4; e += 2 * stride) {
        const float4 a = ldg_stream_f4(g + 4 * e);
        const float4 b = ldg_stream_f4(g + 4 * (e + stride));
        s0 = fmaf(a.x, a.x, s0); s1 = fmaf(a.y, a.y, s1); s2 = fmaf(a.z, a.z, s2); s3 = fmaf(a.w, a.w, s3);
        s0 = fmaf(b.x, b.x, s0); s1 = fmaf(b.y, b.y, s1); s2 = fmaf(b.z, b.z, s2); s3 = fmaf(b.w, b.w, s3);
    }
    for (; e < n4; e += stride) {
        const float4 a = ldg_stream_f4(g + 4 * e);
        s0 = fmaf(a.x, a.x, s0); s1 = fmaf(a.y, a.y, s1); s2 = fmaf(a.z, a.z, s2); s3 = fmaf(a.w, a.w, s3);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const float a = g[(n4 << 2) + threadIdx.x];
        s0 = fmaf(a, a, s0);
    }
    __shared__ double scratch[32];
    __shared__ bool is_last;
    const double b = block_sum<double>((double)((s0 + s1) + (s2 + s3)), scratch);
    if (threadIdx.x == 0) {
        ws->partial[blockIdx.x] = b;
        __threadfence();
        is_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(&ws->partial[i]);
    t = block_sum<double>(t, scratch);
    if (threadIdx.x == 0) {
        *out = (float)t;
        ws->ticket = 0;
    }
}

struct AdamArgs {
    double beta1_d, beta2_d, lr_d;   // for the device-side step counter path
    float w1;          // 1 - beta1
    float beta2;       // beta2
    float omb2;        // 1 - beta2
    float bc2_sqrt;    // sqrt(1 - beta2^t)
    float neg_step;    // -lr / (1 - beta1^t)
    float eps;
    float wd;
    float grad_scale;
    float max_norm;    // < 0: no clipping
};

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, const AdamArgs& a, float coef) {
    // clip (torch: g.mul_(clip_coef_clamped)), then L2 decay g + wd*p (ATen add(alpha) = fmadd)
    g = __fmul_rn(__fmul_rn(g, a.grad_scale), coef);
    if (a.wd != 0.f) g = fmaf(a.wd, p, g);
    m = fmaf(a.w1, __fsub_rn(g, m), m);                                                 // m.lerp_(g, 1-b1)
    v = __fadd_rn(__fmul_rn(v, a.beta2), __fmul_rn(__fmul_rn(a.omb2, g), g));           // v.mul_(b2).addcmul_(g,g,1-b2)
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), a.bc2_sqrt), a.eps);              // sqrt(v)/bc2_sqrt + eps
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(a.neg_step, m), denom));                       // p.addcdiv_(m, denom, -step)
    return p;
}

// Weight copy the GEMMs read, rewritten with the update: one bf16 plane (pb_plane == 0, tensor-core engine) or the three
// planes of the fp32-parity engine (CODAE_F32X3, planes pb_plane elements apart).
__device__ __forceinline__ void store_shadow4(__nv_bfloat16* pb, int64_t pb_plane, int64_t e, const float4& pv) {
    if (pb_plane) {
        store_planes4(pb + 4 * e, pb_plane, pv);
    } else {
        uint2 q;
        q.x = pack_bf16x2(pv.x, pv.y);
        q.y = pack_bf16x2(pv.z, pv.w);
        *reinterpret_cast<uint2*>(pb + 4 * e) = q;
    }
}
__device__ __forceinline__ void store_shadow1(__nv_bfloat16* pb, int64_t pb_plane, int64_t i, float pv) {
    if (pb_plane) store_planes1(pb + i, pb_plane, pv);
    else pb[i] = __float2bfloat16_rn(pv);
}

__global__ void __launch_bounds__(kThreads) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        __nv_bfloat16* __restrict__ pb, int64_t pb_plane, int64_t n, AdamArgs a,
                                                        const float* __restrict__ sqnorm,
                                                        const int32_t* __restrict__ step_dev) {
    pdl_launch_dependents();
    pdl_wait();
    if (step_dev) {
        // CUDA-graph replays cannot change scalar arguments: derive the bias corrections from a device counter,
        // in double like the host path.
        const double t = (double)(*step_dev);
        a.bc2_sqrt = (float)sqrt(1.0 - pow(a.beta2_d, t));
        a.neg_step = (float)(-(a.lr_d / (1.0 - pow(a.beta1_d, t))));
    }
    float coef = 1.0f;
    if (a.max_norm >= 0.f && sqnorm) {
        const float total = __fmul_rn(sqrtf(*sqnorm), a.grad_scale);
        coef = fminf(__fdiv_rn(a.max_norm, __fadd_rn(total, 1e-6f)), 1.0f);
    }
    const int64_t n4 = n >> 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
        float4 pv = *reinterpret_cast<const float4*>(p + 4 * e);
        const float4 gv = ldg_stream_f4(g + 4 * e);
        float4 mv = *reinterpret_cast<const float4*>(m + 4 * e);
        float4 vv = *reinterpret_cast<const float4*>(v + 4 * e);
        adam_one(pv.x, gv.x, mv.x, vv.x, a, coef);
        adam_one(pv.y, gv.y, mv.y, vv.y, a, coef);
        adam_one(pv.z, gv.z, mv.z, vv.z, a, coef);
        adam_one(pv.w, gv.w, mv.w, vv.w, a, coef);
        *reinterpret_cast<float4*>(p + 4 * e) = pv;
        *reinterpret_cast<float4*>(m + 4 * e) = mv;
        *reinterpret_cast<float4*>(v + 4 * e) = vv;
        if (pb) store_shadow4(pb, pb_plane, e, pv);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        float pv = p[i], mv = m[i], vv = v[i];
        adam_one(pv, g[i], mv, vv, a, coef);
        p[i] = pv; m[i] = mv; v[i] = vv;
        if (pb) store_shadow1(pb, pb_plane, i, pv);
    }
}

// Adam with the clip scale derived from per-CTA sums of squares that the weight-gradient kernels left behind
// (codae_linear_wgrad_sq): no norm pass over g at all.  Every CTA sums the (few thousand, L2-resident) partials in the
// same fixed order, so all CTAs use the same scale and the result is bitwise reproducible.
__global__ void __launch_bounds__(kThreads) adam_partials_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                                 float* __restrict__ m, float* __restrict__ v,
                                                                 __nv_bfloat16* __restrict__ pb, int64_t pb_plane, int64_t n, AdamArgs a,
                                                                 const double* __restrict__ partials, int n_partials,
                                                                 float* __restrict__ sqnorm_out,
                                                                 const int32_t* __restrict__ step_dev) {
    __shared__ double scratch[32];
    __shared__ float s_coef;
    pdl_launch_dependents();
    pdl_wait();
    {
        double t = 0.0;
        for (int i = threadIdx.x; i < n_partials; i += blockDim.x) t += __ldcg(&partials[i]);
        t = block_sum<double>(t, scratch);
        if (threadIdx.x == 0) {
            const float sq = (float)t;
            if (blockIdx.x == 0) *sqnorm_out = sq;
            const float total = __fmul_rn(sqrtf(sq), a.grad_scale);
            s_coef = a.max_norm >= 0.f ? fminf(__fdiv_rn(a.max_norm, __fadd_rn(total, 1e-6f)), 1.0f) : 1.0f;
        }
        __syncthreads();
    }
    const float coef = s_coef;
    if (step_dev) {
        const double t = (double)(*step_dev);
        a.bc2_sqrt = (float)sqrt(1.0 - pow(a.beta2_d, t));
        a.neg_step = (float)(-(a.lr_d / (1.0 - pow(a.beta1_d, t))));
    }
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += stride) {
        float4 pv = *reinterpret_cast<const float4*>(p + 4 * e);
        const float4 gv = __ldcg(reinterpret_cast<const float4*>(g) + e);
        float4 mv = *reinterpret_cast<const float4*>(m + 4 * e);
        float4 vv = *reinterpret_cast<const float4*>(v + 4 * e);
        adam_one(pv.x, gv.x, mv.x, vv.x, a, coef);
        adam_one(pv.y, gv.y, mv.y, vv.y, a, coef);
        adam_one(pv.z, gv.z, mv.z, vv.z, a, coef);
        adam_one(pv.w, gv.w, mv.w, vv.w, a, coef);
        *reinterpret_cast<float4*>(p + 4 * e) = pv;
        *reinterpret_cast<float4*>(m + 4 * e) = mv;
        *reinterpret_cast<float4*>(v + 4 * e) = vv;
        if (pb) store_shadow4(pb, pb_plane, e, pv);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        float pv = p[i], mv = m[i], vv = v[i];
        adam_one(pv, g[i], mv, vv, a, coef);
        p[i] = pv; m[i] = mv; v[i] = vv;
        if (pb) store_shadow1(pb, pb_plane, i, pv);
    }
}

// clip_grad_norm_ + Adam in ONE cooperative launch: phase 1 reduces ||g||^2 (per-CTA partials, then every CTA sums the
// partials in the same fixed order), a grid barrier, phase 2 applies the update.  g (4 B/param) is read twice, but the
// second read is served by the 126 MB L2 for the shipped configs (94 MB of gradients), and one launch disappears.
__global__ void __launch_bounds__(kThreads) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                             float* __restrict__ m, float* __restrict__ v,
                                                             __nv_bfloat16* __restrict__ pb, int64_t pb_plane, int64_t n, AdamArgs a,
                                                             NormWs* __restrict__ ws, float* __restrict__ sqnorm_out,
                                                             const int32_t* __restrict__ step_dev) {
    namespace cg = cooperative_groups;
    __shared__ double scratch[32];
    __shared__ float s_coef;
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; e + stride < n4; e += 2 * stride) {
            const float4 x = __ldcg(reinterpret_cast<const float4*>(g) + e);
            const float4 y = __ldcg(reinterpret_cast<const float4*>(g) + e + stride);
            s0 = fmaf(x.x, x.x, s0); s1 = fmaf(x.y, x.y, s1); s2 = fmaf(x.z, x.z, s2); s3 = fmaf(x.w, x.w, s3);
            s0 = fmaf(y.x, y.x, s0); s1 = fmaf(y.y, y.y, s1); s2 = fmaf(y.z, y.z, s2); s3 = fmaf(y.w, y.w, s3);
        }
        for (; e < n4; e += stride) {
            const float4 x = __ldcg(reinterpret_cast<const float4*>(g) + e);
            s0 = fmaf(x.x, x.x, s0); s1 = fmaf(x.y, x.y, s1); s2 = fmaf(x.z, x.z, s2); s3 = fmaf(x.w, x.w, s3);
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
            const float x = g[(n4 << 2) + threadIdx.x];
            s0 = fmaf(x, x, s0);
        }
        const double b = block_sum<double>((double)((s0 + s1) + (s2 + s3)), scratch);
        if (threadIdx.x == 0) ws->partial[blockIdx.x] = b;
    }
    cg::this_grid().sync();
    {
        double t = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(&ws->partial[i]);
        t = block_sum<double>(t, scratch);
        if (threadIdx.x == 0) {
            const float sq = (float)t;
            if (blockIdx.x == 0) *sqnorm_out = sq;
            const float total = __fmul_rn(sqrtf(sq), a.grad_scale);
            s_coef = a.max_norm >= 0.f ? fminf(__fdiv_rn(a.max_norm, __fadd_rn(total, 1e-6f)), 1.0f) : 1.0f;
        }
        __syncthreads();
    }
    const float coef = s_coef;
    if (step_dev) {
        const double t = (double)(*step_dev);
        a.bc2_sqrt = (float)sqrt(1.0 - pow(a.beta2_d, t));
        a.neg_step = (float)(-(a.lr_d / (1.0 - pow(a.beta1_d, t))));
    }
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += stride) {
        float4 pv = *reinterpret_cast<const float4*>(p + 4 * e);
        const float4 gv = __ldcg(reinterpret_cast<const float4*>(g) + e);
        float4 mv = *reinterpret_cast<const float4*>(m + 4 * e);
        float4 vv = *reinterpret_cast<const float4*>(v + 4 * e);
        adam_one(pv.x, gv.x, mv.x, vv.x, a, coef);
        adam_one(pv.y, gv.y, mv.y, vv.y, a, coef);
        adam_one(pv.z, gv.z, mv.z, vv.z, a, coef);
        adam_one(pv.w, gv.w, mv.w, vv.w, a, coef);
        *reinterpret_cast<float4*>(p + 4 * e) = pv;
        *reinterpret_cast<float4*>(m + 4 * e) = mv;
        *reinterpret_cast<float4*>(v + 4 * e) = vv;
        if (pb) store_shadow4(pb, pb_plane, e, pv);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        float pv = p[i], mv = m[i], vv = v[i];
        adam_one(pv, g[i], mv, vv, a, coef);
        p[i] = pv; m[i] = mv; v[i] = vv;
        if (pb) store_shadow1(pb, pb_plane, i, pv);
    }
}

__global__ void __launch_bounds__(kThreads) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                             int64_t n) {
    const int64_t n4 = n >> 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = ldg_stream_f4(src + 4 * e);
        uint2 q;
        q.x = pack_bf16x2(a.x, a.y);
        q.y = pack_bf16x2(a.z, a.w);
        *reinterpret_cast<uint2*>(dst + 4 * e) = q;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        dst[i] = __float2bfloat16_rn(src[i]);
    }
}

__global__ void __launch_bounds__(kThreads) split_x3_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                            int64_t n4, int64_t plane) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x)
        store_planes4(dst + 4 * e, plane, ldg_stream_f4(src + 4 * e));
}

__global__ void counter_add_kernel(int32_t* c, int delta) {
    pdl_launch_dependents();
    pdl_wait();
    *c += delta;
}

// scalar bookkeeping in double, exactly as torch/optim/adam.py does in Python floats
inline AdamArgs make_adam_args(double lr, double beta1, double beta2, double eps, double weight_decay, int step, double max_norm,
                               double grad_scale) {
    AdamArgs a;
    a.beta1_d = beta1; a.beta2_d = beta2; a.lr_d = lr;
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    a.w1 = (float)(1.0 - beta1);
    a.beta2 = (float)beta2;
    a.omb2 = (float)(1.0 - beta2);
    a.bc2_sqrt = (float)sqrt(bc2);
    a.neg_step = (float)(-(lr / bc1));
    a.eps = (float)eps;
    a.wd = (float)weight_decay;
    a.grad_scale = (float)grad_scale;
    a.max_norm = (float)max_norm;
    return a;
}

// The optimizer kernels are launched with a FULL stream dependency, not as programmatic dependents: a dependent grid of
// 592 CTAs becomes resident as soon as its predecessor starts and then sits in griddepcontrol.wait on every SM while the
// last backward GEMMs (main and side stream) still need those SMs.  Measured on the embedding.yaml step (B200):
// 0.3754 ms/step with the programmatic launch, 0.3367 without.  CODAE_ADAM_PDL=1 restores it for A/B runs (read once).
inline bool adam_uses_pdl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CODAE_ADAM_PDL");
        v = e ? (atoi(e) != 0) : 0;
    }
    return v != 0;
}

inline int grid_for(const codae_ctx* ctx, int64_t n4, int per_thread) {
    int64_t blocks = (n4 + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks > kMaxGrid) blocks = kMaxGrid;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

extern "C" {

size_t codae_sqnorm_workspace_bytes(const codae_ctx*) { return sizeof(NormWs); }

int codae_grad_sqnorm(codae_ctx* ctx, const float* g, int64_t n, float* out_sqnorm, void* workspace, size_t ws_bytes,
                      void* stream) {
    CODAE_REQUIRE(ctx, ctx && g && out_sqnorm && workspace && n >= 0, "codae_grad_sqnorm: bad argument");
    CODAE_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(g) & 15) == 0, "codae_grad_sqnorm: g must be 16-byte aligned");
    if (ws_bytes < sizeof(NormWs))
        return codae_fail(ctx, CODAE_ENOMEM, "codae_grad_sqnorm: workspace %zu < %zu bytes", ws_bytes, sizeof(NormWs));
    launch_pdl(ctx, sqnorm_kernel, dim3(grid_for(ctx, n >> 2, 4)), dim3(kThreads), 0, as_stream(stream), g, n, out_sqnorm,
               reinterpret_cast<NormWs*>(workspace));
    return codae_check_launch(ctx, "sqnorm_kernel");
}

int codae_adam_step(codae_ctx* ctx, float* p, const float* g, float* m, float* v, void* p_shadow, int shadow_dtype, int64_t n, double lr,
                    double beta1, double beta2, double eps, double weight_decay, int step, double max_norm,
                    const float* sqnorm, double grad_scale, const int32_t* step_dev, void* stream) {
    CODAE_REQUIRE(ctx, ctx && p && g && m && v && n >= 0 && (step >= 1 || step_dev), "codae_adam_step: bad argument");
    if (step < 1) step = 1;
    CODAE_REQUIRE(ctx, !p_shadow || shadow_dtype == CODAE_BF16 || (shadow_dtype == CODAE_F32X3 && n % 8 == 0),
                  "weight shadow: CODAE_BF16, or CODAE_F32X3 ([3][n] planes, n a multiple of 8)");
    const int64_t pb_plane = (p_shadow && shadow_dtype == CODAE_F32X3) ? n : 0;
    CODAE_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(p_shadow) & 7) == 0,
                  "codae_adam_step: buffers must be 16-byte aligned");
    CODAE_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(p_shadow) & 15) == 0, "codae_adam_step: p_shadow must be 16-byte aligned");
    const AdamArgs a = make_adam_args(lr, beta1, beta2, eps, weight_decay, step, max_norm, grad_scale);
    if (n == 0) return CODAE_OK;
    // one resident wave: 4 CTAs per SM measured best for this access pattern (more CTAs per SM: 156 vs 138 us on 23.6 M
    // parameters; a grid of 8 per SM does not fit at once and leaves a partial second wave)
    int grid = grid_for(ctx, n >> 2, 2);
    if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
    if (adam_uses_pdl()) {
        launch_pdl(ctx, adam_kernel, dim3(grid), dim3(kThreads), 0, as_stream(stream), p, (const float*)g, m, v,
                   reinterpret_cast<__nv_bfloat16*>(p_shadow), pb_plane, n, a, sqnorm, step_dev);
    } else {
        adam_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(p, (const float*)g, m, v, reinterpret_cast<__nv_bfloat16*>(p_shadow), pb_plane,
                                                             n, a, sqnorm, step_dev);
    }
    codae_mark_weights_written(ctx, as_stream(stream));
    return codae_check_launch(ctx, "adam_kernel");
}

int codae_clip_adam_step(codae_ctx* ctx, float* p, const float* g, float* m, float* v, void* p_shadow, int shadow_dtype, int64_t n, double lr,
                         double beta1, double beta2, double eps, double weight_decay, int step, double max_norm,
                         float* sqnorm_out, void* workspace, size_t ws_bytes, double grad_scale, const int32_t* step_dev,
                         void* stream) {
    CODAE_REQUIRE(ctx, ctx && p && g && m && v && sqnorm_out && workspace && n >= 0 && (step >= 1 || step_dev),
                  "codae_clip_adam_step: bad argument");
    CODAE_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(p_shadow) & 7) == 0,
                  "codae_clip_adam_step: buffers must be 16-byte aligned");
    if (ws_bytes < sizeof(NormWs))
        return codae_fail(ctx, CODAE_ENOMEM, "codae_clip_adam_step: workspace %zu < %zu bytes", ws_bytes, sizeof(NormWs));
    if (step < 1) step = 1;
    CODAE_REQUIRE(ctx, !p_shadow || shadow_dtype == CODAE_BF16 || (shadow_dtype == CODAE_F32X3 && n % 8 == 0),
                  "weight shadow: CODAE_BF16, or CODAE_F32X3 ([3][n] planes, n a multiple of 8)");
    const int64_t pb_plane = (p_shadow && shadow_dtype == CODAE_F32X3) ? n : 0;
    const AdamArgs a = make_adam_args(lr, beta1, beta2, eps, weight_decay, step, max_norm, grad_scale);
    if (n == 0) return CODAE_OK;
    // cooperative launch: the grid must be co-resident
    static int max_blocks_per_sm = 0;
    if (!max_blocks_per_sm) {
        cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, clip_adam_kernel, kThreads, 0);
        if (oe != cudaSuccess || max_blocks_per_sm < 1) {
            max_blocks_per_sm = 0;
            cudaGetLastError();
            return codae_fail(ctx, CODAE_ECUDA, "codae_clip_adam_step: occupancy query failed");
        }
    }
    int grid = grid_for(ctx, n >> 2, 2);
    if (grid > max_blocks_per_sm * ctx->sm_count) grid = max_blocks_per_sm * ctx->sm_count;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, clip_adam_kernel, p, g, m, v, reinterpret_cast<__nv_bfloat16*>(p_shadow), pb_plane, n, a,
                                        reinterpret_cast<NormWs*>(workspace), sqnorm_out, step_dev);
    if (le != cudaSuccess) {
        cudaGetLastError();
        return codae_fail(ctx, CODAE_ECUDA, "clip_adam_kernel cooperative launch (grid %d): %s", grid, cudaGetErrorString(le));
    }
    codae_mark_weights_written(ctx, as_stream(stream));
    return codae_check_launch(ctx, "clip_adam_kernel");
}

int codae_adam_step_partials(codae_ctx* ctx, float* p, const float* g, float* m, float* v, void* p_shadow, int shadow_dtype, int64_t n, double lr,
                             double beta1, double beta2, double eps, double weight_decay, int step, double max_norm,
                             const double* sq_partials, int n_partials, float* sqnorm_out, double grad_scale,
                             const int32_t* step_dev, void* stream) {
    CODAE_REQUIRE(ctx, ctx && p && g && m && v && sq_partials && sqnorm_out && n >= 0 && n_partials >= 1 && (step >= 1 || step_dev),
                  "codae_adam_step_partials: bad argument");
    CODAE_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(p_shadow)) & 15) == 0 &&
                            (reinterpret_cast<uintptr_t>(sq_partials) & 7) == 0,
                  "codae_adam_step_partials: buffers must be 16-byte aligned (sq_partials: 8)");
    if (step < 1) step = 1;
    CODAE_REQUIRE(ctx, !p_shadow || shadow_dtype == CODAE_BF16 || (shadow_dtype == CODAE_F32X3 && n % 8 == 0),
                  "weight shadow: CODAE_BF16, or CODAE_F32X3 ([3][n] planes, n a multiple of 8)");
    const int64_t pb_plane = (p_shadow && shadow_dtype == CODAE_F32X3) ? n : 0;
    const AdamArgs a = make_adam_args(lr, beta1, beta2, eps, weight_decay, step, max_norm, grad_scale);
    if (n == 0) return CODAE_OK;
    int grid = grid_for(ctx, n >> 2, 2);
    if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;       // one resident wave, as codae_adam_step
    cudaError_t le;
    if (adam_uses_pdl()) {
        le = launch_pdl(ctx, adam_partials_kernel, dim3(grid), dim3(kThreads), 0, as_stream(stream), p, (const float*)g, m, v,
                        reinterpret_cast<__nv_bfloat16*>(p_shadow), pb_plane, n, a, sq_partials, n_partials, sqnorm_out, step_dev);
    } else {
        adam_partials_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(p, (const float*)g, m, v, reinterpret_cast<__nv_bfloat16*>(p_shadow),
                                                                      pb_plane, n, a, sq_partials, n_partials, sqnorm_out, step_dev);
        le = cudaSuccess;
    }
    if (le != cudaSuccess) {
        cudaGetLastError();
        return codae_fail(ctx, CODAE_ECUDA, "adam_partials_kernel launch: %s", cudaGetErrorString(le));
    }
    codae_mark_weights_written(ctx, as_stream(stream));
    return codae_check_launch(ctx, "adam_partials_kernel");
}

int codae_counter_add(codae_ctx* ctx, int32_t* counter, int delta, void* stream) {
    CODAE_REQUIRE(ctx, ctx && counter, "codae_counter_add: NULL argument");
    launch_pdl(ctx, counter_add_kernel, dim3(1), dim3(1), 0, as_stream(stream), counter, delta);
    return codae_check_launch(ctx, "counter_add_kernel");
}

int codae_cast_bf16(codae_ctx* ctx, const float* src, void* dst, int64_t n, void* stream) {
    CODAE_REQUIRE(ctx, ctx && src && dst && n >= 0, "codae_cast_bf16: bad argument");
    CODAE_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
                  "codae_cast_bf16: misaligned buffers");
    if (n == 0) return CODAE_OK;
    cast_bf16_kernel<<<grid_for(ctx, n >> 2, 4), kThreads, 0, as_stream(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
    codae_mark_weights_written(ctx, as_stream(stream));
    return codae_check_launch(ctx, "cast_bf16_kernel");
}

int codae_split_x3(codae_ctx* ctx, const float* src, void* dst, int64_t n, int64_t plane_stride, void* stream) {
    CODAE_REQUIRE(ctx, ctx && src && dst && n >= 0 && plane_stride >= n, "codae_split_x3: bad argument");
    CODAE_REQUIRE(ctx, (n % 4) == 0 && (plane_stride % 4) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                           (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
                  "codae_split_x3: n and plane_stride must be multiples of 4, src 16-byte and dst 8-byte aligned");
    if (n == 0) return CODAE_OK;
    split_x3_kernel<<<grid_for(ctx, n >> 2, 4), kThreads, 0, as_stream(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n >> 2,
                                                                                 plane_stride);
    codae_mark_weights_written(ctx, as_stream(stream));
    return codae_check_launch(ctx, "split_x3_kernel");
}

}  // extern "C"
