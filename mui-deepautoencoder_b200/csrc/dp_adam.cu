// K4 under data parallelism -- gradient reduce-scatter + clip_grad_norm_ + Adam on this rank's shard + all-gather of the new
// weights as ONE kernel over NVLink peer memory (no NCCL call on the data path).
//
// Replaces, per step and per rank: ncclAllReduce of the whole flat gradient buffer (4 P bytes) followed by a clip + Adam pass
// over all P parameters on every replica.  Here rank r owns the contiguous shard [r S, (r+1) S) of the flat buffers:
//   phase 0  entry barrier: every rank tells every peer "my gradients of this step are complete" (flag in the peer's signal pad)
//   phase 1  g_sum[i] = sum_q grads_q[i] for i in my shard, peers read through NVLink with 128-bit loads, summed in rank order
//            (deterministic); the sum is kept in place in my own gradient buffer; sum(g_sum^2) per CTA, fixed tree
//   phase 1b grid barrier, CTA 0 publishes this shard's sum of squares to every peer; every CTA of every rank then adds the
//            `world` values in rank order -> the SAME clip scale, bit for bit, on every rank
//   phase 2  Adam (torch 2.11 op order, optim.cu's adam_one) on the shard: master weights p, moments m, v (shard-sized) stay
//            local; the new weight is written, in the dtype the GEMMs read (bf16 shadow or f32), into EVERY rank's weight
//            buffer with peer stores -- the all-gather
//   phase 3  grid barrier, system fence, "done" flag to every peer; wait for every peer's flag: when the kernel ends this rank's
//            weight buffer is complete and no peer still reads its gradients.
// Bytes over NVLink per rank and step: (G-1)/G * 4P in + (G-1)/G * 2P out (bf16) instead of 2 (G-1)/G * 4P each way for a ring
// all-reduce; HBM traffic of the optimizer drops from 30 P to about (4 G + 26) P / G.
// Cross-GPU waits are bounded (CODAE_DP_TIMEOUT_S, default 30 s): a rank that never arrives traps instead of hanging the GPU.
//
// NVLS (peers->grads_mc / w_mc = NVSwitch multicast addresses of the same buffers, e.g. torch symmetric memory's multicast_ptr):
// phase 1 is ONE multimem.ld_reduce per 16 bytes -- the switch adds the G ranks' values and returns the sum, so a rank receives
// its shard once instead of G - 1 times -- and phase 2 ONE multimem store per 16 bytes that the switch fans out to every rank.
// Bytes on this rank's links: 4P/G in + 2P/G (bf16) out instead of (G-1)/G * 4P in + (G-1)/G * 2P out.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxGrid = 148 * 8;
constexpr int kMaxWorld = CODAE_DP_MAX_WORLD;
// signal pad layout (u64 words)
constexpr int kReady = 0, kNormFlag = 8, kNormVal = 16, kDone = 24, kEpoch = 32;

struct DpWs {
    double partial[kMaxGrid];
};

struct DpArgs {
    double beta1_d, beta2_d, lr_d;
    float w1, beta2, omb2, bc2_sqrt, neg_step, eps, wd, grad_scale, max_norm;
    int world, rank, w_bf16;
    long long n, shard;
    unsigned long long timeout_ns;
    const float* grads[kMaxWorld];
    void* w_out[kMaxWorld];
    unsigned long long* signals[kMaxWorld];
    const float* grads_mc;       // NVLS multicast address of the gradient buffers (or NULL: peer loads)
    void* w_mc;                  // NVLS multicast address of the weight buffers (or NULL: peer stores)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// NVSwitch in-fabric reduction / fan-out over a multicast address (sm_90+: LDGMC / multicast STG)
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st_f4(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st_bf16x8(void* mc, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spin until *flag >= epoch (bounded).
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long epoch, unsigned long long timeout_ns) {
    if (ld_acquire_sys(flag) >= epoch) return;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(flag) < epoch) {
        __nanosleep(64);
        if (globaltimer_ns() - t0 > timeout_ns) __trap();
    }
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    // peer (NVLink) or local gradient: plain weak load, not cached in L1 (read once)
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const DpArgs& a, float coef) {
    // identical operation order to optim.cu: adam_one (torch 2.11 single-tensor Adam)
    g = __fmul_rn(__fmul_rn(g, a.grad_scale), coef);
    if (a.wd != 0.f) g = fmaf(a.wd, p, g);
    m = fmaf(a.w1, __fsub_rn(g, m), m);
    v = __fadd_rn(__fmul_rn(v, a.beta2), __fmul_rn(__fmul_rn(a.omb2, g), g));
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), a.bc2_sqrt), a.eps);
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(a.neg_step, m), denom));
}

__global__ void __launch_bounds__(kThreads) dp_reduce_adam_gather_kernel(float* __restrict__ p, float* __restrict__ m,
                                                                         float* __restrict__ v, DpArgs a, DpWs* __restrict__ ws,
                                                                         float* __restrict__ sqnorm_out,
                                                                         const int32_t* __restrict__ step_dev) {
    namespace cg = cooperative_groups;
    __shared__ double scratch[32];
    __shared__ float s_coef;
    const int world = a.world, rank = a.rank;
    unsigned long long* my_pad = a.signals[rank];
    const unsigned long long epoch = my_pad[kEpoch] + 1;         // the last thing CTA 0 of the previous call wrote
    // ---- phase 0: my gradients are complete (stream order: this kernel is a full dependent of the backward pass) ----
    if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(a.signals[threadIdx.x] + kReady + rank, epoch);
    if (threadIdx.x < world) wait_flag(my_pad + kReady + threadIdx.x, epoch, a.timeout_ns);
    __syncthreads();
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    // ---- phase 1: reduce my shard across ranks (rank order), keep the sum in place, sum of squares ----
    const long long lo = (long long)rank * a.shard;
    const long long hi = min(a.n, lo + a.shard);
    const long long cnt4 = hi > lo ? (hi - lo) >> 2 : 0;          // shard and n are multiples of 8 (host-checked)
    const long long stride = (long long)gridDim.x * blockDim.x;
    float* gmine = const_cast<float*>(a.grads[rank]);
    {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < cnt4; e += stride) {
            const long long i = lo + 4 * e;
            float4 acc;
            if (a.grads_mc) {
                acc = multimem_ld_reduce_f4(a.grads_mc + i);          // the switch returns the sum over all ranks
            } else {
                float4 part[kMaxWorld];
#pragma unroll
                for (int q = 0; q < kMaxWorld; ++q)
                    if (q < world) part[q] = ld_peer_f4(a.grads[q] + i);
                acc = part[0];
#pragma unroll
                for (int q = 1; q < kMaxWorld; ++q)
                    if (q < world) { acc.x += part[q].x; acc.y += part[q].y; acc.z += part[q].z; acc.w += part[q].w; }
            }
            *reinterpret_cast<float4*>(gmine + i) = acc;
            s0 = fmaf(acc.x, acc.x, s0); s1 = fmaf(acc.y, acc.y, s1); s2 = fmaf(acc.z, acc.z, s2); s3 = fmaf(acc.w, acc.w, s3);
        }
        const double b = block_sum<double>((double)((s0 + s1) + (s2 + s3)), scratch);
        if (threadIdx.x == 0) ws->partial[blockIdx.x] = b;
    }
    cg::this_grid().sync();
    // ---- phase 1b: the global norm = sum over ranks of the shard sums, added in rank order on every rank ----
    if (blockIdx.x == 0) {
        double t = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(&ws->partial[i]);
        t = block_sum<double>(t, scratch);
        if (threadIdx.x == 0) {
            for (int q = 0; q < world; ++q) {
                st_relaxed_sys(a.signals[q] + kNormVal + rank, (unsigned long long)__double_as_longlong(t));
                st_release_sys(a.signals[q] + kNormFlag + rank, epoch);
            }
        }
    }
    if (threadIdx.x < world) wait_flag(my_pad + kNormFlag + threadIdx.x, epoch, a.timeout_ns);
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < world; ++q) t += __longlong_as_double((long long)ld_acquire_sys(my_pad + kNormVal + q));
        const float sq = (float)t;
        if (blockIdx.x == 0) *sqnorm_out = sq;
        const float total = __fmul_rn(sqrtf(sq), a.grad_scale);
        s_coef = a.max_norm >= 0.f ? fminf(__fdiv_rn(a.max_norm, __fadd_rn(total, 1e-6f)), 1.0f) : 1.0f;
    }
    __syncthreads();
    const float coef = s_coef;
    if (step_dev) {
        const double t = (double)(*step_dev);
        a.bc2_sqrt = (float)sqrt(1.0 - pow(a.beta2_d, t));
        a.neg_step = (float)(-(a.lr_d / (1.0 - pow(a.beta1_d, t))));
    }
    // ---- phase 2: Adam on my shard; the new weights go to every rank's weight buffer (8 elements = one 16-byte bf16 store) ----
    const long long cnt8 = cnt4 >> 1;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < cnt8; e += stride) {
        const long long i = lo + 8 * e, j = 8 * e;                 // j: index into the shard-sized moment buffers
        float4 p0 = *reinterpret_cast<const float4*>(p + i), p1 = *reinterpret_cast<const float4*>(p + i + 4);
        const float4 g0 = __ldcg(reinterpret_cast<const float4*>(gmine + i)), g1 = __ldcg(reinterpret_cast<const float4*>(gmine + i + 4));
        float4 m0 = *reinterpret_cast<const float4*>(m + j), m1 = *reinterpret_cast<const float4*>(m + j + 4);
        float4 v0 = *reinterpret_cast<const float4*>(v + j), v1 = *reinterpret_cast<const float4*>(v + j + 4);
        adam_elem(p0.x, g0.x, m0.x, v0.x, a, coef); adam_elem(p0.y, g0.y, m0.y, v0.y, a, coef);
        adam_elem(p0.z, g0.z, m0.z, v0.z, a, coef); adam_elem(p0.w, g0.w, m0.w, v0.w, a, coef);
        adam_elem(p1.x, g1.x, m1.x, v1.x, a, coef); adam_elem(p1.y, g1.y, m1.y, v1.y, a, coef);
        adam_elem(p1.z, g1.z, m1.z, v1.z, a, coef); adam_elem(p1.w, g1.w, m1.w, v1.w, a, coef);
        *reinterpret_cast<float4*>(m + j) = m0; *reinterpret_cast<float4*>(m + j + 4) = m1;
        *reinterpret_cast<float4*>(v + j) = v0; *reinterpret_cast<float4*>(v + j + 4) = v1;
        if (a.w_bf16) {
            *reinterpret_cast<float4*>(p + i) = p0; *reinterpret_cast<float4*>(p + i + 4) = p1;
            uint4 w;
            w.x = pack_bf16x2(p0.x, p0.y); w.y = pack_bf16x2(p0.z, p0.w);
            w.z = pack_bf16x2(p1.x, p1.y); w.w = pack_bf16x2(p1.z, p1.w);
            if (a.w_mc) {
                multimem_st_bf16x8(reinterpret_cast<__nv_bfloat16*>(a.w_mc) + i, w);      // one store, every rank (this one included)
            } else {
#pragma unroll
                for (int q = 0; q < kMaxWorld; ++q)
                    if (q < world) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.w_out[q]) + i) = w;
            }
        } else if (a.w_mc) {
            // f32 engine: the weight buffer IS the master buffer (w_out[rank] == p): the multicast store updates it here as well
            multimem_st_f4(reinterpret_cast<float*>(a.w_mc) + i, p0);
            multimem_st_f4(reinterpret_cast<float*>(a.w_mc) + i + 4, p1);
        } else {
#pragma unroll
            for (int q = 0; q < kMaxWorld; ++q)
                if (q < world) {
                    float* dst = reinterpret_cast<float*>(a.w_out[q]) + i;
                    *reinterpret_cast<float4*>(dst) = p0;
                    *reinterpret_cast<float4*>(dst + 4) = p1;
                }
        }
    }
    // ---- phase 3: my stores have landed everywhere; wait until everybody else's have landed here ----
    __threadfence_system();
    cg::this_grid().sync();
    if (blockIdx.x == 0) {
        if (threadIdx.x < world) {
            st_release_sys(a.signals[threadIdx.x] + kDone + rank, epoch);
            wait_flag(my_pad + kDone + threadIdx.x, epoch, a.timeout_ns);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            my_pad[kEpoch] = epoch;
            __threadfence();
        }
    }
}

}  // namespace

extern "C" {

size_t codae_dp_workspace_bytes(const codae_ctx*) { return sizeof(DpWs); }

int64_t codae_dp_shard_elems(int64_t n, int world) {
    if (n < 0 || world < 1) return 0;
    const int64_t per = (n + world - 1) / world;
    return (per + 7) / 8 * 8;
}

int codae_dp_adam_step(codae_ctx* ctx, const codae_dp_peers* peers, float* p, float* m, float* v, int w_dtype, int64_t n,
                       double lr, double beta1, double beta2, double eps, double weight_decay, int step, double max_norm,
                       float* sqnorm_out, void* workspace, size_t ws_bytes, double grad_scale, const int32_t* step_dev,
                       void* stream) {
    CODAE_REQUIRE(ctx, ctx && peers && p && m && v && sqnorm_out && workspace && n >= 0 && (step >= 1 || step_dev),
                  "codae_dp_adam_step: bad argument");
    CODAE_REQUIRE(ctx, peers->world >= 1 && peers->world <= kMaxWorld && peers->rank >= 0 && peers->rank < peers->world,
                  "codae_dp_adam_step: world %d / rank %d (at most %d ranks)", peers->world, peers->rank, kMaxWorld);
    CODAE_REQUIRE(ctx, w_dtype == CODAE_BF16 || w_dtype == CODAE_F32, "codae_dp_adam_step: bad weight dtype %d", w_dtype);
    CODAE_REQUIRE(ctx, (n % 8) == 0, "codae_dp_adam_step: n = %lld must be a multiple of 8 (the flat layout pads rows to 64)", (long long)n);
    if (ws_bytes < sizeof(DpWs)) return codae_fail(ctx, CODAE_ENOMEM, "codae_dp_adam_step: workspace %zu < %zu bytes", ws_bytes, sizeof(DpWs));
    uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v);
    for (int q = 0; q < peers->world; ++q) {
        CODAE_REQUIRE(ctx, peers->grads[q] && peers->w_out[q] && peers->signals[q], "codae_dp_adam_step: NULL peer pointer (rank %d)", q);
        al |= reinterpret_cast<uintptr_t>(peers->grads[q]) | reinterpret_cast<uintptr_t>(peers->w_out[q]);
        CODAE_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(peers->signals[q]) & 7) == 0, "codae_dp_adam_step: signal pads must be 8-byte aligned");
    }
    CODAE_REQUIRE(ctx, (al & 15) == 0, "codae_dp_adam_step: buffers must be 16-byte aligned");
    CODAE_REQUIRE(ctx, w_dtype != CODAE_F32 || peers->w_out[peers->rank] == (void*)p,
                  "codae_dp_adam_step: with f32 weights the local weight buffer must be the master buffer");
    if (step < 1) step = 1;
    if (n == 0) return CODAE_OK;
    DpArgs a;
    a.beta1_d = beta1; a.beta2_d = beta2; a.lr_d = lr;
    a.w1 = (float)(1.0 - beta1);
    a.beta2 = (float)beta2;
    a.omb2 = (float)(1.0 - beta2);
    a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    a.neg_step = (float)(-(lr / (1.0 - pow(beta1, (double)step))));
    a.eps = (float)eps;
    a.wd = (float)weight_decay;
    a.grad_scale = (float)grad_scale;
    a.max_norm = (float)max_norm;
    a.world = peers->world; a.rank = peers->rank; a.w_bf16 = w_dtype == CODAE_BF16;
    a.n = n; a.shard = codae_dp_shard_elems(n, peers->world);
    static long long timeout_s = -1;
    if (timeout_s < 0) {
        const char* e = getenv("CODAE_DP_TIMEOUT_S");
        timeout_s = (e && atoll(e) > 0) ? atoll(e) : 30;
    }
    a.timeout_ns = (unsigned long long)timeout_s * 1000000000ull;
    for (int q = 0; q < kMaxWorld; ++q) {
        const bool in = q < peers->world;
        a.grads[q] = in ? peers->grads[q] : nullptr;
        a.w_out[q] = in ? peers->w_out[q] : nullptr;
        a.signals[q] = in ? reinterpret_cast<unsigned long long*>(peers->signals[q]) : nullptr;
    }
    // NVLS when both multicast addresses are given AND the group is large enough for the byte saving to beat the lower throughput of
    // the multimem path.  Measured on B200s (tools/r02_nvls_ab.sh, r02_nvls_ab8.sh; ms/step, multimem vs peer loops):
    //   embedding.yaml bf16  2 GPUs 0.464 / 0.394   4 GPUs 0.448 / 0.426   8 GPUs 0.439 / 0.470
    //   embedding.yaml fp32  2 GPUs 0.830 / 0.647   4 GPUs 0.815 / 0.802   8 GPUs 0.784 / 0.918
    //   polyvore bf16        2 GPUs 7.99  / 7.70    4 GPUs 7.82  / 7.68    8 GPUs 7.82  / 7.89
    // CODAE_DP_NVLS = 0 | 1 forces the choice (read once).
    static const int nvls_env = getenv("CODAE_DP_NVLS") ? atoi(getenv("CODAE_DP_NVLS")) : -1;
    const bool nvls_on = nvls_env < 0 ? peers->world >= 8 : nvls_env != 0;
    const bool nvls = nvls_on && peers->grads_mc && peers->w_mc;
    CODAE_REQUIRE(ctx, !nvls || ((reinterpret_cast<uintptr_t>(peers->grads_mc) | reinterpret_cast<uintptr_t>(peers->w_mc)) & 15) == 0,
                  "codae_dp_adam_step: multicast addresses must be 16-byte aligned");
    a.grads_mc = nvls ? peers->grads_mc : nullptr;
    a.w_mc = nvls ? peers->w_mc : nullptr;
    static int max_blocks_per_sm = 0;
    if (!max_blocks_per_sm) {
        cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, dp_reduce_adam_gather_kernel, kThreads, 0);
        if (oe != cudaSuccess || max_blocks_per_sm < 1) {
            max_blocks_per_sm = 0;
            cudaGetLastError();
            return codae_fail(ctx, CODAE_ECUDA, "codae_dp_adam_step: occupancy query failed");
        }
    }
    // one co-resident wave; 4 CTAs per SM keep enough 128-bit peer loads in flight to cover the NVLink round trip
    int per_sm = max_blocks_per_sm < 4 ? max_blocks_per_sm : 4;
    long long want = ((a.shard >> 2) + kThreads - 1) / kThreads;
    int grid = (int)(want < 1 ? 1 : (want > (long long)per_sm * ctx->sm_count ? (long long)per_sm * ctx->sm_count : want));
    if (grid > kMaxGrid) grid = kMaxGrid;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, dp_reduce_adam_gather_kernel, p, m, v, a, reinterpret_cast<DpWs*>(workspace), sqnorm_out, step_dev);
    if (le != cudaSuccess) {
        cudaGetLastError();
        return codae_fail(ctx, CODAE_ECUDA, "dp_reduce_adam_gather_kernel cooperative launch (grid %d): %s", grid, cudaGetErrorString(le));
    }
    codae_mark_weights_written(ctx, as_stream(stream));
    return codae_check_launch(ctx, "dp_reduce_adam_gather_kernel");
}

}  // extern "C"
