// Shared device/host helpers for libcodae_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <mutex>

#include "../../include/codae_b200.h"

struct CUtensorMap_st;

struct codae_ctx {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    char err[512];
    void* encode_tiled;  // cuTensorMapEncodeTiled, resolved through cudaGetDriverEntryPoint
    int splitk;          // 1: small-batch contractions may use cluster split-K (default on)
    int persistent;      // 1: large contractions use the persistent, TMEM-double-buffered kernel (default on)
    int pdl;             // 1: training-step kernels are launched with programmatic dependent launch (default on)
    int weight_prefetch; // 1: fwd / dgrad GEMMs issue the TMA loads of their WEIGHT tiles before griddepcontrol.wait
    int tma_store;       // 1: single-pass f32 output tiles leave through TMA bulk stores (default on)
    int tma_store_persistent;  // 1: the persistent kernel's epilogue warps store through per-warp TMA boxes (default on)
    int cta_pair;        // 1: 256-wide persistent contractions run as CTA pairs (cta_group::2 MMAs on 256 x 256 tiles)
    int weights_dirty;   // a weight-writing kernel (Adam, clip+Adam, bf16 cast) was the last codae launch on dirty_stream
    cudaStream_t dirty_stream;
    std::mutex mu;
};

extern char g_codae_last_error[512];

int codae_fail(codae_ctx* ctx, int code, const char* fmt, ...);
int codae_check_launch(codae_ctx* ctx, const char* what);

#define CODAE_REQUIRE(ctx, cond, ...)                                     \
    do {                                                                  \
        if (!(cond)) return codae_fail((ctx), CODAE_EINVAL, __VA_ARGS__); \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: the kernel may start (prologue, barrier init, TMEM allocation) while its stream
// predecessor is still running; it must execute pdl_wait() before touching any global memory.
// Weight tiles may be fetched BEFORE griddepcontrol.wait (weight_prefetch), i.e. while any number of stream predecessors
// are still running.  That is only sound if no predecessor can still be writing weights: every entry point that writes
// weights calls codae_mark_weights_written, and the next launch on that stream is then made WITHOUT the programmatic
// attribute -- a full stream dependency (complete + flushed).  All later pre-wait portions start after that launch began.
inline void codae_mark_weights_written(codae_ctx* ctx, cudaStream_t s) {
    ctx->weights_dirty = 1;
    ctx->dirty_stream = s;
}
inline bool codae_pdl_allowed(const codae_ctx* cctx, cudaStream_t s) {
    codae_ctx* ctx = const_cast<codae_ctx*>(cctx);
    if (!ctx || !ctx->pdl) return false;
    if (ctx->weights_dirty && ctx->dirty_stream == s) {
        ctx->weights_dirty = 0;
        return false;
    }
    return true;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(const codae_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = codae_pdl_allowed(ctx, s) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// Resident CTAs per SM of a kernel (register / shared-memory limited): grid-stride kernels are launched as ONE resident wave,
// a larger grid only adds a partial last wave.
template <typename K>
inline int resident_ctas_per_sm(K kernel, int threads, size_t smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = 1;
    }
    return n;
}

// ---- device helpers ---------------------------------------------------------------------------
// No-ops when the grid was not launched as a programmatic dependent.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream_u2(void* p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// CODAE_F32X3 -- three-way bf16 split of an fp32 value: x = hi + mid + lo with |x - (hi + mid + lo)| <= 2^-24 |x| (both
// residuals are exact in fp32).  The fp32-parity tensor-core engine (gemm_tcgen05.cu, NP = 3) contracts such triples.
__device__ __forceinline__ void split3(float x, float& h, float& m, float& l) {
    h = __bfloat162float(__float2bfloat16_rn(x));
    const float r1 = __fsub_rn(x, h);
    m = __bfloat162float(__float2bfloat16_rn(r1));
    l = __fsub_rn(r1, m);
}
// Four consecutive fp32 values -> the same four columns of the three bf16 planes (8-byte stores; `dst` = plane 0).
__device__ __forceinline__ void store_planes4(__nv_bfloat16* dst, int64_t plane_stride, float4 x) {
    float h[4], m[4], l[4];
    split3(x.x, h[0], m[0], l[0]);
    split3(x.y, h[1], m[1], l[1]);
    split3(x.z, h[2], m[2], l[2]);
    split3(x.w, h[3], m[3], l[3]);
    uint2 q;
    q.x = pack_bf16x2(h[0], h[1]); q.y = pack_bf16x2(h[2], h[3]);
    *reinterpret_cast<uint2*>(dst) = q;
    q.x = pack_bf16x2(m[0], m[1]); q.y = pack_bf16x2(m[2], m[3]);
    *reinterpret_cast<uint2*>(dst + plane_stride) = q;
    q.x = pack_bf16x2(l[0], l[1]); q.y = pack_bf16x2(l[2], l[3]);
    *reinterpret_cast<uint2*>(dst + 2 * plane_stride) = q;
}
__device__ __forceinline__ void store_planes1(__nv_bfloat16* dst, int64_t plane_stride, float x) {
    float h, m, l;
    split3(x, h, m, l);
    dst[0] = __float2bfloat16_rn(h);
    dst[plane_stride] = __float2bfloat16_rn(m);
    dst[2 * plane_stride] = __float2bfloat16_rn(l);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum with a fixed tree (warp shuffles, then warp 0 over the per-warp partials).
// Result valid in thread 0.  `scratch` holds >= 32 elements of T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    T r = T(0);
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? scratch[lane] : T(0);
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}
