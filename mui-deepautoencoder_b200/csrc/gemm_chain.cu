// K2, small batches -- a whole CHAIN of dependent Linear layers in one persistent launch (opt-in, codae_linear_chain).
//
// Why: at B <= 128 every layer of embedding.yaml is one 128-row tile deep; as separate launches each layer costs ~7 us
// (launch hand-over, first operand fetch, MMAs, cluster reduce, store, teardown -- profiles/r01_notes.md has the timeline)
// although its data moves in < 1 us.  Here ONE grid of C clusters x S CTAs stays resident for all layers:
//   * layer l, output tile t (128 x 64) belongs to cluster t mod C; the S CTAs of the cluster split its k-blocks (split-K)
//     and reduce the partial tiles through distributed shared memory in rank order, exactly like tc05_gemm_kernel;
//   * WEIGHT tiles of the next (layer, tile) are requested while the current one is still being reduced -- they do not
//     depend on anything computed here -- so only the 128 x K activation slice sits on the critical path of a layer;
//   * layers are separated by a grid-wide barrier on a global counter (one arrival per CTA and layer, release/acquire at
//     GPU scope, generic->async proxy fences on both sides because the next layer reads the activations through TMA).
// The grid must be co-resident (it spins on the counters): the host sizes it with cudaOccupancyMaxActiveClusters and the
// waits are bounded (a protocol bug traps instead of hanging the GPU).
// Same arithmetic as the per-layer path: same tiles, same k order, partials summed in rank order (bitwise reproducible).
#include <cuda.h>

#include "common.cuh"
#include "gemm.h"
#include "tc05_ptx.cuh"
#include "chain_geo.h"

namespace {

constexpr int kThreads = 192;
constexpr int CBN = kChainBN;                                     // output tile width (chain_geo.h)
constexpr int kStages = 7;
constexpr uint32_t kBTileBytes = CBN * BK * 2;                    // 8 KB
constexpr uint32_t kStageBytes = kATileBytes + kBTileBytes;       // 24 KB
constexpr int kPitch = CBN + 4;                                   // floats per row of the staged partial tile
constexpr uint32_t kTileBytes = BM * kPitch * 4;                  // 34 KB, its own region: the ring keeps prefetching
constexpr uint32_t kSmemBytes = kStages * kStageBytes + kTileBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kMaxCluster = 8;                                    // portable cluster size limit
constexpr int kMaxLayers = CODAE_CHAIN_MAX_LAYERS;

struct ChainLayer {
    CUtensorMap map_a;            // A(m,k) = input activations [M rows, K cols] bf16, box {64 k, 128 m}
    CUtensorMap map_b;            // weights: K-major [N rows, K cols] box {64 k, 64 n}; MN-major [K rows, N cols] box {64 n, 64 k}
    void* C;
    long long ldc;
    const __nv_bfloat16* mask_src;
    long long ldm;
    int N, K;
    int c_bf16, act, b_kmajor, pad;
};
struct ChainParams {
    ChainLayer layer[kMaxLayers];
    unsigned int* layer_done;     // [num_layers] arrival counters, zeroed by the host before the launch
    int num_layers, M;
};

// work decomposition (chain_geo.h: shared with the host-side unit test tests/test_chain_geometry.py)
__device__ __forceinline__ Geo layer_geo(const ChainParams& P, int l, int S, int rank) {
    return chain_layer_geo(P.layer[l].N, P.layer[l].K, S, rank);
}
// first (layer, tile) at or after (l, t) that exists for this cluster
__device__ __forceinline__ void normalize_item(const ChainParams& P, int cid, int& l, int& t) {
    while (l < P.num_layers && t >= chain_tiles(P.layer[l].N)) { ++l; t = cid; }
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kThreads, 1) tc05_chain_kernel(const __grid_constant__ ChainParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* tile = reinterpret_cast<float*>(smem + kStages * kStageBytes);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kTileBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t S_u, rank_u;
    asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(S_u));
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
    const int S = (int)S_u, rank = (int)rank_u;
    const int C = gridDim.x / S, cid = blockIdx.x / S;
    const unsigned int G = gridDim.x;
    const int L = P.num_layers;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        mbar_init(tmem_empty_bar, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, CBN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                                   // no-op unless launched as a programmatic dependent

    // ring cursors (meaningful in the thread that owns the role)
    int ring_b = 0; uint32_t ring_b_ph = 0;                       // producer: next slot to arm + request the weight tile for
    int ring_a = 0;                                               // producer: next armed slot that still lacks its activation tile
    int ring_m = 0; uint32_t ring_m_ph = 0;                       // MMA issuer
    int pre_done = 0;                                             // producer: k-blocks of the CURRENT item whose weight tile is in flight
    int arrived = 0;                                              // thread 0: layers [0, arrived) signalled by this CTA
    int waited = 0;                                               // thread 0: outputs of layers [0, waited) are known complete
    int acc_it = 0;                                               // items with MMAs so far (all threads; uniform per CTA)

    // one weight tile: arm the slot for both operands, request B
    auto issue_b = [&](int l, int n0, int kb_abs) {
        mbar_wait(&empty_bar[ring_b], ring_b_ph ^ 1);
        mbar_expect_tx(&full_bar[ring_b], kStageBytes);
        uint8_t* b_dst = smem + ring_b * kStageBytes + kATileBytes;
        const int k0 = kb_abs * BK;
        if (P.layer[l].b_kmajor) tma_load_2d(&P.layer[l].map_b, &full_bar[ring_b], b_dst, k0, n0);
        else tma_load_2d(&P.layer[l].map_b, &full_bar[ring_b], b_dst, n0, k0);
        if (++ring_b == kStages) { ring_b = 0; ring_b_ph ^= 1; }
    };
    auto issue_a = [&](int l, int kb_abs) {
        tma_load_2d(&P.layer[l].map_a, &full_bar[ring_a], smem + ring_a * kStageBytes, kb_abs * BK, 0);
        if (++ring_a == kStages) ring_a = 0;
    };
    auto spin_trap = [](long long t0) { if (clock64() - t0 > 4000000000LL) __trap(); };

    int l = 0, t = cid;
    normalize_item(P, cid, l, t);
    if (threadIdx.x == 0 && l < L) {                              // weight tiles of the first item
        const Geo g0 = layer_geo(P, l, S, rank);
        pre_done = min(g0.num_kb, kStages);
        for (int kb = 0; kb < pre_done; ++kb) issue_b(l, t * CBN, g0.kb_begin + kb);
    }

    while (l < L) {
        const Geo geo = layer_geo(P, l, S, rank);
        const ChainLayer& Ly = P.layer[l];
        const int n0 = t * CBN;
        int nl = l, nt = t + C;
        normalize_item(P, cid, nl, nt);

        if (warp == 0) {
            if (lane == 0) {
                // ===== producer (and this CTA's voice on the layer barriers) =====
                while (arrived < l) {                            // layers this CTA is done with (or had no tile of)
                    __threadfence();
                    asm volatile("fence.proxy.async;" ::: "memory");
                    atomicAdd(&P.layer_done[arrived], 1u);
                    ++arrived;
                }
                if (geo.num_kb > 0) {
                    if (l > waited) {                            // the activations of this layer are the previous layer's outputs
                        const long long t0 = clock64();
                        while (ld_acquire_gpu(&P.layer_done[l - 1]) < G) spin_trap(t0);
                        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy stores of other SMs -> our TMA loads
                        waited = l;
                    }
                    for (int kb = 0; kb < geo.num_kb; ++kb) {
                        if (kb >= pre_done) issue_b(l, n0, geo.kb_begin + kb);
                        issue_a(l, geo.kb_begin + kb);
                    }
                }
                pre_done = 0;
                if (nl < L) {                                    // weight tiles of the NEXT item: nothing here depends on them
                    const Geo g2 = layer_geo(P, nl, S, rank);
                    pre_done = min(g2.num_kb, kStages);
                    for (int kb = 0; kb < pre_done; ++kb) issue_b(nl, nt * CBN, g2.kb_begin + kb);
                }
            }
        } else if (warp == 1) {
            if (lane == 0 && geo.num_kb > 0) {
                // ===== MMA issuer =====
                mbar_wait(tmem_empty_bar, (uint32_t)((acc_it & 1) ^ 1));     // epilogue has drained the accumulator (first use: free)
                tc_fence_after();
                const bool bk = Ly.b_kmajor != 0;
                const uint32_t idesc = make_idesc(CBN, true, bk);
                const uint32_t b_adv = bk ? (UMMA_K * 2) : (UMMA_K * 128);
                for (int kb = 0; kb < geo.num_kb; ++kb) {
                    mbar_wait(&full_bar[ring_m], ring_m_ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + ring_m * kStageBytes);
                    const uint32_t b_addr = a_addr + kATileBytes;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(tmem_base, make_desc(a_addr + k * (UMMA_K * 2), true), make_desc(b_addr + k * b_adv, bk), idesc,
                                  (kb | k) != 0);
                    umma_commit(&empty_bar[ring_m]);
                    if (++ring_m == kStages) { ring_m = 0; ring_m_ph ^= 1; }
                }
                umma_commit(tmem_full_bar);
            }
        } else if (geo.num_kb > 0) {
            // ===== epilogue, part 1: TMEM -> fp32 partial tile in shared memory =====
            mbar_wait(tmem_full_bar, (uint32_t)(acc_it & 1));
            tc_fence_after();
            const int q = warp & 3;
            float* stage_row = tile + (size_t)(q * 32 + lane) * kPitch;
#pragma unroll 1
            for (int c = 0; c < CBN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
                float4* dst = reinterpret_cast<float4*>(stage_row + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                         __uint_as_float(v[4 * j + 3]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar);
        }

        cluster_sync();                                           // every partial tile of the cluster is staged

        if (warp >= 2) {
            // ===== epilogue, part 2: CTA r reduces rows [128 r / S, 128 (r+1) / S) of the partials in rank order =====
            const int tt = threadIdx.x - 64;
            const int r_begin = (rank * BM) / S, r_end = ((rank + 1) * BM) / S;
            constexpr int kVecPerRow = CBN / 4;
            const int items = (r_end - r_begin) * kVecPerRow;
            const uint32_t tile_base = smem_u32(tile);
            for (int it = tt; it < items; it += 128) {
                const int rl = r_begin + it / kVecPerRow, c4 = it % kVecPerRow;
                const uint32_t off = (uint32_t)(rl * kPitch + 4 * c4) * 4u;
                float4 part[kMaxCluster];
#pragma unroll
                for (int sp = 0; sp < kMaxCluster; ++sp)
                    if (sp < geo.nsplit) part[sp] = ld_dsmem_f4(tile_base + off, (uint32_t)sp);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int sp = 0; sp < kMaxCluster; ++sp)
                    if (sp < geo.nsplit) { acc.x += part[sp].x; acc.y += part[sp].y; acc.z += part[sp].z; acc.w += part[sp].w; }
                const int row = rl, col = n0 + 4 * c4;             // one row tile: the batch
                if (row >= P.M || col >= Ly.N) continue;
                float f[4] = {acc.x, acc.y, acc.z, acc.w};
                const bool full = col + 4 <= Ly.N;
                if (Ly.act == CODAE_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) f[j] = fmaxf(f[j], 0.f);
                }
                if (Ly.mask_src) {
                    const __nv_bfloat16* mrow = Ly.mask_src + (long long)row * Ly.ldm + col;
                    if (full) {
                        const uint2 mv = *reinterpret_cast<const uint2*>(mrow);
                        if (!(bf16_lo(mv.x) > 0.f)) f[0] = 0.f;
                        if (!(bf16_hi(mv.x) > 0.f)) f[1] = 0.f;
                        if (!(bf16_lo(mv.y) > 0.f)) f[2] = 0.f;
                        if (!(bf16_hi(mv.y) > 0.f)) f[3] = 0.f;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (col + j < Ly.N && !(__bfloat162float(mrow[j]) > 0.f)) f[j] = 0.f;
                    }
                }
                if (Ly.c_bf16) {
                    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(Ly.C) + (long long)row * Ly.ldc + col;
                    if (full) {
                        uint2 o;
                        o.x = pack_bf16x2(f[0], f[1]);
                        o.y = pack_bf16x2(f[2], f[3]);
                        *reinterpret_cast<uint2*>(crow) = o;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (col + j < Ly.N) crow[j] = __float2bfloat16_rn(f[j]);
                    }
                } else {
                    float* crow = reinterpret_cast<float*>(Ly.C) + (long long)row * Ly.ldc + col;
                    if (full) {
                        *reinterpret_cast<float4*>(crow) = make_float4(f[0], f[1], f[2], f[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (col + j < Ly.N) crow[j] = f[j];
                    }
                }
            }
            __threadfence();                                      // this thread's outputs are visible GPU-wide before the barrier below
        }

        cluster_sync();                                           // peers are done reading this CTA's partial tile; outputs are fenced

        if (geo.num_kb > 0) ++acc_it;
        l = nl;
        t = nt;
    }

    if (threadIdx.x == 0) {
        while (arrived < L - 1) {                                 // nobody waits for the last layer
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
            atomicAdd(&P.layer_done[arrived], 1u);
            ++arrived;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, CBN);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int chain_map(codae_ctx* ctx, CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols,
              int box_rows) {
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
        map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return codae_fail(ctx, CODAE_ECUDA, "codae_linear_chain: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return CODAE_OK;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

size_t codae_linear_chain_workspace_bytes(const codae_ctx*) { return kMaxLayers * sizeof(unsigned int); }

int codae_linear_chain(codae_ctx* ctx, const codae_chain_layer* layers, int n_layers, int M, void* workspace, size_t ws_bytes,
                       void* stream) {
    CODAE_REQUIRE(ctx, ctx && layers && workspace, "codae_linear_chain: NULL argument");
    CODAE_REQUIRE(ctx, ctx->encode_tiled, "codae_linear_chain: cuTensorMapEncodeTiled is not available");
    CODAE_REQUIRE(ctx, n_layers >= 1 && n_layers <= kMaxLayers, "codae_linear_chain: 1..%d layers, got %d", kMaxLayers, n_layers);
    CODAE_REQUIRE(ctx, M >= 1 && M <= BM, "codae_linear_chain: the chain kernel holds the batch in one 128-row tile, got M=%d", M);
    if (ws_bytes < kMaxLayers * sizeof(unsigned int))
        return codae_fail(ctx, CODAE_ENOMEM, "codae_linear_chain: workspace %zu < %zu bytes", ws_bytes, kMaxLayers * sizeof(unsigned int));
    ChainParams P;
    memset(&P, 0, sizeof(P));
    int t_max = 1;
    for (int l = 0; l < n_layers; ++l) {
        const codae_chain_layer& s = layers[l];
        CODAE_REQUIRE(ctx, s.A && s.B && s.C && s.N >= 32 && s.K >= 32, "codae_linear_chain: layer %d: bad operand / shape", l);
        CODAE_REQUIRE(ctx, al16(s.A) && al16(s.B) && al16(s.C) && (s.lda % 8) == 0 && (s.ldb % 8) == 0 && s.lda >= s.K,
                      "codae_linear_chain: layer %d: operands must be 16-byte aligned with pitches that are multiples of 8", l);
        CODAE_REQUIRE(ctx, s.c_dtype == CODAE_BF16 ? (s.ldc % 8) == 0 : (s.c_dtype == CODAE_F32 && (s.ldc % 4) == 0),
                      "codae_linear_chain: layer %d: bad output dtype / pitch", l);
        CODAE_REQUIRE(ctx, s.ldc >= s.N && (s.b_kmajor ? s.ldb >= s.K : s.ldb >= s.N), "codae_linear_chain: layer %d: pitch < width", l);
        CODAE_REQUIRE(ctx, !s.mask_src || (al16(s.mask_src) && (s.ldm % 8) == 0 && s.ldm >= s.N), "codae_linear_chain: layer %d: bad mask", l);
        CODAE_REQUIRE(ctx, s.act == CODAE_ACT_NONE || s.act == CODAE_ACT_RELU, "codae_linear_chain: layer %d: bad activation", l);
        ChainLayer& d = P.layer[l];
        int rc = chain_map(ctx, &d.map_a, s.A, M, s.K, s.lda, BK, BM);
        if (rc) return rc;
        if (s.b_kmajor) rc = chain_map(ctx, &d.map_b, s.B, s.N, s.K, s.ldb, BK, CBN);
        else rc = chain_map(ctx, &d.map_b, s.B, s.K, s.N, s.ldb, 64, BK);
        if (rc) return rc;
        d.C = s.C; d.ldc = s.ldc;
        d.mask_src = reinterpret_cast<const __nv_bfloat16*>(s.mask_src); d.ldm = s.ldm;
        d.N = s.N; d.K = s.K;
        d.c_bf16 = s.c_dtype == CODAE_BF16; d.act = s.act; d.b_kmajor = s.b_kmajor ? 1 : 0;
        const int tiles = (s.N + CBN - 1) / CBN;
        if (tiles > t_max) t_max = tiles;
    }
    P.layer_done = reinterpret_cast<unsigned int*>(workspace);
    P.num_layers = n_layers;
    P.M = M;

    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc05_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return codae_fail(ctx, CODAE_ECUDA, "codae_linear_chain: cudaFuncSetAttribute(smem=%u): %s", kSmemBytes, cudaGetErrorString(e));
        attr_set = true;
    }
    // S CTAs per cluster split the k-blocks of a tile; C clusters share the tiles of a layer.  The whole grid must be resident
    // (it spins on the layer counters), and clusters are placed inside one GPC (16 / 18 / 20 SMs on this part), so the number
    // of co-resident clusters depends on S: take the largest S that (a) leaves no rank without k-blocks in every layer and
    // (b) still hosts one cluster per tile of the widest layer; failing (b), the S with the most resident CTAs.
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int S0 = ctx->sm_count / t_max;
    if (S0 > kMaxCluster) S0 = kMaxCluster;
    if (S0 < 1) S0 = 1;
    int S = 0, C = 0, best_S = 0, best_C = 0;
    // the decision only depends on the layer shapes: remember the last one (occupancy queries are host work per call, and
    // the first call of a shape happens outside CUDA-graph capture)
    constexpr int kCache = 4;                            // forward and input-gradient chains of a step alternate
    static int cache_key[kCache][2 * kMaxLayers + 1], cache_S[kCache] = {0}, cache_C[kCache] = {0}, cache_next = 0;
    int key[2 * kMaxLayers + 1] = {0};
    key[0] = n_layers;
    for (int l = 0; l < n_layers; ++l) { key[1 + 2 * l] = layers[l].N; key[2 + 2 * l] = layers[l].K; }
    for (int e = 0; e < kCache; ++e)
        if (cache_S[e] > 0 && memcmp(key, cache_key[e], sizeof(key)) == 0) { S = cache_S[e]; C = cache_C[e]; }
    for (int s_try = S0; s_try >= 1 && S == 0; --s_try) {
        int need = 1;
        for (int l = 0; l < n_layers; ++l) {
            const int ns = chain_layer_geo(layers[l].N, layers[l].K, s_try, 0).nsplit;
            if (ns > need) need = ns;
        }
        if (need < s_try) continue;                      // a smaller cluster gives the same split without idle ranks
        attr[0].val.clusterDim.x = s_try;
        cfg.gridDim = dim3(s_try * t_max);
        int max_clusters = 0;
        cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, tc05_chain_kernel, &cfg);
        if (oe != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        const int c_try = t_max < max_clusters ? t_max : max_clusters;
        if (c_try * s_try > best_C * best_S) { best_S = s_try; best_C = c_try; }
        if (c_try == t_max) break;
    }
    if (S == 0) {
        S = best_S; C = best_C;
        if (S > 0) {
            memcpy(cache_key[cache_next], key, sizeof(key));
            cache_S[cache_next] = S;
            cache_C[cache_next] = C;
            cache_next = (cache_next + 1) % kCache;
        }
    }
    if (S < 1 || C < 1)
        return codae_fail(ctx, CODAE_ECUDA, "codae_linear_chain: no resident cluster with %u bytes of shared memory per CTA", kSmemBytes);
    attr[0].val.clusterDim.x = S;
    cfg.gridDim = dim3(S * C);
    cudaError_t me = cudaMemsetAsync(workspace, 0, kMaxLayers * sizeof(unsigned int), as_stream(stream));
    if (me != cudaSuccess) return codae_fail(ctx, CODAE_ECUDA, "codae_linear_chain: cudaMemsetAsync: %s", cudaGetErrorString(me));
    cudaError_t le = cudaLaunchKernelEx(&cfg, tc05_chain_kernel, P);
    if (le != cudaSuccess) {
        cudaGetLastError();
        return codae_fail(ctx, CODAE_ECUDA, "tc05_chain_kernel launch (%d clusters x %d CTAs, smem %u): %s", C, S, kSmemBytes, cudaGetErrorString(le));
    }
    return codae_check_launch(ctx, "tc05_chain_kernel");
}

}  // extern "C"
