// K3 -- complementarity inference: sweep a catalog shard against query vectors staged in shared
// memory, reduce each candidate's score with warp shuffles and keep the best k per query in-kernel.
// HBM-bound: E*sizeof(dtype) bytes per candidate, nothing written per candidate.
// Ordering is the total order (score, index) -> the selected set and its order do not depend on the
// sweep schedule, so results are bit-reproducible and shard-invariant.
#include <float.h>

#include "common.cuh"

namespace {

constexpr int kWarps = 8;                 // warps per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kMaxK = 128;
constexpr int64_t kEmptyIdx = 0x7fffffffffffffffLL;

// "a is strictly better than b" under the total order; sqerr: smaller score first, cosine: larger first.
template <int METRIC>
__device__ __forceinline__ bool better(float sa, int64_t ia, float sb, int64_t ib) {
    if (METRIC == CODAE_METRIC_SQERR) return sa < sb || (sa == sb && ia < ib);
    return sa > sb || (sa == sb && ia < ib);
}
template <int METRIC>
__device__ __forceinline__ float worst_score() { return METRIC == CODAE_METRIC_SQERR ? INFINITY : -INFINITY; }

// Warp-cooperative sorted insertion into a k-entry list held in shared memory (best first).
// Returns true if the entry made it into the list.  All lanes must call with identical (s, i).
template <int METRIC>
__device__ __forceinline__ bool list_insert(float* ls, int64_t* li, int k, float s, int64_t i, int lane) {
    if (!better<METRIC>(s, i, ls[k - 1], li[k - 1])) return false;
    // position = number of entries strictly better than the new one
    int pos = 0;
    for (int b = 0; b < k; b += 32) {
        const int e = b + lane;
        const bool bt = e < k && better<METRIC>(ls[e], li[e], s, i);
        pos += __popc(__ballot_sync(0xffffffffu, bt));
    }
    // shift [pos, k-1) one slot towards the tail, highest chunk first
    for (int b = ((k - 1) / 32) * 32; b >= 0; b -= 32) {
        const int e = b + lane;
        float ts = 0.f; int64_t ti = 0;
        const bool mv = e > pos && e < k;
        if (mv) { ts = ls[e - 1]; ti = li[e - 1]; }
        __syncwarp();
        if (mv) { ls[e] = ts; li[e] = ti; }
        __syncwarp();
    }
    if (lane == 0) { ls[pos] = s; li[pos] = i; }
    __syncwarp();
    return true;
}

// R candidate rows against QC queries (smem), partial sums per lane.  All loads of a 256-element (bf16) / 128-element (f32)
// column block of the R rows are issued before the first use: R x 16 bytes (bf16) or R x 16 bytes x 1 (f32, four column blocks
// per unrolled pass) in flight per lane -- a 1 KB bf16 row alone leaves a warp with 2 loads in flight and the sweep at half of
// the HBM rate (measured 0.52 of the copy bandwidth; fp32 rows are twice as long and reached 0.94).
// The per-row arithmetic (order of the FMAs inside a lane, then the shuffle tree) does not depend on R: scores are bit-identical
// to the one-row-at-a-time sweep.
//   SQERR : acc[q] = sum (q_d - c_d*inv_scale)^2 ; COSINE: acc[q] = sum c_d q_d, cc = sum c_d^2
template <int METRIC, int QC, bool kBf16, int R>
__device__ __forceinline__ void rows_partial(const unsigned char* const (&row)[R], int E, const float* __restrict__ qs,
                                             float inv_scale, int lane, float (&acc)[R][QC], float (&cc)[R]) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
        cc[j] = 0.f;
#pragma unroll
        for (int q = 0; q < QC; ++q) acc[j][q] = 0.f;
    }
    if (kBf16) {
        for (int c = lane * 8; c < E; c += 256) {
            uint4 p[R];
#pragma unroll
            for (int j = 0; j < R; ++j) p[j] = ldg_stream_u4(reinterpret_cast<const __nv_bfloat16*>(row[j]) + c);
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const float v[8] = {bf16_lo(p[j].x), bf16_hi(p[j].x), bf16_lo(p[j].y), bf16_hi(p[j].y),
                                    bf16_lo(p[j].z), bf16_hi(p[j].z), bf16_lo(p[j].w), bf16_hi(p[j].w)};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (METRIC == CODAE_METRIC_COSINE) cc[j] = fmaf(v[e], v[e], cc[j]);
#pragma unroll
                    for (int q = 0; q < QC; ++q) {
                        const float qv = qs[q * E + c + e];
                        if (METRIC == CODAE_METRIC_SQERR) { const float d = qv - v[e] * inv_scale; acc[j][q] = fmaf(d, d, acc[j][q]); }
                        else acc[j][q] = fmaf(v[e], qv, acc[j][q]);
                    }
                }
            }
        }
    } else {
        for (int c = lane * 4; c < E; c += 128) {
            float4 p[R];
#pragma unroll
            for (int j = 0; j < R; ++j) p[j] = ldg_stream_f4(reinterpret_cast<const float*>(row[j]) + c);
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const float v[4] = {p[j].x, p[j].y, p[j].z, p[j].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (METRIC == CODAE_METRIC_COSINE) cc[j] = fmaf(v[e], v[e], cc[j]);
#pragma unroll
                    for (int q = 0; q < QC; ++q) {
                        const float qv = qs[q * E + c + e];
                        if (METRIC == CODAE_METRIC_SQERR) { const float d = qv - v[e] * inv_scale; acc[j][q] = fmaf(d, d, acc[j][q]); }
                        else acc[j][q] = fmaf(v[e], qv, acc[j][q]);
                    }
                }
            }
        }
    }
}

// One row (rank mode).
template <int METRIC, int QC, bool kBf16>
__device__ __forceinline__ void row_partial(const void* __restrict__ row, int E, const float* __restrict__ qs, float inv_scale,
                                            int lane, float (&acc)[QC], float& cc) {
    const unsigned char* const rows[1] = {reinterpret_cast<const unsigned char*>(row)};
    float a[1][QC], c1[1];
    rows_partial<METRIC, QC, kBf16, 1>(rows, E, qs, inv_scale, lane, a, c1);
#pragma unroll
    for (int q = 0; q < QC; ++q) acc[q] = a[0][q];
    cc = c1[0];
}

template <int METRIC>
__device__ __forceinline__ float finish_score(float acc, float cc, float qq) {
    if (METRIC == CODAE_METRIC_SQERR) return acc;
    return acc / fmaxf(sqrtf(cc * qq), 1e-8f);
}

// Workspace layout: [grid][QC][k] entries per launch (scores then indices).
template <int METRIC, int QC, bool kBf16>
__global__ void __launch_bounds__(kThreads) score_topk_kernel(const void* __restrict__ catalog, int64_t n_rows, int64_t ld,
                                                              int E, int64_t row_offset, const float* __restrict__ query,
                                                              float inv_scale, int k, float* __restrict__ ws_score,
                                                              int64_t* __restrict__ ws_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* qs = reinterpret_cast<float*>(smem_raw);                        // [QC][E]
    float* qq = qs + QC * E;                                               // [QC] |q|^2
    int64_t* li = reinterpret_cast<int64_t*>(qq + ((QC + 3) & ~3));         // [kWarps][QC][k]
    float* ls = reinterpret_cast<float*>(li + kWarps * QC * k);            // [kWarps][QC][k]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < QC * E; e += kThreads) qs[e] = query[e];
    for (int e = threadIdx.x; e < kWarps * QC * k; e += kThreads) { ls[e] = worst_score<METRIC>(); li[e] = kEmptyIdx; }
    __syncthreads();
    if (warp < QC) {
        float s = 0.f;
        for (int c = lane; c < E; c += 32) s = fmaf(qs[warp * E + c], qs[warp * E + c], s);
        s = warp_sum(s);
        if (lane == 0) qq[warp] = s;
    }
    __syncthreads();
    float* my_s = ls + warp * QC * k;
    int64_t* my_i = li + warp * QC * k;
    const size_t esz = kBf16 ? 2 : 4;
    const int64_t wstride = (int64_t)gridDim.x * kWarps;
    // rows per warp pass = independent 16-byte loads per lane and column block (fewer with four queries: register budget)
    constexpr int R = (kBf16 ? 8 : 4) / (QC == 1 ? 1 : 2);
    for (int64_t r = (int64_t)blockIdx.x * kWarps + warp; r < n_rows; r += R * wstride) {
        const unsigned char* row[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int64_t rj = r + j * wstride;
            row[j] = reinterpret_cast<const unsigned char*>(catalog) + (size_t)(rj < n_rows ? rj : r) * ld * esz;   // tail: re-read row r, result unused
        }
        float acc[R][QC], cc[R];
        rows_partial<METRIC, QC, kBf16, R>(row, E, qs, inv_scale, lane, acc, cc);
        // R x QC independent shuffle trees: the compiler interleaves them
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if (METRIC == CODAE_METRIC_COSINE) cc[j] = warp_sum(cc[j]);
#pragma unroll
            for (int q = 0; q < QC; ++q) acc[j][q] = warp_sum(acc[j][q]);
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int64_t rj = r + j * wstride;
            if (rj >= n_rows) break;
#pragma unroll
            for (int q = 0; q < QC; ++q) {
                const float s = finish_score<METRIC>(acc[j][q], cc[j], qq[q]);
                if (s == s) list_insert<METRIC>(my_s + q * k, my_i + q * k, k, s, row_offset + rj, lane);  // NaN never ranks
            }
        }
    }
    __syncthreads();
    // CTA merge: warp w (< QC) folds the lists of query w from all warps; sorted inputs -> stop at first reject
    for (int q = warp; q < QC; q += kWarps) {
        float* dst_s = ls + q * k;            // warp 0's list of query q
        int64_t* dst_i = li + q * k;
        for (int w = 1; w < kWarps; ++w) {
            const float* src_s = ls + (w * QC + q) * k;
            const int64_t* src_i = li + (w * QC + q) * k;
            for (int e = 0; e < k; ++e) {
                if (src_i[e] == kEmptyIdx) break;
                if (!list_insert<METRIC>(dst_s, dst_i, k, src_s[e], src_i[e], lane)) break;
            }
        }
        for (int e = lane; e < k; e += 32) {
            ws_score[((int64_t)blockIdx.x * QC + q) * k + e] = dst_s[e];
            ws_idx[((int64_t)blockIdx.x * QC + q) * k + e] = dst_i[e];
        }
    }
}

// Merge `n_lists` sorted k-lists per query (layout [n_lists][Q][k]) into out [Q][k].  One CTA per query;
// every warp folds a strided subset of the lists, then warp 0 folds the per-warp results.
// A warp first copies its lists into shared memory with all lanes loading (hundreds of independent loads in flight), then
// folds them from there: folding straight from global memory is a chain of dependent ~0.6 us loads per list (measured 48.9 us
// for the 592 lists of one sweep).
constexpr int kMergeStage = 1024;          // entries per warp staged per pass (12 KB)
inline size_t merge_smem() { return (size_t)kWarps * kMergeStage * 12; }
template <int METRIC>
__global__ void __launch_bounds__(kThreads) topk_merge_kernel(const float* __restrict__ in_s, const int64_t* __restrict__ in_i,
                                                              int n_lists, int Q, int k, float* __restrict__ out_s,
                                                              int64_t* __restrict__ out_i, int q_out_offset, int q_out_stride) {
    __shared__ float ls[kWarps][kMaxK];
    __shared__ int64_t li[kWarps][kMaxK];
    extern __shared__ __align__(16) unsigned char merge_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = blockIdx.x;
    int64_t* st_i = reinterpret_cast<int64_t*>(merge_raw) + warp * kMergeStage;
    float* st_s = reinterpret_cast<float*>(merge_raw + (size_t)kWarps * kMergeStage * 8) + warp * kMergeStage;
    for (int e = lane; e < k; e += 32) { ls[warp][e] = worst_score<METRIC>(); li[warp][e] = kEmptyIdx; }
    __syncwarp();
    const int per_pass = kMergeStage / k;                              // lists per pass (k <= 128 -> at least 8)
    const int mine = n_lists > warp ? (n_lists - warp + kWarps - 1) / kWarps : 0;    // lists warp, warp + kWarps, ...
    for (int j0 = 0; j0 < mine; j0 += per_pass) {
        const int nl = min(per_pass, mine - j0);
        for (int t = lane; t < nl * k; t += 32) {
            const int j = t / k, e = t - j * k;
            const int64_t src = ((int64_t)(warp + (j0 + j) * kWarps) * Q + q) * k + e;
            st_s[t] = in_s[src];
            st_i[t] = in_i[src];
        }
        __syncwarp();
        for (int j = 0; j < nl; ++j)
            for (int e = 0; e < k; ++e) {
                const int64_t ii = st_i[j * k + e];
                if (ii == kEmptyIdx || ii < 0) break;
                if (!list_insert<METRIC>(ls[warp], li[warp], k, st_s[j * k + e], ii, lane)) break;
            }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < kWarps; ++w)
            for (int e = 0; e < k; ++e) {
                if (li[w][e] == kEmptyIdx) break;
                if (!list_insert<METRIC>(ls[0], li[0], k, ls[w][e], li[w][e], lane)) break;
            }
        const int qo = q_out_offset + q * q_out_stride;
        for (int e = lane; e < k; e += 32) {
            const bool empty = li[0][e] == kEmptyIdx;
            out_s[(int64_t)qo * k + e] = ls[0][e];
            out_i[(int64_t)qo * k + e] = empty ? -1 : li[0][e];
        }
    }
}

inline void merge_attrs() {      // the staging buffers need the opt-in shared-memory size (once per process)
    static const bool done = (cudaFuncSetAttribute(topk_merge_kernel<CODAE_METRIC_SQERR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem()),
                              cudaFuncSetAttribute(topk_merge_kernel<CODAE_METRIC_COSINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem()), true);
    (void)done;
}

// Rank mode: s_true[q] = score(query q, row true_idx[q]); out_rank[q] = #{j in subset : s_true[q] better-than s_q[j]}.
template <int METRIC, bool kBf16>
__global__ void __launch_bounds__(kThreads) score_rank_kernel(const void* __restrict__ catalog, int64_t n_rows, int64_t ld,
                                                              int E, const float* __restrict__ query, int Q, float inv_scale,
                                                              const int64_t* __restrict__ true_idx,
                                                              const int64_t* __restrict__ subset, int64_t n_subset,
                                                              unsigned long long* __restrict__ out_rank) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_true = reinterpret_cast<float*>(smem_raw);     // [Q]
    float* qq = s_true + Q;                                 // [Q]
    int* cnt = reinterpret_cast<int*>(qq + Q);              // [Q]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t esz = kBf16 ? 2 : 4;
    for (int q = threadIdx.x; q < Q; q += kThreads) cnt[q] = 0;
    for (int q = warp; q < Q; q += kWarps) {
        float s = 0.f;
        for (int c = lane; c < E; c += 32) { const float v = query[(int64_t)q * E + c]; s = fmaf(v, v, s); }
        s = warp_sum(s);
        float acc[1], cc;
        row_partial<METRIC, 1, kBf16>(reinterpret_cast<const unsigned char*>(catalog) + (size_t)true_idx[q] * ld * esz, E,
                                      query + (int64_t)q * E, inv_scale, lane, acc, cc);
        if (METRIC == CODAE_METRIC_COSINE) cc = warp_sum(cc);
        const float st = finish_score<METRIC>(warp_sum(acc[0]), cc, s);
        if (lane == 0) { qq[q] = s; s_true[q] = st; }
    }
    __syncthreads();
    const int64_t wstride = (int64_t)gridDim.x * kWarps;
    for (int64_t t = (int64_t)blockIdx.x * kWarps + warp; t < n_subset; t += wstride) {
        const int64_t r = subset ? subset[t] : t;
        const unsigned char* row = reinterpret_cast<const unsigned char*>(catalog) + (size_t)r * ld * esz;
        for (int q = 0; q < Q; ++q) {
            float acc[1], cc;
            row_partial<METRIC, 1, kBf16>(row, E, query + (int64_t)q * E, inv_scale, lane, acc, cc);
            if (METRIC == CODAE_METRIC_COSINE) cc = warp_sum(cc);
            const float s = finish_score<METRIC>(warp_sum(acc[0]), cc, qq[q]);
            const bool b = METRIC == CODAE_METRIC_SQERR ? (s_true[q] < s) : (s_true[q] > s);
            if (lane == 0 && b) atomicAdd(&cnt[q], 1);
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < Q; q += kThreads)
        if (cnt[q]) atomicAdd(&out_rank[q], (unsigned long long)cnt[q]);
}

// ---- rank mode for large query batches: the Q x n dot products come from the fp32-parity tensor-core GEMM ---------------------
// RankingLoss.get runs for every validation batch (train_dae_on_embedding.py:241-261; metering.py:46-79).  For Q >= 64 queries of
// one category the dot products q . c_j are one [Q, E] x [E, n] contraction (codae_linear_fwd_x3: fp32 fidelity on tensor cores);
// these two kernels supply the norms and turn the score matrix into ranks.
//   row_sqnorm_kernel  out[r] = sum_d X[r, d]^2                  (warp per row, fixed shuffle tree)
//   rank_count_kernel  out_rank[q] = #{ j < n : cos(q, true_q) > cos(q, c_j) } with cos = dot / max(sqrt(|c|^2 |q|^2), 1e-8),
//                      dot(q, true_q) = scores[q, n + q] (the true rows are appended to the catalog operand so that a true item
//                      that is itself in the subset gets bit-identical scores on both sides of the strict comparison)
__global__ void __launch_bounds__(kThreads) row_sqnorm_kernel(const float* __restrict__ X, int64_t ld, int64_t rows, int E,
                                                              float* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t r = (int64_t)blockIdx.x * kWarps + warp; r < rows; r += (int64_t)gridDim.x * kWarps) {
        const float* row = X + r * ld;
        float s = 0.f;
        for (int c = lane; c < E; c += 32) { const float v = row[c]; s = fmaf(v, v, s); }
        s = warp_sum(s);
        if (lane == 0) out[r] = s;
    }
}

__global__ void __launch_bounds__(kThreads) rank_count_kernel(const float* __restrict__ scores, int64_t ld, int Q, int64_t n,
                                                              const float* __restrict__ cc, const float* __restrict__ qq,
                                                              unsigned long long* __restrict__ out_rank) {
    __shared__ unsigned int part[kWarps];
    const int q = blockIdx.y;
    const float* row = scores + (int64_t)q * ld;
    const float nq = qq[q];
    const float s_true = row[n + q] / fmaxf(sqrtf(cc[n + q] * nq), 1e-8f);
    unsigned int cnt = 0;
    for (int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += (int64_t)gridDim.x * kThreads) {
        const float s = row[j] / fmaxf(sqrtf(cc[j] * nq), 1e-8f);
        cnt += (s_true > s) ? 1u : 0u;
    }
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
        for (int w = 0; w < kWarps; ++w) t += part[w];
        if (t) atomicAdd(&out_rank[q], (unsigned long long)t);      // integer counts: order-independent
    }
}

// ---- candidate SWAPS scored by full reconstruction error --------------------------------------------------------
// Second reading of "scores candidate item swaps by reconstruction error": put candidate j into slot c of the outfit, run
// the DAE on the swapped outfit (the tensor-core GEMMs, batched over candidates) and score the swap by
// || DAE(outfit_j) - outfit_j ||^2 over all io dimensions.  Two small kernels bracket the forward pass:
//   swap_build_kernel      x'[b, :] = outfit with columns [slot*E, (slot+1)*E) replaced by catalog[first_row + b] * inv_scale
//   swap_error_topk_kernel err_b = sum_d (y[b, d] - x'[b, d])^2 (x' recomputed in fp32), best k per CTA, then topk_merge_kernel
template <bool kBf16Cat, bool kBf16Out>
__global__ void __launch_bounds__(256) swap_build_kernel(const float* __restrict__ outfit, const void* __restrict__ catalog,
                                                         int64_t ld_cat, int64_t first_row, int B, int E, int slot, int io,
                                                         float inv_scale, void* __restrict__ out, int64_t ld_out) {
    const int64_t total = (int64_t)B * io;
    const int lo = slot * E, hi = lo + E;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(e / io), d = (int)(e - (int64_t)b * io);
        float v;
        if (d >= lo && d < hi) {
            const int64_t off = (first_row + b) * ld_cat + (d - lo);
            v = (kBf16Cat ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(catalog)[off])
                          : reinterpret_cast<const float*>(catalog)[off]) * inv_scale;
        } else {
            v = outfit[d];
        }
        if (kBf16Out) reinterpret_cast<__nv_bfloat16*>(out)[(int64_t)b * ld_out + d] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(out)[(int64_t)b * ld_out + d] = v;
    }
}

template <bool kBf16Cat>
__global__ void __launch_bounds__(kThreads) swap_error_topk_kernel(const float* __restrict__ outfit, const void* __restrict__ catalog,
                                                                   int64_t ld_cat, int64_t first_row, int B, int E, int slot, int io,
                                                                   float inv_scale, const float* __restrict__ y, int64_t ld_y,
                                                                   int64_t row_offset, int k, float* __restrict__ ws_score,
                                                                   int64_t* __restrict__ ws_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int64_t* li = reinterpret_cast<int64_t*>(smem_raw);                    // [kWarps][k]
    float* ls = reinterpret_cast<float*>(li + kWarps * k);                 // [kWarps][k]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < kWarps * k; e += kThreads) { ls[e] = INFINITY; li[e] = kEmptyIdx; }
    __syncthreads();
    const int lo = slot * E, hi = lo + E;
    for (int b = blockIdx.x * kWarps + warp; b < B; b += gridDim.x * kWarps) {
        float acc = 0.f;
        for (int d = lane; d < io; d += 32) {
            float t;
            if (d >= lo && d < hi) {
                const int64_t off = (first_row + b) * ld_cat + (d - lo);
                t = (kBf16Cat ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(catalog)[off])
                              : reinterpret_cast<const float*>(catalog)[off]) * inv_scale;
            } else {
                t = outfit[d];
            }
            const float df = y[(int64_t)b * ld_y + d] - t;
            acc = fmaf(df, df, acc);
        }
        acc = warp_sum(acc);
        if (acc == acc) list_insert<CODAE_METRIC_SQERR>(ls + warp * k, li + warp * k, k, acc, row_offset + first_row + b, lane);
    }
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < kWarps; ++w)
            for (int e = 0; e < k; ++e) {
                if (li[w * k + e] == kEmptyIdx) break;
                if (!list_insert<CODAE_METRIC_SQERR>(ls, li, k, ls[w * k + e], li[w * k + e], lane)) break;
            }
        for (int e = lane; e < k; e += 32) {
            ws_score[(int64_t)blockIdx.x * k + e] = ls[e];
            ws_idx[(int64_t)blockIdx.x * k + e] = li[e];
        }
    }
}

inline int sweep_grid(const codae_ctx* ctx, int64_t n_rows) {
    int64_t g = (n_rows + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)ctx->sm_count * 4;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

inline size_t topk_smem(int QC, int E, int k) {
    return (size_t)QC * E * 4 + (size_t)((QC + 3) & ~3) * 4 + (size_t)kWarps * QC * k * 12;
}

}  // namespace

extern "C" {

size_t codae_score_topk_workspace_bytes(const codae_ctx* ctx, int Q, int k) {
    if (!ctx || Q < 1 || k < 1) return 0;
    return (size_t)ctx->sm_count * 4 * 4 /*QC max*/ * (size_t)k * 12 + 256;
}

int codae_score_topk(codae_ctx* ctx, const void* catalog, int cat_dtype, int64_t n_rows, int64_t ld, int E,
                     int64_t row_offset, const float* query, int Q, float inv_scale, int metric, int k, float* out_score,
                     int64_t* out_idx, void* workspace, size_t ws_bytes, void* stream) {
    CODAE_REQUIRE(ctx, ctx && catalog && query && out_score && out_idx && workspace, "codae_score_topk: NULL argument");
    CODAE_REQUIRE(ctx, k >= 1 && k <= kMaxK, "codae_score_topk: k %d outside [1, %d]", k, kMaxK);
    CODAE_REQUIRE(ctx, Q >= 1 && n_rows >= 0, "codae_score_topk: bad Q / n_rows");
    CODAE_REQUIRE(ctx, metric == CODAE_METRIC_SQERR || metric == CODAE_METRIC_COSINE, "codae_score_topk: bad metric %d", metric);
    const bool bf = cat_dtype == CODAE_BF16;
    CODAE_REQUIRE(ctx, cat_dtype == CODAE_F32 || bf, "codae_score_topk: bad cat_dtype %d", cat_dtype);
    const int vec = bf ? 8 : 4;
    CODAE_REQUIRE(ctx, E >= vec && E <= 4096 && E % vec == 0 && ld >= E && ld % vec == 0 &&
                           (reinterpret_cast<uintptr_t>(catalog) & 15) == 0,
                  "codae_score_topk: E=%d ld=%lld must be multiples of %d (16-byte rows), E <= 4096", E, (long long)ld, vec);
    if (ws_bytes < codae_score_topk_workspace_bytes(ctx, Q, k))
        return codae_fail(ctx, CODAE_ENOMEM, "codae_score_topk: workspace %zu < %zu bytes", ws_bytes,
                          codae_score_topk_workspace_bytes(ctx, Q, k));
    cudaStream_t s = as_stream(stream);
    merge_attrs();
    const int grid0 = sweep_grid(ctx, n_rows);
    int64_t* ws_idx = reinterpret_cast<int64_t*>(workspace);
    float* ws_score = reinterpret_cast<float*>(ws_idx + (size_t)grid0 * 4 * k);
    for (int q0 = 0; q0 < Q; q0 += 4) {
        int grid = grid0;
        const int qc = (Q - q0 >= 4) ? 4 : 1;
        const int reps = (Q - q0 >= 4) ? 1 : (Q - q0);  // leftover queries one at a time
        for (int rpt = 0; rpt < reps; ++rpt) {
            const float* qptr = query + (int64_t)(q0 + rpt) * E;
            const size_t smem = topk_smem(qc, E, k);
#define LAUNCH(MET, QC, BF)                                                                                       \
    do {                                                                                                          \
        cudaFuncSetAttribute(score_topk_kernel<MET, QC, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        /* one resident wave: the grid-stride sweep balances itself only if every CTA runs from the start */          \
        const int occ = resident_ctas_per_sm(score_topk_kernel<MET, QC, BF>, kThreads, smem);                          \
        if (grid > ctx->sm_count * occ) grid = ctx->sm_count * occ;                                                  \
        score_topk_kernel<MET, QC, BF><<<grid, kThreads, smem, s>>>(catalog, n_rows, ld, E, row_offset, qptr, inv_scale, k, \
                                                                    ws_score, ws_idx);                           \
    } while (0)
            if (metric == CODAE_METRIC_SQERR) {
                if (qc == 4) { if (bf) LAUNCH(CODAE_METRIC_SQERR, 4, true); else LAUNCH(CODAE_METRIC_SQERR, 4, false); }
                else { if (bf) LAUNCH(CODAE_METRIC_SQERR, 1, true); else LAUNCH(CODAE_METRIC_SQERR, 1, false); }
            } else {
                if (qc == 4) { if (bf) LAUNCH(CODAE_METRIC_COSINE, 4, true); else LAUNCH(CODAE_METRIC_COSINE, 4, false); }
                else { if (bf) LAUNCH(CODAE_METRIC_COSINE, 1, true); else LAUNCH(CODAE_METRIC_COSINE, 1, false); }
            }
#undef LAUNCH
            int rc = codae_check_launch(ctx, "score_topk_kernel");
            if (rc) return rc;
            if (metric == CODAE_METRIC_SQERR)
                topk_merge_kernel<CODAE_METRIC_SQERR><<<qc, kThreads, merge_smem(), s>>>(ws_score, ws_idx, grid, qc, k, out_score, out_idx, q0 + rpt, 1);
            else
                topk_merge_kernel<CODAE_METRIC_COSINE><<<qc, kThreads, merge_smem(), s>>>(ws_score, ws_idx, grid, qc, k, out_score, out_idx, q0 + rpt, 1);
            rc = codae_check_launch(ctx, "topk_merge_kernel");
            if (rc) return rc;
        }
    }
    return CODAE_OK;
}

int codae_topk_merge(codae_ctx* ctx, const float* scores, const int64_t* idx, int G, int Q, int k, int metric,
                     float* out_score, int64_t* out_idx, void* stream) {
    CODAE_REQUIRE(ctx, ctx && scores && idx && out_score && out_idx, "codae_topk_merge: NULL argument");
    CODAE_REQUIRE(ctx, G >= 1 && Q >= 1 && k >= 1 && k <= kMaxK, "codae_topk_merge: bad shape");
    merge_attrs();
    if (metric == CODAE_METRIC_SQERR)
        topk_merge_kernel<CODAE_METRIC_SQERR><<<Q, kThreads, merge_smem(), as_stream(stream)>>>(scores, idx, G, Q, k, out_score, out_idx, 0, 1);
    else if (metric == CODAE_METRIC_COSINE)
        topk_merge_kernel<CODAE_METRIC_COSINE><<<Q, kThreads, merge_smem(), as_stream(stream)>>>(scores, idx, G, Q, k, out_score, out_idx, 0, 1);
    else
        return codae_fail(ctx, CODAE_EINVAL, "codae_topk_merge: bad metric %d", metric);
    return codae_check_launch(ctx, "topk_merge_kernel");
}

int codae_swap_build(codae_ctx* ctx, const float* outfit, const void* catalog, int cat_dtype, int64_t ld_cat, int64_t first_row,
                     int B, int E, int slot, int io, float inv_scale, void* out_x, int x_dtype, int64_t ld_x, void* stream) {
    CODAE_REQUIRE(ctx, ctx && outfit && catalog && out_x, "codae_swap_build: NULL argument");
    CODAE_REQUIRE(ctx, B >= 1 && E >= 1 && slot >= 0 && (slot + 1) * E <= io && ld_x >= io && ld_cat >= E && first_row >= 0,
                  "codae_swap_build: bad shape (B=%d E=%d slot=%d io=%d)", B, E, slot, io);
    const bool bc = cat_dtype == CODAE_BF16, bo = x_dtype == CODAE_BF16;
    int64_t blocks = ((int64_t)B * io + 256 * 4 - 1) / (256 * 4);
    if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
    cudaStream_t s = as_stream(stream);
#define LAUNCH(BC, BO) swap_build_kernel<BC, BO><<<(unsigned)blocks, 256, 0, s>>>(outfit, catalog, ld_cat, first_row, B, E, slot, io, inv_scale, out_x, ld_x)
    if (bc && bo) LAUNCH(true, true); else if (bc) LAUNCH(true, false); else if (bo) LAUNCH(false, true); else LAUNCH(false, false);
#undef LAUNCH
    return codae_check_launch(ctx, "swap_build_kernel");
}

int codae_swap_error_topk(codae_ctx* ctx, const float* outfit, const void* catalog, int cat_dtype, int64_t ld_cat,
                          int64_t first_row, int B, int E, int slot, int io, float inv_scale, const float* y, int64_t ld_y,
                          int64_t row_offset, int k, float* out_score, int64_t* out_idx, void* workspace, size_t ws_bytes,
                          void* stream) {
    CODAE_REQUIRE(ctx, ctx && outfit && catalog && y && out_score && out_idx && workspace, "codae_swap_error_topk: NULL argument");
    CODAE_REQUIRE(ctx, B >= 1 && E >= 1 && slot >= 0 && (slot + 1) * E <= io && ld_y >= io && ld_cat >= E && k >= 1 && k <= kMaxK,
                  "codae_swap_error_topk: bad shape");
    if (ws_bytes < codae_score_topk_workspace_bytes(ctx, 1, k))
        return codae_fail(ctx, CODAE_ENOMEM, "codae_swap_error_topk: workspace %zu < %zu bytes", ws_bytes,
                          codae_score_topk_workspace_bytes(ctx, 1, k));
    cudaStream_t s = as_stream(stream);
    const int grid = sweep_grid(ctx, B);
    int64_t* ws_idx = reinterpret_cast<int64_t*>(workspace);
    float* ws_score = reinterpret_cast<float*>(ws_idx + (size_t)grid * k);
    const size_t smem = (size_t)kWarps * k * 12;
    if (cat_dtype == CODAE_BF16)
        swap_error_topk_kernel<true><<<grid, kThreads, smem, s>>>(outfit, catalog, ld_cat, first_row, B, E, slot, io, inv_scale, y, ld_y, row_offset, k, ws_score, ws_idx);
    else
        swap_error_topk_kernel<false><<<grid, kThreads, smem, s>>>(outfit, catalog, ld_cat, first_row, B, E, slot, io, inv_scale, y, ld_y, row_offset, k, ws_score, ws_idx);
    int rc = codae_check_launch(ctx, "swap_error_topk_kernel");
    if (rc) return rc;
    merge_attrs();
    topk_merge_kernel<CODAE_METRIC_SQERR><<<1, kThreads, merge_smem(), s>>>(ws_score, ws_idx, grid, 1, k, out_score, out_idx, 0, 1);
    return codae_check_launch(ctx, "topk_merge_kernel");
}

int codae_score_rank(codae_ctx* ctx, const void* catalog, int cat_dtype, int64_t n_rows, int64_t ld, int E,
                     const float* query, int Q, float inv_scale, int metric, const int64_t* true_idx,
                     const int64_t* subset_idx, int64_t n_subset, int64_t* out_rank, void* stream) {
    CODAE_REQUIRE(ctx, ctx && catalog && query && true_idx && out_rank, "codae_score_rank: NULL argument");
    CODAE_REQUIRE(ctx, Q >= 1 && Q <= 1024, "codae_score_rank: Q %d outside [1, 1024]", Q);
    const bool bf = cat_dtype == CODAE_BF16;
    const int vec = bf ? 8 : 4;
    CODAE_REQUIRE(ctx, E >= vec && E <= 4096 && E % vec == 0 && ld >= E && ld % vec == 0 &&
                           (reinterpret_cast<uintptr_t>(catalog) & 15) == 0 && (reinterpret_cast<uintptr_t>(query) & 15) == 0,
                  "codae_score_rank: E=%d ld=%lld must be multiples of %d", E, (long long)ld, vec);
    cudaStream_t s = as_stream(stream);
    if (!subset_idx) n_subset = n_rows;
    cudaError_t e = cudaMemsetAsync(out_rank, 0, sizeof(int64_t) * Q, s);
    if (e != cudaSuccess) return codae_fail(ctx, CODAE_ECUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const int grid = sweep_grid(ctx, n_subset);
    const size_t smem = (size_t)Q * 12;
    unsigned long long* out = reinterpret_cast<unsigned long long*>(out_rank);
#define LAUNCH(MET, BF) score_rank_kernel<MET, BF><<<grid, kThreads, smem, s>>>(catalog, n_rows, ld, E, query, Q, inv_scale, true_idx, subset_idx, n_subset, out)
    if (metric == CODAE_METRIC_SQERR) { if (bf) LAUNCH(CODAE_METRIC_SQERR, true); else LAUNCH(CODAE_METRIC_SQERR, false); }
    else if (metric == CODAE_METRIC_COSINE) { if (bf) LAUNCH(CODAE_METRIC_COSINE, true); else LAUNCH(CODAE_METRIC_COSINE, false); }
    else return codae_fail(ctx, CODAE_EINVAL, "codae_score_rank: bad metric %d", metric);
#undef LAUNCH
    return codae_check_launch(ctx, "score_rank_kernel");
}

int codae_row_sqnorm(codae_ctx* ctx, const float* X, int64_t ld, int64_t rows, int E, float* out, void* stream) {
    CODAE_REQUIRE(ctx, ctx && X && out && rows >= 0 && E >= 1 && ld >= E, "codae_row_sqnorm: bad argument");
    if (rows == 0) return CODAE_OK;
    row_sqnorm_kernel<<<sweep_grid(ctx, rows), kThreads, 0, as_stream(stream)>>>(X, ld, rows, E, out);
    return codae_check_launch(ctx, "row_sqnorm_kernel");
}

int codae_rank_count(codae_ctx* ctx, const float* scores, int64_t ld, int Q, int64_t n, const float* cc, const float* qq,
                     int64_t* out_rank, void* stream) {
    CODAE_REQUIRE(ctx, ctx && scores && cc && qq && out_rank, "codae_rank_count: NULL argument");
    CODAE_REQUIRE(ctx, Q >= 1 && Q <= 65535 && n >= 0 && ld >= n + Q, "codae_rank_count: bad shape Q=%d n=%lld ld=%lld", Q, (long long)n, (long long)ld);
    cudaStream_t s = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(out_rank, 0, sizeof(int64_t) * Q, s);
    if (e != cudaSuccess) return codae_fail(ctx, CODAE_ECUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    int gx = (int)((n + kThreads * 8 - 1) / (kThreads * 8));
    if (gx < 1) gx = 1;
    if (gx > ctx->sm_count) gx = ctx->sm_count;
    rank_count_kernel<<<dim3(gx, Q), kThreads, 0, s>>>(scores, ld, Q, n, cc, qq, reinterpret_cast<unsigned long long*>(out_rank));
    return codae_check_launch(ctx, "rank_count_kernel");
}

}  // extern "C"
