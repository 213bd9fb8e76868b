// K1-loss -- reconstruction loss forward + backward + monitor sums in one HBM pass, and the
// tabular (abalone) CombinedCriterion loss / monitor kernels.
#include "common.cuh"

namespace {

constexpr int kLossThreads = 256;
constexpr int kLossUnroll = 4;
constexpr int kLossMaxGrid = 148 * 8 * 2;  // workspace is sized for this many CTA partials

struct LossWs {
    unsigned int ticket;
    unsigned int pad[3];
    double partial[kLossMaxGrid][2];
};

__device__ __forceinline__ float keep_of(uint64_t bits, uint32_t var) { return ((bits >> var) & 1ull) ? 0.0f : 1.0f; }

// kDy: CODAE_F32 / CODAE_BF16 / CODAE_F32X3 (dL/dy as three bf16 planes B * ld_dy elements apart: fp32-parity engine, kVec only)
template <bool kBf16Y, int kDy, bool kVec>
__global__ void __launch_bounds__(kLossThreads) mse_loss_kernel(
    const float* __restrict__ x, int64_t ld_x, const int64_t* __restrict__ batch_idx, const void* __restrict__ y,
    int64_t ld_y, const int32_t* __restrict__ mask_id, const uint64_t* __restrict__ mask_bits,
    const uint8_t* __restrict__ col_var, int B, int io, int log2g, float grad_scale, void* __restrict__ dy, int64_t ld_dy,
    double* __restrict__ acc, LossWs* __restrict__ ws) {
    pdl_launch_dependents();
    pdl_wait();
    float s_full = 0.f, s_part = 0.f;
    if (kVec) {
        // a group of G = 2^log2g threads owns one row at a time (same schedule as corrupt_fwd_vec_kernel): per-row lookups
        // once, the next row's lookups under the loads in flight, four independent 128-bit loads of x and y per thread
        const int io4 = io >> 2;
        const int G = 1 << log2g, rows_per_cta = kLossThreads >> log2g;
        const int gl = threadIdx.x & (G - 1);
        const int row_stride = gridDim.x * rows_per_cta;
        int row = blockIdx.x * rows_per_cta + (threadIdx.x >> log2g);
        int64_t obs = 0;
        uint64_t bits = 0;
        if (row < B) {
            obs = batch_idx ? batch_idx[row] : (int64_t)row;
            bits = mask_bits[mask_id[row]];
        }
        for (; row < B; row += row_stride) {
            const float* xrow = x + obs * ld_x;
            int64_t n_obs = 0;
            uint64_t n_bits = 0;
            bool first_chunk = true;
            for (int c0 = 0; c0 < io4; c0 += kLossUnroll * G) {
                float4 xv[kLossUnroll], yv[kLossUnroll];
#pragma unroll
                for (int j = 0; j < kLossUnroll; ++j) {
                    const int c = 4 * (c0 + gl + j * G);
                    if (c < io) {
                        xv[j] = ldg_stream_f4(xrow + c);
                        if (kBf16Y) {
                            const uint2 p = ldg_stream_u2(reinterpret_cast<const __nv_bfloat16*>(y) + (int64_t)row * ld_y + c);
                            yv[j] = make_float4(bf16_lo(p.x), bf16_hi(p.x), bf16_lo(p.y), bf16_hi(p.y));
                        } else {
                            yv[j] = ldg_stream_f4(reinterpret_cast<const float*>(y) + (int64_t)row * ld_y + c);
                        }
                    }
                }
                if (first_chunk) {
                    first_chunk = false;
                    const int nrow = row + row_stride;
                    if (nrow < B) {
                        n_obs = batch_idx ? batch_idx[nrow] : (int64_t)nrow;
                        n_bits = mask_bits[mask_id[nrow]];
                    }
                }
#pragma unroll
                for (int j = 0; j < kLossUnroll; ++j) {
                    const int c = 4 * (c0 + gl + j * G);
                    if (c >= io) continue;
                    const uchar4 var = *reinterpret_cast<const uchar4*>(col_var + c);
                    const float d0 = yv[j].x - xv[j].x, d1 = yv[j].y - xv[j].y, d2 = yv[j].z - xv[j].z, d3 = yv[j].w - xv[j].w;
                    const float q0 = d0 * d0, q1 = d1 * d1, q2 = d2 * d2, q3 = d3 * d3;
                    s_full += (q0 + q1) + (q2 + q3);
                    s_part += ((1.f - keep_of(bits, var.x)) * q0 + (1.f - keep_of(bits, var.y)) * q1) +
                              ((1.f - keep_of(bits, var.z)) * q2 + (1.f - keep_of(bits, var.w)) * q3);
                    if (dy) {
                        if (kDy == CODAE_F32X3) {
                            store_planes4(reinterpret_cast<__nv_bfloat16*>(dy) + (int64_t)row * ld_dy + c, (int64_t)B * ld_dy,
                                          make_float4(grad_scale * d0, grad_scale * d1, grad_scale * d2, grad_scale * d3));
                        } else if (kDy == CODAE_BF16) {
                            uint2 p;
                            p.x = pack_bf16x2(grad_scale * d0, grad_scale * d1);
                            p.y = pack_bf16x2(grad_scale * d2, grad_scale * d3);
                            stg_stream_u2(reinterpret_cast<__nv_bfloat16*>(dy) + (int64_t)row * ld_dy + c, p);
                        } else {
                            stg_stream_f4(reinterpret_cast<float*>(dy) + (int64_t)row * ld_dy + c,
                                          make_float4(grad_scale * d0, grad_scale * d1, grad_scale * d2, grad_scale * d3));
                        }
                    }
                }
            }
            obs = n_obs; bits = n_bits;
        }
    } else {
        const int64_t total = (int64_t)B * io;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
            const int row = (int)(e / io);
            const int c = (int)(e - (int64_t)row * io);
            const int64_t obs = batch_idx ? batch_idx[row] : (int64_t)row;
            const float xv = x[obs * ld_x + c];
            const float yv = kBf16Y ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(y)[(int64_t)row * ld_y + c])
                                    : reinterpret_cast<const float*>(y)[(int64_t)row * ld_y + c];
            const float d = yv - xv, q = d * d;
            s_full += q;
            s_part += (1.f - keep_of(mask_bits[mask_id[row]], col_var[c])) * q;
            if (dy) {
                if (kDy == CODAE_BF16) reinterpret_cast<__nv_bfloat16*>(dy)[(int64_t)row * ld_dy + c] = __float2bfloat16_rn(grad_scale * d);
                else reinterpret_cast<float*>(dy)[(int64_t)row * ld_dy + c] = grad_scale * d;
            }
        }
    }
    __shared__ double scratch[32];
    __shared__ bool is_last;
    const double bf = block_sum<double>((double)s_full, scratch);
    const double bp = block_sum<double>((double)s_part, scratch);
    if (threadIdx.x == 0) {
        ws->partial[blockIdx.x][0] = bf;
        ws->partial[blockIdx.x][1] = bp;
        __threadfence();
        const unsigned int t = atomicAdd(&ws->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last CTA: fixed-order sum of the per-CTA partials (deterministic regardless of arrival order)
    double f = 0.0, p = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        f += __ldcg(&ws->partial[i][0]);
        p += __ldcg(&ws->partial[i][1]);
    }
    f = block_sum<double>(f, scratch);
    p = block_sum<double>(p, scratch);
    if (threadIdx.x == 0) {
        acc[0] += f;
        acc[1] += p;
        acc[2] += (double)B;
        acc[3] = f;
        ws->ticket = 0;  // leave the workspace ready for the next call
    }
}

// ---- tabular CombinedCriterion("mean") forward + backward: one CTA ---------------------------------
constexpr int kMixedThreads = 256;
constexpr int kMaxVar = 64;

__device__ __forceinline__ void row_softmax_stats(const float* yb, const float* xb, int size, float& lse, int& tgt) {
    float mx = yb[0], xm = xb[0];
    tgt = 0;
    for (int c = 1; c < size; ++c) {
        mx = fmaxf(mx, yb[c]);
        if (xb[c] > xm) { xm = xb[c]; tgt = c; }
    }
    float s = 0.f;
    for (int c = 0; c < size; ++c) s += expf(yb[c] - mx);
    lse = mx + logf(s);
}

__global__ void __launch_bounds__(kMixedThreads) mixed_loss_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                   int B, int io, int64_t ld, int V,
                                                                   const int32_t* __restrict__ var_pos,
                                                                   const int32_t* __restrict__ var_size,
                                                                   const int32_t* __restrict__ var_type,
                                                                   const float* __restrict__ weight, float* __restrict__ dy,
                                                                   float* __restrict__ loss_out) {
    __shared__ double scratch[32];
    __shared__ float l_var[kMaxVar];
    for (int v = 0; v < V; ++v) {
        const int p = var_pos[v], s = var_size[v];
        double part = 0.0;
        if (var_type[v] == CODAE_VAR_REGRESSION) {
            for (int e = threadIdx.x; e < B * s; e += blockDim.x) {
                const int r = e / s, c = p + e % s;
                const float d = y[(int64_t)r * ld + c] - x[(int64_t)r * ld + c];
                part += (double)(d * d);
            }
        } else {
            for (int r = threadIdx.x; r < B; r += blockDim.x) {
                float lse; int tgt;
                row_softmax_stats(y + (int64_t)r * ld + p, x + (int64_t)r * ld + p, s, lse, tgt);
                part += (double)(lse - y[(int64_t)r * ld + p + tgt]);
            }
        }
        const double tot = block_sum<double>(part, scratch);
        if (threadIdx.x == 0) {
            if (var_type[v] == CODAE_VAR_REGRESSION) l_var[v] = sqrtf((float)(tot / ((double)B * s)));
            else l_var[v] = (float)(tot / (double)B);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
        for (int v = 0; v < V; ++v) { total += weight[v] * l_var[v]; loss_out[1 + v] = l_var[v]; }
        loss_out[0] = total / (float)V;
    }
    for (int v = 0; v < V; ++v) {
        const int p = var_pos[v], s = var_size[v];
        const float wv = weight[v] / (float)V;
        if (var_type[v] == CODAE_VAR_REGRESSION) {
            const float denom = (float)B * (float)s * l_var[v];
            for (int e = threadIdx.x; e < B * s; e += blockDim.x) {
                const int r = e / s, c = p + e % s;
                const float d = y[(int64_t)r * ld + c] - x[(int64_t)r * ld + c];
                dy[(int64_t)r * ld + c] = wv * d / denom;
            }
        } else {
            for (int r = threadIdx.x; r < B; r += blockDim.x) {
                float lse; int tgt;
                const float* yb = y + (int64_t)r * ld + p;
                row_softmax_stats(yb, x + (int64_t)r * ld + p, s, lse, tgt);
                for (int c = 0; c < s; ++c)
                    dy[(int64_t)r * ld + p + c] = wv * (expf(yb[c] - lse) - (c == tgt ? 1.f : 0.f)) / (float)B;
            }
        }
    }
}

// ---- tabular monitor: de-normalise, per-row per-variable losses, per-k / partial accumulators ----------
__global__ void __launch_bounds__(kMixedThreads) mixed_monitor_kernel(
    const float* __restrict__ x, const float* __restrict__ y, int B, int io, int64_t ld, int V,
    const int32_t* __restrict__ var_pos, const int32_t* __restrict__ var_size, const int32_t* __restrict__ var_type,
    const float* __restrict__ norm_scale, const float* __restrict__ norm_min, int norm_first,
    const int32_t* __restrict__ mask_id, const uint64_t* __restrict__ mask_bits, const uint8_t* __restrict__ nb_missing,
    int k_max, float* __restrict__ out_loss, double* __restrict__ acc) {
    // phase 1: loss matrix [B, V]
    for (int e = threadIdx.x; e < B * V; e += blockDim.x) {
        const int r = e / V, v = e % V;
        const int p = var_pos[v], s = var_size[v];
        float xb[kMaxVar], yb[kMaxVar];
        const int sc = s < kMaxVar ? s : kMaxVar;
        for (int c = 0; c < sc; ++c) {
            float xv = x[(int64_t)r * ld + p + c], yv = y[(int64_t)r * ld + p + c];
            if (p + c >= norm_first) {
                const float sca = norm_scale[p + c - norm_first], mn = norm_min[p + c - norm_first];
                xv = xv * sca + mn;
                yv = yv * sca + mn;
            }
            xb[c] = xv; yb[c] = yv;
        }
        float l;
        if (var_type[v] == CODAE_VAR_REGRESSION) {
            const float d = xb[0] - yb[0];
            l = d * d;
        } else {
            float lse; int tgt;
            row_softmax_stats(yb, xb, sc, lse, tgt);
            l = lse - yb[tgt];
        }
        out_loss[e] = l;
    }
    __syncthreads();
    // phase 2: one thread per accumulator, rows in order (deterministic)
    const int n_acc = 2 * k_max * V;
    for (int a = threadIdx.x; a < n_acc; a += blockDim.x) {
        const bool part = a >= k_max * V;
        const int k = (a % (k_max * V)) / V, v = a % V;
        double s = 0.0;
        for (int r = 0; r < B; ++r) {
            const int mid = mask_id[r];
            if (nb_missing[mid] - 1 != k) continue;
            if (part && !((mask_bits[mid] >> v) & 1ull)) continue;
            s += (double)out_loss[r * V + v];
        }
        acc[2 + a] += s;
    }
    if (threadIdx.x == 0) {
        double f = 0.0, p = 0.0;
        for (int r = 0; r < B; ++r) {
            const uint64_t bits = mask_bits[mask_id[r]];
            for (int v = 0; v < V; ++v) {
                const double l = (double)out_loss[r * V + v];
                f += l;
                if ((bits >> v) & 1ull) p += l;
            }
        }
        acc[0] += f;
        acc[1] += p;
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

size_t codae_loss_workspace_bytes(const codae_ctx*) { return sizeof(LossWs); }

int codae_mse_loss_fwd_bwd(codae_ctx* ctx, const float* x, int64_t ld_x, const int64_t* batch_idx, const void* y,
                           int y_dtype, int64_t ld_y, const int32_t* mask_id, const uint64_t* mask_bits,
                           const uint8_t* col_var, int B, int io, float grad_scale, void* dy, int dy_dtype,
                           int64_t ld_dy, double* acc, void* workspace, size_t ws_bytes, void* stream) {
    CODAE_REQUIRE(ctx, ctx && x && y && mask_id && mask_bits && col_var && acc && workspace,
                  "codae_mse_loss_fwd_bwd: NULL argument");
    CODAE_REQUIRE(ctx, B >= 1 && io >= 1 && ld_x >= io && ld_y >= io && (!dy || ld_dy >= io),
                  "codae_mse_loss_fwd_bwd: bad shape B=%d io=%d", B, io);
    if (ws_bytes < sizeof(LossWs))
        return codae_fail(ctx, CODAE_ENOMEM, "codae_mse_loss_fwd_bwd: workspace %zu < %zu bytes", ws_bytes, sizeof(LossWs));
    const bool by = y_dtype == CODAE_BF16, bd = dy_dtype == CODAE_BF16 || dy_dtype == CODAE_F32X3;   // bd: 2-byte elements
    const bool x3 = dy && dy_dtype == CODAE_F32X3;
    const bool vec = (io % 4 == 0) && (ld_x % 4 == 0) && (ld_y % 4 == 0) && (!dy || ld_dy % 4 == 0) && aligned16(x) &&
                     ((reinterpret_cast<uintptr_t>(y) & (by ? 7 : 15)) == 0) &&
                     (!dy || (reinterpret_cast<uintptr_t>(dy) & (bd ? 7 : 15)) == 0) &&
                     ((reinterpret_cast<uintptr_t>(col_var) & 3) == 0);
    int log2g = 5;
    while (log2g < 8 && (kLossUnroll << log2g) < io / 4) ++log2g;
    const int rows_per_cta = kLossThreads >> log2g;
    int64_t blocks = vec ? ((int64_t)B + rows_per_cta - 1) / rows_per_cta : ((int64_t)B * io + kLossThreads * 4 - 1) / (kLossThreads * 4);
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks > kLossMaxGrid) blocks = kLossMaxGrid;
    if (blocks < 1) blocks = 1;
    LossWs* ws = reinterpret_cast<LossWs*>(workspace);
    cudaStream_t s = as_stream(stream);
#define LAUNCH(BY, BD, VEC)                                                                                         \
    do {                                                                                                            \
        static int occ = 0;                                                                                         \
        if (!occ) occ = resident_ctas_per_sm(mse_loss_kernel<BY, BD, VEC>, kLossThreads, 0);                        \
        if (blocks > (int64_t)ctx->sm_count * occ) blocks = (int64_t)ctx->sm_count * occ;   /* one resident wave */ \
        launch_pdl(ctx, mse_loss_kernel<BY, BD, VEC>, dim3((unsigned)blocks), dim3(kLossThreads), 0, s, x, ld_x, batch_idx, y, \
                   ld_y, mask_id, mask_bits, col_var, B, io, log2g, grad_scale, dy, ld_dy, acc, ws);                \
    } while (0)
    if (x3) {
        if (!vec || by)
            return codae_fail(ctx, CODAE_EINVAL, "codae_mse_loss_fwd_bwd: CODAE_F32X3 dy needs f32 y, io and all pitches multiples of 4, aligned buffers");
        LAUNCH(false, CODAE_F32X3, true);
    } else if (vec) {
        if (by && bd) LAUNCH(true, CODAE_BF16, true);
        else if (by) LAUNCH(true, CODAE_F32, true);
        else if (bd) LAUNCH(false, CODAE_BF16, true);
        else LAUNCH(false, CODAE_F32, true);
    } else {
        if (by && bd) LAUNCH(true, CODAE_BF16, false);
        else if (by) LAUNCH(true, CODAE_F32, false);
        else if (bd) LAUNCH(false, CODAE_BF16, false);
        else LAUNCH(false, CODAE_F32, false);
    }
#undef LAUNCH
    return codae_check_launch(ctx, "mse_loss_kernel");
}

int codae_mixed_loss_fwd_bwd(codae_ctx* ctx, const float* x, const float* y, int B, int io, int64_t ld, int V,
                             const int32_t* var_pos, const int32_t* var_size, const int32_t* var_type,
                             const float* weight, float* dy, float* loss_out, void* stream) {
    CODAE_REQUIRE(ctx, ctx && x && y && var_pos && var_size && var_type && weight && dy && loss_out,
                  "codae_mixed_loss_fwd_bwd: NULL argument");
    CODAE_REQUIRE(ctx, B >= 1 && io >= 1 && ld >= io && V >= 1 && V <= kMaxVar, "codae_mixed_loss_fwd_bwd: bad shape (V <= %d)", kMaxVar);
    mixed_loss_kernel<<<1, kMixedThreads, 0, as_stream(stream)>>>(x, y, B, io, ld, V, var_pos, var_size, var_type, weight,
                                                                  dy, loss_out);
    return codae_check_launch(ctx, "mixed_loss_kernel");
}

int codae_mixed_monitor(codae_ctx* ctx, const float* x, const float* y, int B, int io, int64_t ld, int V,
                        const int32_t* var_pos, const int32_t* var_size, const int32_t* var_type,
                        const float* norm_scale, const float* norm_min, int norm_first, const int32_t* mask_id,
                        const uint64_t* mask_bits, const uint8_t* nb_missing, int k_max, float* out_loss, double* acc,
                        void* stream) {
    CODAE_REQUIRE(ctx, ctx && x && y && var_pos && var_size && var_type && mask_id && mask_bits && nb_missing && out_loss && acc,
                  "codae_mixed_monitor: NULL argument");
    CODAE_REQUIRE(ctx, B >= 1 && io >= 1 && ld >= io && V >= 1 && V <= kMaxVar && k_max >= 1, "codae_mixed_monitor: bad shape");
    CODAE_REQUIRE(ctx, norm_first >= io || (norm_scale && norm_min), "codae_mixed_monitor: normalizer tables missing");
    mixed_monitor_kernel<<<1, kMixedThreads, 0, as_stream(stream)>>>(x, y, B, io, ld, V, var_pos, var_size, var_type,
                                                                     norm_scale, norm_min, norm_first, mask_id, mask_bits,
                                                                     nb_missing, k_max, out_loss, acc);
    return codae_check_launch(ctx, "mixed_monitor_kernel");
}

}  // extern "C"
