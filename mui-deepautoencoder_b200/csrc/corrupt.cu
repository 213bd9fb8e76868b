// K1 -- corruption: Philox mask-id table, fused row gather + slot-mask corruption, dense masks.
// HBM-bound elementwise work: 128-bit streaming loads/stores, one mask id per row looked up from the
// (observation, run) table; the dense mask is computed from (mask_bits, col_var) and never stored.
#include "common.cuh"

namespace {

// ---- Philox4x32-10 (Salmon et al., SC'11); the host restatement is oracle/philox.py ---------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

constexpr int kMaxRun = 1024;

// One thread per observation: Fisher-Yates over range(nb_run) held in the output row itself.
__global__ void mask_table_philox_kernel(uint64_t seed, int64_t first_obs, int64_t n_obs, int nb_run,
                                         int16_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_obs) return;
    int16_t* row = out + i * nb_run;
    for (int t = 0; t < nb_run; ++t) row[t] = (int16_t)t;
    const uint64_t obs = (uint64_t)(first_obs + i);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t blk[4];
    for (int t = 0; t + 1 < nb_run; ++t) {
        if ((t & 3) == 0) philox4x32_10((uint32_t)obs, (uint32_t)(obs >> 32), (uint32_t)(t >> 2), 0x434F4441u, k0, k1, blk);
        const uint32_t draw = blk[t & 3];
        const int j = t + (int)(((uint64_t)draw * (uint64_t)(nb_run - t)) >> 32);
        const int16_t a = row[t];
        row[t] = row[j];
        row[j] = a;
    }
}

__device__ __forceinline__ float keep_of(uint64_t bits, uint32_t var) { return ((bits >> var) & 1ull) ? 0.0f : 1.0f; }
// The row gather is the one place of the step where a caller-supplied index addresses memory: an observation id outside
// [0, n_rows) traps (reported as a launch failure) instead of reading somebody else's bytes -- the reference raises IndexError.
__device__ __forceinline__ int64_t checked_obs(int64_t obs, int64_t n_rows) {
    if ((uint64_t)obs >= (uint64_t)n_rows) __trap();
    return obs;
}

// Vector path: io % 4 == 0, all pitches % 4 == 0, 16-byte aligned bases.
// A group of G = 2^log2g threads (32..256) owns one row at a time: the row's (observation, mask id, mask bits) chain is
// looked up once per row -- and for the NEXT row while the current row's data is in flight -- and every thread keeps
// four independent 128-bit loads outstanding.
constexpr int kRowUnroll = 4;
// kOut: CODAE_F32 / CODAE_BF16 / CODAE_F32X3 (three bf16 planes B * ld_cx elements apart: the operand of the fp32-parity engine)
template <int kOut>
__global__ void __launch_bounds__(256) corrupt_fwd_vec_kernel(const float* __restrict__ data, int64_t n_rows, int64_t ld_data,
                                                              const int64_t* __restrict__ batch_idx, int B,
                                                              const int16_t* __restrict__ mask_table, int nb_run, int run,
                                                              const uint64_t* __restrict__ mask_bits,
                                                              const uint8_t* __restrict__ col_var, int io4, int log2g,
                                                              void* __restrict__ out_cx, int64_t ld_cx,
                                                              float* __restrict__ out_x, int64_t ld_x,
                                                              int32_t* __restrict__ out_mask_id) {
    pdl_launch_dependents();
    pdl_wait();
    const int G = 1 << log2g, rows_per_cta = 256 >> log2g;
    const int gl = threadIdx.x & (G - 1);
    const int row_stride = gridDim.x * rows_per_cta;
    int row = blockIdx.x * rows_per_cta + (threadIdx.x >> log2g);
    int64_t obs = 0;
    int mid = 0;
    uint64_t bits = 0;
    if (row < B) {
        obs = checked_obs(batch_idx ? batch_idx[row] : (int64_t)row, n_rows);
        mid = mask_table[obs * nb_run + run];
        bits = mask_bits[mid];
    }
    for (; row < B; row += row_stride) {
        const float* src = data + obs * ld_data;
        int64_t n_obs = 0;
        int n_mid = 0;
        uint64_t n_bits = 0;
        bool first_chunk = true;
        for (int c0 = 0; c0 < io4; c0 += kRowUnroll * G) {
            float4 x[kRowUnroll];
#pragma unroll
            for (int j = 0; j < kRowUnroll; ++j) {
                const int c4 = c0 + gl + j * G;
                if (c4 < io4) x[j] = ldg_stream_f4(src + 4 * (int64_t)c4);
            }
            if (first_chunk) {                     // next row's lookups ride under the loads just issued
                first_chunk = false;
                const int nrow = row + row_stride;
                if (nrow < B) {
                    n_obs = checked_obs(batch_idx ? batch_idx[nrow] : (int64_t)nrow, n_rows);
                    n_mid = mask_table[n_obs * nb_run + run];
                    n_bits = mask_bits[n_mid];
                }
            }
#pragma unroll
            for (int j = 0; j < kRowUnroll; ++j) {
                const int c4 = c0 + gl + j * G;
                if (c4 >= io4) continue;
                const uchar4 var = *reinterpret_cast<const uchar4*>(col_var + 4 * c4);
                float4 cx;
                cx.x = x[j].x * keep_of(bits, var.x);
                cx.y = x[j].y * keep_of(bits, var.y);
                cx.z = x[j].z * keep_of(bits, var.z);
                cx.w = x[j].w * keep_of(bits, var.w);
                if (kOut == CODAE_F32X3) {
                    store_planes4(reinterpret_cast<__nv_bfloat16*>(out_cx) + (int64_t)row * ld_cx + 4 * c4, (int64_t)B * ld_cx, cx);
                } else if (kOut == CODAE_BF16) {
                    uint2 p;
                    p.x = pack_bf16x2(cx.x, cx.y);
                    p.y = pack_bf16x2(cx.z, cx.w);
                    stg_stream_u2(reinterpret_cast<__nv_bfloat16*>(out_cx) + (int64_t)row * ld_cx + 4 * c4, p);
                } else {
                    stg_stream_f4(reinterpret_cast<float*>(out_cx) + (int64_t)row * ld_cx + 4 * c4, cx);
                }
                if (out_x) stg_stream_f4(out_x + (int64_t)row * ld_x + 4 * c4, x[j]);
            }
        }
        if (out_mask_id && gl == 0) out_mask_id[row] = mid;
        obs = n_obs; mid = n_mid; bits = n_bits;
    }
}

// Scalar path for tabular widths (io = 11 for abalone) and unaligned pitches.
template <bool kBf16Out>
__global__ void __launch_bounds__(256) corrupt_fwd_scalar_kernel(const float* __restrict__ data, int64_t n_rows, int64_t ld_data,
                                                                 const int64_t* __restrict__ batch_idx, int B,
                                                                 const int16_t* __restrict__ mask_table, int nb_run,
                                                                 int run, const uint64_t* __restrict__ mask_bits,
                                                                 const uint8_t* __restrict__ col_var, int io,
                                                                 void* __restrict__ out_cx, int64_t ld_cx,
                                                                 float* __restrict__ out_x, int64_t ld_x,
                                                                 int32_t* __restrict__ out_mask_id) {
    const int64_t total = (int64_t)B * io;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(e / io);
        const int c = (int)(e - (int64_t)row * io);
        const int64_t obs = checked_obs(batch_idx ? batch_idx[row] : (int64_t)row, n_rows);
        const int mid = mask_table[obs * nb_run + run];
        const float x = data[obs * ld_data + c];
        const float cx = x * keep_of(mask_bits[mid], col_var[c]);
        if (kBf16Out) reinterpret_cast<__nv_bfloat16*>(out_cx)[(int64_t)row * ld_cx + c] = __float2bfloat16_rn(cx);
        else reinterpret_cast<float*>(out_cx)[(int64_t)row * ld_cx + c] = cx;
        if (out_x) out_x[(int64_t)row * ld_x + c] = x;
        if (out_mask_id && c == 0) out_mask_id[row] = mid;
    }
}

__global__ void __launch_bounds__(256) dense_masks_kernel(const int64_t* __restrict__ batch_idx, int B,
                                                          const int16_t* __restrict__ mask_table, int nb_run, int run,
                                                          const uint64_t* __restrict__ mask_bits,
                                                          const uint8_t* __restrict__ nb_missing,
                                                          const uint8_t* __restrict__ col_var, int io, int k_max,
                                                          float* __restrict__ out_masks, float* __restrict__ out_fmask) {
    const int64_t total = (int64_t)B * io;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(e / io);
        const int c = (int)(e - (int64_t)row * io);
        const int64_t obs = batch_idx ? batch_idx[row] : (int64_t)row;
        const int mid = mask_table[obs * nb_run + run];
        const float keep = keep_of(mask_bits[mid], col_var[c]);
        const int kk = nb_missing[mid] - 1;
        for (int k = 0; k < k_max; ++k) out_masks[((int64_t)k * B + row) * io + c] = (k == kk) ? keep : 0.0f;
        out_fmask[e] = keep;
    }
}

__global__ void __launch_bounds__(256) mul_mask_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                                       float* __restrict__ out, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = x[e] * m[e];
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int grid_for(const codae_ctx* ctx, int64_t work_items, int threads, int per_thread) {
    int64_t blocks = (work_items + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread);
    const int64_t cap = (int64_t)ctx->sm_count * 8;  // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

extern "C" {

int codae_mask_table_philox(codae_ctx* ctx, uint64_t seed, int64_t first_obs, int64_t n_obs, int nb_run, int16_t* out,
                            void* stream) {
    CODAE_REQUIRE(ctx, ctx && out, "codae_mask_table_philox: NULL argument");
    CODAE_REQUIRE(ctx, nb_run >= 1 && nb_run <= kMaxRun, "codae_mask_table_philox: nb_run %d outside [1, %d]", nb_run,
                  kMaxRun);
    CODAE_REQUIRE(ctx, n_obs >= 0 && first_obs >= 0, "codae_mask_table_philox: negative size");
    if (n_obs == 0) return CODAE_OK;
    const int threads = 128;
    mask_table_philox_kernel<<<(unsigned)((n_obs + threads - 1) / threads), threads, 0, as_stream(stream)>>>(
        seed, first_obs, n_obs, nb_run, out);
    return codae_check_launch(ctx, "mask_table_philox_kernel");
}

int codae_corrupt_fwd(codae_ctx* ctx, const float* data, int64_t n_rows, int64_t ld_data, const int64_t* batch_idx, int B,
                      const int16_t* mask_table, int nb_run, int run, const uint64_t* mask_bits, const uint8_t* col_var,
                      int io, void* out_cx, int cx_dtype, int64_t ld_cx, float* out_x, int64_t ld_x,
                      int32_t* out_mask_id, void* stream) {
    CODAE_REQUIRE(ctx, ctx && data && mask_table && mask_bits && col_var && out_cx, "codae_corrupt_fwd: NULL argument");
    CODAE_REQUIRE(ctx, B >= 0 && io >= 1 && n_rows >= 1 && (batch_idx || B <= n_rows), "codae_corrupt_fwd: bad shape B=%d io=%d n_rows=%lld", B, io, (long long)n_rows);
    CODAE_REQUIRE(ctx, run >= 0 && run < nb_run, "codae_corrupt_fwd: run %d outside [0, %d)", run, nb_run);
    CODAE_REQUIRE(ctx, ld_data >= io && ld_cx >= io && (!out_x || ld_x >= io), "codae_corrupt_fwd: pitch < io");
    CODAE_REQUIRE(ctx, cx_dtype == CODAE_F32 || cx_dtype == CODAE_BF16 || cx_dtype == CODAE_F32X3, "codae_corrupt_fwd: bad cx_dtype %d", cx_dtype);
    if (B == 0) return CODAE_OK;
    const bool bf = cx_dtype == CODAE_BF16;
    const bool vec = (io % 4 == 0) && (ld_data % 4 == 0) && (ld_cx % 4 == 0) && (!out_x || ld_x % 4 == 0) &&
                     aligned16(data) && aligned16(out_cx) && (!out_x || aligned16(out_x)) &&
                     ((reinterpret_cast<uintptr_t>(col_var) & 3) == 0);
    cudaStream_t s = as_stream(stream);
    if (vec) {
        // threads per row: the smallest power of two that covers the row with four 128-bit loads per thread
        int log2g = 5;
        while (log2g < 8 && (kRowUnroll << log2g) < io / 4) ++log2g;
        const int rows_per_cta = 256 >> log2g;
        static int occ_bf = 0, occ_f32 = 0, occ_x3 = 0;
        if (!occ_bf) {
            occ_bf = resident_ctas_per_sm(corrupt_fwd_vec_kernel<CODAE_BF16>, 256, 0);
            occ_f32 = resident_ctas_per_sm(corrupt_fwd_vec_kernel<CODAE_F32>, 256, 0);
            occ_x3 = resident_ctas_per_sm(corrupt_fwd_vec_kernel<CODAE_F32X3>, 256, 0);
        }
        int g = (B + rows_per_cta - 1) / rows_per_cta;
        const int cap = ctx->sm_count * (bf ? occ_bf : (cx_dtype == CODAE_F32X3 ? occ_x3 : occ_f32));
        if (g > cap) g = cap;
        if (bf) launch_pdl(ctx, corrupt_fwd_vec_kernel<CODAE_BF16>, dim3(g), dim3(256), 0, s, data, n_rows, ld_data, batch_idx, B, mask_table, nb_run, run, mask_bits, col_var, io / 4, log2g, out_cx, ld_cx, out_x, ld_x, out_mask_id);
        else if (cx_dtype == CODAE_F32X3) launch_pdl(ctx, corrupt_fwd_vec_kernel<CODAE_F32X3>, dim3(g), dim3(256), 0, s, data, n_rows, ld_data, batch_idx, B, mask_table, nb_run, run, mask_bits, col_var, io / 4, log2g, out_cx, ld_cx, out_x, ld_x, out_mask_id);
        else launch_pdl(ctx, corrupt_fwd_vec_kernel<CODAE_F32>, dim3(g), dim3(256), 0, s, data, n_rows, ld_data, batch_idx, B, mask_table, nb_run, run, mask_bits, col_var, io / 4, log2g, out_cx, ld_cx, out_x, ld_x, out_mask_id);
    } else if (cx_dtype == CODAE_F32X3) {
        return codae_fail(ctx, CODAE_EINVAL, "codae_corrupt_fwd: CODAE_F32X3 output needs io and all pitches to be multiples of 4 and 16-byte aligned buffers");
    } else {
        const int g = grid_for(ctx, (int64_t)B * io, 256, 4);
        if (bf) corrupt_fwd_scalar_kernel<true><<<g, 256, 0, s>>>(data, n_rows, ld_data, batch_idx, B, mask_table, nb_run, run, mask_bits, col_var, io, out_cx, ld_cx, out_x, ld_x, out_mask_id);
        else corrupt_fwd_scalar_kernel<false><<<g, 256, 0, s>>>(data, n_rows, ld_data, batch_idx, B, mask_table, nb_run, run, mask_bits, col_var, io, out_cx, ld_cx, out_x, ld_x, out_mask_id);
    }
    return codae_check_launch(ctx, "corrupt_fwd_kernel");
}

int codae_dense_masks(codae_ctx* ctx, const int64_t* batch_idx, int B, const int16_t* mask_table, int nb_run, int run,
                      const uint64_t* mask_bits, const uint8_t* nb_missing, const uint8_t* col_var, int io, int k_max,
                      float* out_masks, float* out_fmask, void* stream) {
    CODAE_REQUIRE(ctx, ctx && mask_table && mask_bits && nb_missing && col_var && out_masks && out_fmask,
                  "codae_dense_masks: NULL argument");
    CODAE_REQUIRE(ctx, B >= 0 && io >= 1 && k_max >= 1, "codae_dense_masks: bad shape");
    CODAE_REQUIRE(ctx, run >= 0 && run < nb_run, "codae_dense_masks: run %d outside [0, %d)", run, nb_run);
    if (B == 0) return CODAE_OK;
    dense_masks_kernel<<<grid_for(ctx, (int64_t)B * io, 256, 2), 256, 0, as_stream(stream)>>>(
        batch_idx, B, mask_table, nb_run, run, mask_bits, nb_missing, col_var, io, k_max, out_masks, out_fmask);
    return codae_check_launch(ctx, "dense_masks_kernel");
}

int codae_mul_mask(codae_ctx* ctx, const float* x, const float* mask, float* out, int64_t n, void* stream) {
    CODAE_REQUIRE(ctx, ctx && x && mask && out && n >= 0, "codae_mul_mask: bad argument");
    if (n == 0) return CODAE_OK;
    mul_mask_kernel<<<grid_for(ctx, n, 256, 4), 256, 0, as_stream(stream)>>>(x, mask, out, n);
    return codae_check_launch(ctx, "mul_mask_kernel");
}

}  // extern "C"
