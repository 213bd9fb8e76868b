// K2 ABI -- nn.Linear forward / input-gradient / weight-gradient, dispatched on (dtype, shape):
//   CODAE_F32  -> exact-fp32 FFMA engine (gemm_simt.cu)
//   CODAE_BF16 -> tcgen05 tensor-core engine (gemm_tcgen05.cu).  Shapes it cannot tile return CODAE_EINVAL:
//                 there is no silent fallback; the host layer picks CODAE_F32 for tabular widths up front.
//   codae_linear_*_x3 -> the same tensor-core kernels on CODAE_F32X3 operands (three bf16 planes per fp32 tensor, six MMAs
//                 per k-step): the reference's fp32 precision on tensor cores.
#include "common.cuh"
#include "gemm.h"

namespace {
inline bool tc_shape_ok(int M, int N, int K) { return M >= 1 && N >= 32 && K >= 32; }  // layer dims; the batch may be 1
}

extern "C" {

int codae_linear_engine(const codae_ctx* ctx, int dtype, int M, int N, int K) {
    if (!ctx) return CODAE_EINVAL;
    if (dtype == CODAE_BF16 && ctx->encode_tiled && tc_shape_ok(M, N, K))   // pitches/alignment are checked per call
        return CODAE_ENGINE_TCGEN05_BF16;
    if (dtype == CODAE_F32X3 && ctx->encode_tiled && tc_shape_ok(M, N, K)) return CODAE_ENGINE_TCGEN05_F32X3;
    return CODAE_ENGINE_SIMT_F32;
}

int codae_linear_fwd(codae_ctx* ctx, const void* X, int64_t ldx, const void* W, int64_t ldw, const float* bias, void* Y,
                     int64_t ldy, int M, int N, int K, int act, int dtype, int out_dtype, void* stream) {
    CODAE_REQUIRE(ctx, ctx && X && W && Y, "codae_linear_fwd: NULL argument");
    CODAE_REQUIRE(ctx, M >= 1 && N >= 1 && K >= 1 && ldx >= K && ldw >= K && ldy >= N, "codae_linear_fwd: bad shape M=%d N=%d K=%d", M, N, K);
    CODAE_REQUIRE(ctx, act == CODAE_ACT_NONE || act == CODAE_ACT_RELU, "codae_linear_fwd: bad activation %d", act);
    if (dtype == CODAE_F32) {
        CODAE_REQUIRE(ctx, out_dtype == CODAE_F32, "codae_linear_fwd: f32 engine writes f32");
        return codae_simt_linear_fwd(ctx, (const float*)X, ldx, (const float*)W, ldw, bias, (float*)Y, ldy, M, N, K, act, as_stream(stream));
    }
    CODAE_REQUIRE(ctx, dtype == CODAE_BF16, "codae_linear_fwd: bad dtype %d", dtype);
    Tc05Gemm g{X, ldx, true, W, ldw, true, Y, ldy, out_dtype, M, N, K, bias, act, nullptr, 0, true};
    return codae_tc05_gemm(ctx, g, as_stream(stream));
}

int codae_linear_dgrad(codae_ctx* ctx, const void* dY, int64_t lddy, const void* W, int64_t ldw, const void* A_prev, int64_t lda,
                       void* dX, int64_t lddx, int M, int N, int K, int dtype, int out_dtype, void* stream) {
    CODAE_REQUIRE(ctx, ctx && dY && W && dX, "codae_linear_dgrad: NULL argument");
    CODAE_REQUIRE(ctx, M >= 1 && N >= 1 && K >= 1 && lddy >= N && ldw >= K && lddx >= K && (!A_prev || lda >= K),
                  "codae_linear_dgrad: bad shape M=%d N=%d K=%d", M, N, K);
    if (dtype == CODAE_F32) {
        CODAE_REQUIRE(ctx, out_dtype == CODAE_F32, "codae_linear_dgrad: f32 engine writes f32");
        return codae_simt_linear_dgrad(ctx, (const float*)dY, lddy, (const float*)W, ldw, (const float*)A_prev, lda, (float*)dX, lddx, M, N, K, as_stream(stream));
    }
    CODAE_REQUIRE(ctx, dtype == CODAE_BF16, "codae_linear_dgrad: bad dtype %d", dtype);
    // dX[M,K] = sum_n dY[m,n] W[n,k]: contraction over N.  A = dY (contraction contiguous), B(k_out, n) = W[n, k_out]
    // (contraction index is the row of W -> MN-major operand).
    Tc05Gemm g{dY, lddy, true, W, ldw, false, dX, lddx, out_dtype, M, K, N, nullptr, CODAE_ACT_NONE, A_prev, lda, true};
    return codae_tc05_gemm(ctx, g, as_stream(stream));
}

int codae_linear_wgrad(codae_ctx* ctx, const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw,
                       float* db, int M, int N, int K, int dtype, void* stream) {
    CODAE_REQUIRE(ctx, ctx && dY && X && dW, "codae_linear_wgrad: NULL argument");
    CODAE_REQUIRE(ctx, M >= 1 && N >= 1 && K >= 1 && lddy >= N && ldx >= K && lddw >= K, "codae_linear_wgrad: bad shape M=%d N=%d K=%d", M, N, K);
    int rc;
    if (dtype == CODAE_F32) {
        rc = codae_simt_linear_wgrad(ctx, (const float*)dY, lddy, (const float*)X, ldx, dW, lddw, M, N, K, as_stream(stream));
    } else {
        CODAE_REQUIRE(ctx, dtype == CODAE_BF16, "codae_linear_wgrad: bad dtype %d", dtype);
        // dW[N,K] = sum_m dY[m,n] X[m,k]: contraction over the batch; both operands are MN-major.
        Tc05Gemm g{dY, lddy, false, X, ldx, false, dW, lddw, CODAE_F32, N, K, M, nullptr, CODAE_ACT_NONE, nullptr, 0};
        rc = codae_tc05_gemm(ctx, g, as_stream(stream));
    }
    if (rc) return rc;
    if (db) rc = codae_colsum(ctx, dY, dtype, lddy, M, N, db, as_stream(stream));
    return rc;
}

int codae_linear_wgrad_sq_slots(const codae_ctx* ctx, int M, int N, int K, int dtype) {
    if (!ctx || (dtype != CODAE_BF16 && dtype != CODAE_F32X3) || !tc_shape_ok(M, N, K)) return 0;
    Tc05Gemm g{nullptr, 0, false, nullptr, 0, false, nullptr, 0, CODAE_F32, N, K, M, nullptr, CODAE_ACT_NONE, nullptr, 0};
    g.planes = dtype == CODAE_F32X3 ? 3 : 1;
    return codae_tc05_gemm_ctas(ctx, g);
}

int codae_linear_wgrad_sq(codae_ctx* ctx, const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw,
                          int M, int N, int K, int dtype, double* sq_partials, int n_slots, void* stream) {
    CODAE_REQUIRE(ctx, ctx && dY && X && dW && sq_partials, "codae_linear_wgrad_sq: NULL argument");
    CODAE_REQUIRE(ctx, M >= 1 && N >= 1 && K >= 1 && lddy >= N && ldx >= K && lddw >= K, "codae_linear_wgrad_sq: bad shape M=%d N=%d K=%d", M, N, K);
    CODAE_REQUIRE(ctx, dtype == CODAE_BF16, "codae_linear_wgrad_sq: tensor-core engine only (dtype %d)", dtype);
    CODAE_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(sq_partials) & 7) == 0, "codae_linear_wgrad_sq: sq_partials must be 8-byte aligned");
    Tc05Gemm g{dY, lddy, false, X, ldx, false, dW, lddw, CODAE_F32, N, K, M, nullptr, CODAE_ACT_NONE, nullptr, 0};
    g.sq_partial = sq_partials;
    g.sq_slots = n_slots;
    return codae_tc05_gemm(ctx, g, as_stream(stream));
}

// ---- fp32-parity engine: CODAE_F32X3 operands (three bf16 planes), same tensor-core kernels with six MMAs per k-step ----------
int codae_linear_fwd_x3(codae_ctx* ctx, const void* X, int64_t ldx, int64_t x_plane, const void* W, int64_t ldw, int64_t w_plane,
                        void* Y, int64_t ldy, int64_t y_plane, int M, int N, int K, int act, int out_dtype, void* stream) {
    CODAE_REQUIRE(ctx, ctx && X && W && Y, "codae_linear_fwd_x3: NULL argument");
    CODAE_REQUIRE(ctx, M >= 1 && N >= 1 && K >= 1 && ldx >= K && ldw >= K && ldy >= N, "codae_linear_fwd_x3: bad shape M=%d N=%d K=%d", M, N, K);
    CODAE_REQUIRE(ctx, act == CODAE_ACT_NONE || act == CODAE_ACT_RELU, "codae_linear_fwd_x3: bad activation %d", act);
    CODAE_REQUIRE(ctx, out_dtype == CODAE_F32 || out_dtype == CODAE_F32X3, "codae_linear_fwd_x3: output is f32 or its bf16 triple");
    Tc05Gemm g{X, ldx, true, W, ldw, true, Y, ldy, out_dtype, M, N, K, nullptr, act, nullptr, 0, true};
    g.planes = 3; g.a_plane_stride = x_plane; g.b_plane_stride = w_plane; g.c_plane_stride = y_plane;
    return codae_tc05_gemm(ctx, g, as_stream(stream));
}

int codae_linear_dgrad_x3(codae_ctx* ctx, const void* dY, int64_t lddy, int64_t dy_plane, const void* W, int64_t ldw,
                          int64_t w_plane, const void* A_prev_hi, int64_t lda, void* dX, int64_t lddx, int64_t dx_plane, int M,
                          int N, int K, int out_dtype, void* stream) {
    CODAE_REQUIRE(ctx, ctx && dY && W && dX, "codae_linear_dgrad_x3: NULL argument");
    CODAE_REQUIRE(ctx, M >= 1 && N >= 1 && K >= 1 && lddy >= N && ldw >= K && lddx >= K && (!A_prev_hi || lda >= K),
                  "codae_linear_dgrad_x3: bad shape M=%d N=%d K=%d", M, N, K);
    CODAE_REQUIRE(ctx, out_dtype == CODAE_F32 || out_dtype == CODAE_F32X3, "codae_linear_dgrad_x3: output is f32 or its bf16 triple");
    Tc05Gemm g{dY, lddy, true, W, ldw, false, dX, lddx, out_dtype, M, K, N, nullptr, CODAE_ACT_NONE, A_prev_hi, lda, true};
    g.planes = 3; g.a_plane_stride = dy_plane; g.b_plane_stride = w_plane; g.c_plane_stride = dx_plane;
    return codae_tc05_gemm(ctx, g, as_stream(stream));
}

int codae_linear_wgrad_x3(codae_ctx* ctx, const void* dY, int64_t lddy, int64_t dy_plane, const void* X, int64_t ldx,
                          int64_t x_plane, float* dW, int64_t lddw, int M, int N, int K, double* sq_partials, int n_slots,
                          void* stream) {
    CODAE_REQUIRE(ctx, ctx && dY && X && dW, "codae_linear_wgrad_x3: NULL argument");
    CODAE_REQUIRE(ctx, M >= 1 && N >= 1 && K >= 1 && lddy >= N && ldx >= K && lddw >= K, "codae_linear_wgrad_x3: bad shape M=%d N=%d K=%d", M, N, K);
    CODAE_REQUIRE(ctx, !sq_partials || (reinterpret_cast<uintptr_t>(sq_partials) & 7) == 0, "codae_linear_wgrad_x3: sq_partials must be 8-byte aligned");
    Tc05Gemm g{dY, lddy, false, X, ldx, false, dW, lddw, CODAE_F32, N, K, M, nullptr, CODAE_ACT_NONE, nullptr, 0};
    g.planes = 3; g.a_plane_stride = dy_plane; g.b_plane_stride = x_plane;
    if (sq_partials) { g.sq_partial = sq_partials; g.sq_slots = n_slots; }
    return codae_tc05_gemm(ctx, g, as_stream(stream));
}

}  // extern "C"
