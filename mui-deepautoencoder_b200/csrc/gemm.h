// Internal (non-ABI) entry points of the two GEMM engines; linear.cu dispatches between them.
#pragma once
#include "common.cuh"

int codae_simt_linear_fwd(codae_ctx* ctx, const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                          float* Y, int64_t ldy, int M, int N, int K, int act, cudaStream_t s);
int codae_simt_linear_dgrad(codae_ctx* ctx, const float* dY, int64_t lddy, const float* W, int64_t ldw, const float* A_prev,
                            int64_t lda, float* dX, int64_t lddx, int M, int N, int K, cudaStream_t s);
int codae_simt_linear_wgrad(codae_ctx* ctx, const float* dY, int64_t lddy, const float* X, int64_t ldx, float* dW,
                            int64_t lddw, int M, int N, int K, cudaStream_t s);
int codae_colsum(codae_ctx* ctx, const void* dY, int dtype, int64_t ld, int M, int N, float* db, cudaStream_t s);

// tcgen05 engine (gemm_tcgen05.cu).  All operands bf16, fp32 accumulation in TMEM.
//   a_kmajor / b_kmajor: whether the contraction index is the contiguous one of that operand.
//   C[M,N] (f32 or bf16, pitch ldc) = epilogue(sum_k A(m,k) B(n,k))
struct Tc05Gemm {
    const void* A; int64_t lda; bool a_kmajor;   // A(m,k): kmajor -> A[m*lda+k], else A[k*lda+m]
    const void* B; int64_t ldb; bool b_kmajor;   // B(n,k): kmajor -> B[n*ldb+k], else B[k*ldb+n]
    void* C; int64_t ldc; int c_dtype;
    int M, N, K;
    const float* bias; int act;                  // bias[n] + activation (fwd)
    const void* mask_src; int64_t ldm;           // bf16 [M, ldm]: C *= (mask_src > 0) (dgrad)
    bool b_is_weight = false;                    // B holds layer weights: its tiles may be fetched ahead of the stream dependency
    double* sq_partial = nullptr;                // f32 outputs: every CTA writes the sum of squares of what it stored into its slot
    int sq_slots = 0;                            // must equal codae_tc05_gemm_ctas(ctx, g)
    // fp32-parity mode (CODAE_F32X3): A and B are THREE bf16 planes each (hi, mid, lo: x = hi + mid + lo to 2^-24), plane p at
    // base + p * plane_stride elements; six MMAs per k-step (hh | hm, mh | mm, hl, lh) into three TMEM accumulators by
    // magnitude class, summed in the epilogue.  C is f32 (c_dtype CODAE_F32) or three bf16 planes (c_dtype CODAE_F32X3).
    int planes = 1;
    int64_t a_plane_stride = 0, b_plane_stride = 0, c_plane_stride = 0;
};
bool codae_tc05_supported(const codae_ctx* ctx, const Tc05Gemm& g);
int codae_tc05_gemm(codae_ctx* ctx, const Tc05Gemm& g, cudaStream_t s);
// CTAs codae_tc05_gemm launches for this shape under the context's current options (0: shape not supported).
int codae_tc05_gemm_ctas(const codae_ctx* ctx, const Tc05Gemm& g);
