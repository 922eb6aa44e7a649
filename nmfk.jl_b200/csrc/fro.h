// Variant FRO (NMF.jl MultUpdate(obj=:mse), reached by the reference at NMFkExecute.jl:763-766): internal interface.
#pragma once
#include "nmfk_internal.h"

namespace nmfk {

// Float32 stacked GEMM on tcgen05 (fro_gemm.cu): C[M x N] = A[M x K] B[N x K]^T, 3-term TF32 split, all row-major
bool fro_gemm_supported(long long N, long long K);
int fro_gemm_slices(int M, int N, int K);  // split-K slices the launcher should be given for this shape
cudaError_t launch_fro_gemm(const float* Ahi, const float* Alo, long long lda, const float* Bhi, const float* Blo, long long ldb, float* C,
                            long long ldc, int M, int N, int K, int S, long long pstride, int* d_errflag, cudaStream_t s);
// Float64 stacked GEMM on the FP64 tensor pipe (fro_gemm_f64.cu): same contract, DMMA m8n8k4
cudaError_t launch_fro_gemm_f64(const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc, int M, int N,
                                int K, cudaStream_t s);

// C[i] = sum over the S split-K partial products C[s * pstride + i], in slice order
cudaError_t launch_sum_slices(float* C, int S, long long pstride, long long len, cudaStream_t s);
// lo image of an FP32 array for the 3-term split: lo = x - tf32_truncate(x)
cudaError_t launch_split_lo(const float* x, float* lo, long long len, cudaStream_t s);

// the whole solve (fro_solve.cu); Xlo / Xtlo: lo images of a.X / a.Xt (Float32 only, else nullptr)
cudaError_t solve_fro(const SolveArgs& a, int dtype, const void* Xlo, const void* Xtlo, cudaStream_t s, int64_t* launches);

// NMFsparsity (sparsity.cu): beta-divergence multiplicative updates with an L1 penalty on H and unit-norm columns of W
cudaError_t solve_sparsity(const SolveArgs& a, int dtype, double beta, double sparsity, double lam, cudaStream_t s, int64_t* launches);

}  // namespace nmfk
