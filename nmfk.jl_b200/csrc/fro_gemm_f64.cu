// Variant FRO, Float64: the stacked-restart GEMM  C[M x N] = A[M x K] B[N x K]^T  (see fro_gemm.cu for what M, N, K are) on the
// FP64 tensor pipe.  tcgen05 has no FP64 kind, so this is mma.sync.aligned.m8n8k4.f64 (DMMA) fed from shared memory:
// CTA tile 128 x 128, K-blocks of 16 doubles staged by a 3-stage cp.async ring (16-byte copies; both operands are K-major,
// rows padded to a pitch of 20 doubles = 4 x odd so that the 64-bit fragment loads of a half warp hit 32 distinct banks),
// 8 warps as 2 x 4, each warp owns 64 x 32 of the tile = 8 x 4 DMMA tiles (64 accumulator registers per thread).
#include "fro.h"

namespace nmfk {
namespace {

constexpr int FBM = 128, FBN = 128, FBK = 16, FST = 3, FPITCH = FBK + 4;
constexpr int FTHREADS = 256;
constexpr size_t FSMEM = (size_t)FST * (FBM + FBN) * FPITCH * sizeof(double);

__device__ __forceinline__ void cp16(void* dst, const void* src, bool ok) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp8(void* dst, const void* src, bool ok) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(src), "r"(bytes) : "memory");
}

template <bool ALIGNED>
__global__ void __launch_bounds__(FTHREADS, 1)
    fro_gemm_f64_kernel(const double* __restrict__ A, long long lda, const double* __restrict__ B, long long ldb, double* __restrict__ C,
                        long long ldc, int M, int N, int K, int mtiles) {
    extern __shared__ __align__(16) unsigned char fsm[];
    double* As = reinterpret_cast<double*>(fsm);                 // [FST][FBM][FPITCH]
    double* Bs = As + (size_t)FST * FBM * FPITCH;                // [FST][FBN][FPITCH]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;                     // 2 x 4 warps: 64 rows x 32 columns each
    const int m0 = (blockIdx.x % mtiles) * FBM, n0 = (blockIdx.x / mtiles) * FBN;  // M fastest: concurrent CTAs share B rows in L2
    const int kblocks = (K + FBK - 1) / FBK;

    auto issue = [&](int kb) {
        const int s = kb % FST;
        double* as = As + (size_t)s * FBM * FPITCH;
        double* bs = Bs + (size_t)s * FBN * FPITCH;
        const int k0 = kb * FBK;
        if (ALIGNED) {  // 16-byte copies: 8 per row, 128 rows per operand -> 1024 copies each, 4 per thread
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = tid + q * FTHREADS, row = e >> 3, c2 = (e & 7) * 2;
                const bool kok = k0 + c2 < K;
                cp16(as + row * FPITCH + c2, A + (long long)(m0 + row) * lda + k0 + c2, kok && (m0 + row) < M);
                cp16(bs + row * FPITCH + c2, B + (long long)(n0 + row) * ldb + k0 + c2, kok && (n0 + row) < N);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int e = tid + q * FTHREADS, row = e >> 4, c = e & 15;
                const bool kok = k0 + c < K;
                cp8(as + row * FPITCH + c, A + (long long)(m0 + row) * lda + k0 + c, kok && (m0 + row) < M);
                cp8(bs + row * FPITCH + c, B + (long long)(n0 + row) * ldb + k0 + c, kok && (n0 + row) < N);
            }
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int p = 0; p < FST - 1; ++p) {
        if (p < kblocks) issue(p);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int fr = lane >> 2, fk = lane & 3;  // fragment row / column index and k index of this lane
    for (int kb = 0; kb < kblocks; ++kb) {
        asm volatile("cp.async.wait_group %0;" ::"n"(FST - 2) : "memory");
        __syncthreads();
        if (kb + FST - 1 < kblocks) issue(kb + FST - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* as = As + (size_t)(kb % FST) * FBM * FPITCH + (size_t)(wm * 64 + fr) * FPITCH + fk;
        const double* bs = Bs + (size_t)(kb % FST) * FBN * FPITCH + (size_t)(wn * 32 + fr) * FPITCH + fk;
#pragma unroll
        for (int k4 = 0; k4 < FBK; k4 += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = as[(size_t)i * 8 * FPITCH + k4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = bs[(size_t)j * 8 * FPITCH + k4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                                 : "d"(af[i]), "d"(bf[j]));
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // C fragment of m8n8k4: lane holds row = lane / 4, columns 2 * (lane % 4) and + 1
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + wm * 64 + i * 8 + fr;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + wn * 32 + j * 8 + fk * 2;
            if (col < N) C[(long long)row * ldc + col] = acc[i][j][0];
            if (col + 1 < N) C[(long long)row * ldc + col + 1] = acc[i][j][1];
        }
    }
}

}  // namespace

cudaError_t launch_fro_gemm_f64(const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc, int M, int N,
                                int K, cudaStream_t s) {
    const int mtiles = (M + FBM - 1) / FBM, ntiles = (N + FBN - 1) / FBN;
    const long long grid = (long long)mtiles * ntiles;
    if (grid > 2147483647ll) return cudaErrorInvalidValue;
    const bool aligned = (lda % 2 == 0) && (ldb % 2 == 0) && (K % 2 == 0) && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) % 16 == 0);
    cudaError_t e;
    if (aligned) {
        if ((e = cudaFuncSetAttribute(fro_gemm_f64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FSMEM)) != cudaSuccess) return e;
        fro_gemm_f64_kernel<true><<<(unsigned)grid, FTHREADS, FSMEM, s>>>(A, lda, B, ldb, C, ldc, M, N, K, mtiles);
    } else {
        if ((e = cudaFuncSetAttribute(fro_gemm_f64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FSMEM)) != cudaSuccess) return e;
        fro_gemm_f64_kernel<false><<<(unsigned)grid, FTHREADS, FSMEM, s>>>(A, lda, B, ldb, C, ldc, M, N, K, mtiles);
    }
    return cudaGetLastError();
}

}  // namespace nmfk
