// Variant FRO, Float32: the STACKED-restart GEMM of BASELINE.json's north star on the 5th-generation tensor cores.
//
//   H-update   N1[R*k x m] = [W_1 ... W_R]^T X      one GEMM over X for all R restarts       (NMF.jl W'X, for every restart)
//   W-update   N2[R*k x n] = [H_1 ; ... ; H_R] X^T  one GEMM over X^T for all R restarts      (NMF.jl X*H', transposed)
// Both are  C[M x N] = A[M x K] B[N x K]^T  with M = R*k (the stacked factor rows), N and K the two dimensions of X, and
// BOTH operands K-major: the factor stack is stored [R*k x K] row-major, X (column-major n x m) is [m x n] row-major and
// its transpose copy is [n x m] row-major - no data movement, only tensor maps.
//
// Float32 accuracy on TF32 tensor cores: the 3-term split (never plain TF32).  x = hi + lo with hi = x truncated to TF32 -
// which is what tcgen05.mma kind::tf32 does to an FP32 operand anyway, so the "hi image" IS the FP32 array itself - and
// lo = x - hi stored as a second array.  C = Alo*Bhi + Ahi*Blo + Ahi*Bhi, FP32 accumulation in tensor memory.
// tcgen05.mma accumulates with round-toward-zero (measured, tools/umma_selftest.py --timing): a chain of c accumulating
// instructions is biased by about -0.23 c ulp.  The K loop is therefore cut in CHUNKS of 8 K-blocks (96 chained
// instructions, bias ~1e-6 relative); every chunk starts a fresh tensor-memory accumulator (two of them, double-buffered)
// that the epilogue warps add into FP32 registers with round-to-nearest while the next chunk is being multiplied.
//
// Kernel (persistent, one CTA per SM, 320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B boxes of 32 floats x rows) of the four operand images
//               of a K-block into a 2-stage shared-memory ring, mbarrier transaction counts
//   warp 1      MMA issuer (one elected lane): per K-block 4 K-steps x 3 products of tcgen05.mma.cta_group::1.kind::tf32
//               M = 128, N = 256, K = 8, operands by shared-memory descriptors; tcgen05.commit frees the stage / publishes
//               the chunk accumulator
//   warps 2-9   accumulator drain + epilogue: tcgen05.ld 32x32b, FP32 adds, plain stores of the C tile
// CLUSTER = 2: two CTAs with neighbouring M tiles and the same N tile form a cluster; each loads HALF of the B (= X) tile
// and multicasts it to both (cp.async.bulk.tensor ... .multicast::cluster), halving the L2 -> SM traffic of X.
//
// The element-wise halves of the update (Gram matrices, H .*= N1 ./ (G*H + d), convergence sums) are fro_solve.cu.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "fro.h"
#include "tc_ptx.cuh"
#include "tma.cuh"

namespace nmfk {
namespace {

constexpr int BM = 128, BN = 256, BK = 32;  // BK floats = 128 bytes = one SWIZZLE_128B row
constexpr int STAGES = 2;
constexpr int CHUNK = 8;                     // K-blocks per tensor-memory accumulator (round-toward-zero chains stay short)
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // A hi | A lo | B hi | B lo
constexpr int EPI_WARPS = 8, THREADS = (2 + EPI_WARPS) * 32;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

// [rows x kdim] row-major FP32 matrix (leading dimension ld floats) -> boxes of box_rows x 32 floats, 128-byte swizzle;
// out-of-range elements read as zero (edge tiles contribute nothing)
bool make_map(CUtensorMap* map, const void* base, long long rows, long long kdim, long long ld, int box_rows) {
    return tma_make_map_f32(map, base, kdim, rows, ld, BK, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

// K-major operand tile in SWIZZLE_128B layout: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO); the leading byte
// offset is not used by swizzled K-major layouts; version 1 (sm_100), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            tc::smem_u32(dst)),
        "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tc::smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CLUSTER>
__global__ void __launch_bounds__(THREADS, 1)
    fro_gemm_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                    const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo, float* __restrict__ Cout,
                    int M, int N, int K, long long ldc, int S, long long pstride, int* errflag) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
    uint64_t* full = bars;                // [STAGES] operand images of a K-block landed (transaction bytes)
    uint64_t* empty = full + STAGES;      // [STAGES] the MMAs that read the stage have completed (in every CTA of the cluster)
    uint64_t* acc_full = empty + STAGES;  // [2] chunk accumulator complete
    uint64_t* acc_empty = acc_full + 2;   // [2] chunk accumulator drained by the 8 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = CLUSTER > 1 ? cluster_rank() : 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mAhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mAlo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mBhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mBlo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], CLUSTER);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&acc_full[b], 1);
            tc::mbar_init(&acc_empty[b], EPI_WARPS);
        }
        tc::mbar_fence_init();
    }
    if (warp == 2) tc::tmem_alloc<512>(tmem_slot);
    tc::tc_fence_before_sync();
    if (CLUSTER > 1)
        cluster_sync_all();
    else
        __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tbase = *tmem_slot;

    const int mtiles = (M + BM - 1) / BM, ntiles = (N + BN - 1) / BN;
    const int kblocks_all = (K + BK - 1) / BK;
    // work item = (M tile group, K slice, N tile), M index fastest: the CTAs running at the same time read the same rows and
    // the same K range of B (= X) through L2; a cluster takes CLUSTER neighbouring M tiles.  The K range is cut in S slices
    // (split-K) when the tiles alone cannot fill the SMs evenly; slice s writes its partial product to Cout + s * pstride and
    // the consumer (fro_apply_kernel) adds the slices in order.
    const int mgroups = (mtiles + CLUSTER - 1) / CLUSTER;
    const int ngroupsTotal = mgroups * S * ntiles;
    const int cluster_id = blockIdx.x / CLUSTER, nclusters = gridDim.x / CLUSTER;
    auto item = [&](int g, int& m0, int& n0, int& kb0, int& kb1, int& sl) {
        m0 = ((g % mgroups) * CLUSTER + (int)crank) * BM;
        sl = (g / mgroups) % S;
        n0 = (g / (mgroups * S)) * BN;
        kb0 = (int)((long long)kblocks_all * sl / S);
        kb1 = (int)((long long)kblocks_all * (sl + 1) / S);
    };

    if (warp == 0) {
        // ===== TMA producer =====
        uint32_t it = 0;
        for (int g = cluster_id; g < ngroupsTotal; g += nclusters) {
            int m0, n0, kb0, kb1, sl;
            item(g, m0, n0, kb0, kb1, sl);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1, errflag, 1);
                if (tc::elect_one()) {
                    unsigned char* st = smem + (size_t)s * STAGE_BYTES;
                    tc::mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
                    tma_load_2d(st, &mAhi, kb * BK, m0, &full[s]);
                    tma_load_2d(st + A_BYTES, &mAlo, kb * BK, m0, &full[s]);
                    if (CLUSTER == 1) {
                        tma_load_2d(st + 2 * A_BYTES, &mBhi, kb * BK, n0, &full[s]);
                        tma_load_2d(st + 2 * A_BYTES + B_BYTES, &mBlo, kb * BK, n0, &full[s]);
                    } else {
                        const int half = BN / CLUSTER;
                        const uint16_t mask = (uint16_t)((1u << CLUSTER) - 1);
                        tma_load_2d_mc(st + 2 * A_BYTES + crank * half * BK * 4, &mBhi, kb * BK, n0 + crank * half, &full[s], mask);
                        tma_load_2d_mc(st + 2 * A_BYTES + B_BYTES + crank * half * BK * 4, &mBlo, kb * BK, n0 + crank * half, &full[s], mask);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = tc::idesc_tf32(BM, BN, 0);
        uint32_t it = 0, chunk_no = 0;
        for (int g = cluster_id; g < ngroupsTotal; g += nclusters) {
            int m0, n0, kb0, kb1, sl;
            item(g, m0, n0, kb0, kb1, sl);
            const int nchunks = (kb1 - kb0 + CHUNK - 1) / CHUNK;
            for (int c = 0; c < nchunks; ++c, ++chunk_no) {
                const int buf = chunk_no & 1;
                tc::mbar_wait(&acc_empty[buf], ((chunk_no >> 1) & 1) ^ 1, errflag, 2);
                tc::tc_fence_after_sync();
                const int kb_begin = kb0 + c * CHUNK, kb_end = min(kb1, kb_begin + CHUNK);
                for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                    const int s = it % STAGES;
                    tc::mbar_wait(&full[s], (it / STAGES) & 1, errflag, 3);
                    tc::tc_fence_after_sync();
                    if (tc::elect_one()) {
                        const uint32_t sa = tc::smem_u32(smem + (size_t)s * STAGE_BYTES);
                        const uint64_t ahi = sw128_desc(sa), alo = sw128_desc(sa + A_BYTES);
                        const uint64_t bhi = sw128_desc(sa + 2 * A_BYTES), blo = sw128_desc(sa + 2 * A_BYTES + B_BYTES);
                        const uint32_t d = tbase + buf * BN;
#pragma unroll
                        for (int kk = 0; kk < BK / 8; ++kk) {
                            const uint64_t off = (uint64_t)(kk * 2);  // 8 floats = 32 bytes along K inside the swizzle row
                            tc::mma_tf32_ss(d, alo + off, bhi + off, idesc, (kb > kb_begin || kk > 0) ? 1u : 0u);
                            tc::mma_tf32_ss(d, ahi + off, blo + off, idesc, 1u);
                            tc::mma_tf32_ss(d, ahi + off, bhi + off, idesc, 1u);
                        }
                        if (CLUSTER == 1)
                            tc::mma_commit(&empty[s]);
                        else
                            commit_mc(&empty[s], (uint16_t)((1u << CLUSTER) - 1));
                        if (kb == kb_end - 1) tc::mma_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== accumulator drain + epilogue (warps 2..9) =====
        const int e = warp - 2;
        const int quarter = warp & 3;   // the 32 tensor-memory lanes this warp may touch
        const int half = e >> 2;        // which 128 of the 256 accumulator columns
        const uint32_t lane_addr = tbase + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * (BN / 2));
        float acc[BN / 2];
        uint32_t chunk_no = 0;
        for (int g = cluster_id; g < ngroupsTotal; g += nclusters) {
            int m0, n0, kb0, kb1, sl;
            item(g, m0, n0, kb0, kb1, sl);
            const int nchunks = (kb1 - kb0 + CHUNK - 1) / CHUNK;
#pragma unroll
            for (int j = 0; j < BN / 2; ++j) acc[j] = 0.f;
            for (int c = 0; c < nchunks; ++c, ++chunk_no) {
                const int buf = chunk_no & 1;
                tc::mbar_wait(&acc_full[buf], (chunk_no >> 1) & 1, errflag, 4);
                tc::tc_fence_after_sync();
#pragma unroll
                for (int c0 = 0; c0 < BN / 2; c0 += 32) {
                    uint32_t v[32];
                    tc::tmem_ld32(lane_addr + buf * BN + c0, v);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
                }
                tc::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
            }
            const int row = m0 + quarter * 32 + lane;
            if (row < M) {
                float* dst = Cout + (long long)sl * pstride + (long long)row * ldc + n0 + half * (BN / 2);
                const int ncols = min(BN / 2, N - (n0 + half * (BN / 2)));
                if (ncols == BN / 2 && (ldc & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < BN / 2; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < BN / 2; ++j)
                        if (j < ncols) dst[j] = acc[j];
                }
            }
        }
    }
    tc::tc_fence_before_sync();
    if (CLUSTER > 1)
        cluster_sync_all();
    else
        __syncthreads();
    if (warp == 2) tc::tmem_dealloc<512>(tbase);
}

}  // namespace

bool fro_gemm_supported(long long N, long long K) { return tma_encode_fn() != nullptr && (K % 4) == 0 && (N % 4) == 0 && K >= 4; }

static int fro_cluster(int M) {
    static int cluster_env = -1;
    if (cluster_env < 0) cluster_env = getenv("NMFK_FRO_CLUSTER") ? atoi(getenv("NMFK_FRO_CLUSTER")) : 2;
    return (cluster_env == 2 && (M + BM - 1) / BM >= 2) ? 2 : 1;
}

// K slices (split-K) of the stacked GEMM: the tile count alone rarely fills the 148 SMs evenly (C3: 8 x 40 tiles = 2.16 waves,
// 72 % of the last wave idle; C5's H-update: 6 x 4 tiles for 148 SMs) while K is long.  Picks the slice count with the best
// wave efficiency, charging 1.5 % per extra slice for the partial products the consumer has to add.
int fro_gemm_slices(int M, int N, int K) {
    if (getenv("NMFK_FRO_SLICES")) return std::max(1, atoi(getenv("NMFK_FRO_SLICES")));
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int cluster = fro_cluster(M);
    const int mtiles = (M + BM - 1) / BM, ntiles = (N + BN - 1) / BN, kblocks = (K + BK - 1) / BK;
    const double groups = (double)((mtiles + cluster - 1) / cluster) * ntiles, slots = (double)(sms / cluster);
    int best = 1;
    double best_score = -1.0;
    for (int S = 1; S <= 64 && kblocks / S >= 2 * CHUNK; ++S) {
        const double x = groups * S / slots;
        const double score = x / std::ceil(x) - 0.015 * (S - 1);
        if (score > best_score + 1e-9) {
            best_score = score;
            best = S;
        }
    }
    return best;
}

// C[M x N] (row-major, leading dimension ldc) = A[M x K] B[N x K]^T with the 3-term TF32 split; A / B row-major with
// leading dimensions lda / ldb (floats, multiples of 4), *lo = the x - tf32(x) images.
// S > 1: slice s of the K range writes its partial product to C + s * pstride (floats); the consumer adds them in order.
cudaError_t launch_fro_gemm(const float* Ahi, const float* Alo, long long lda, const float* Bhi, const float* Blo, long long ldb, float* C,
                            long long ldc, int M, int N, int K, int S, long long pstride, int* d_errflag, cudaStream_t s) {
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int mtiles = (M + BM - 1) / BM, ntiles = (N + BN - 1) / BN;
    const int cluster = fro_cluster(M);
    if (S < 1) S = 1;
    CUtensorMap mAhi, mAlo, mBhi, mBlo;
    if (!make_map(&mAhi, Ahi, M, K, lda, BM) || !make_map(&mAlo, Alo, M, K, lda, BM) || !make_map(&mBhi, Bhi, N, K, ldb, BN / cluster) ||
        !make_map(&mBlo, Blo, N, K, ldb, BN / cluster))
        return cudaErrorInvalidValue;
    cudaError_t e;
    const int groups = ((mtiles + cluster - 1) / cluster) * ntiles * S;
    int grid = std::min(groups * cluster, sms / cluster * cluster);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cluster == 2) {
        if ((e = cudaFuncSetAttribute(fro_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES)) != cudaSuccess) return e;
        return cudaLaunchKernelEx(&cfg, fro_gemm_kernel<2>, mAhi, mAlo, mBhi, mBlo, C, M, N, K, ldc, S, pstride, d_errflag);
    }
    if ((e = cudaFuncSetAttribute(fro_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES)) != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, fro_gemm_kernel<1>, mAhi, mAlo, mBhi, mBlo, C, M, N, K, ldc, S, pstride, d_errflag);
}

}  // namespace nmfk
