// Device self-test of the tcgen05 building blocks of the Float32 tiled engine (tc_ptx.cuh): one CTA,
// M = 128 own indices, TS = 64 steps, KP = 16 columns.
//   bit 0: P_ss  = U V^T            A and B from shared memory (K-major, un-swizzled)
//   bit 1: P_ts  = U V^T            A from tensor memory
//   bit 2: ACC_a = (Qhi + Qlo) V    A from tensor memory, B = second K-major copy of V
//   bit 3: ACC_b = (Qhi + Qlo) V    A from tensor memory, B = the FIRST copy of V read as an MN-major operand
// with Q = 0.5 * (U V^T) computed by scalar FMAs and split as hi = q & 0xffffe000, lo = q - hi.
// tests/test_gpu_parity.py compares the four outputs with NumPy.
#include "nmfk_internal.h"
#include <cstdio>

#include "tc_ptx.cuh"

namespace nmfk {
namespace {

constexpr int M = 128, TS = 64, KP = 16;
constexpr uint32_t LBO = 128;
constexpr uint32_t SBO_K16 = (KP / 4) * 128;  // row groups of an operand whose K extent is KP
constexpr uint32_t SBO_K64 = (TS / 4) * 128;  // ... whose K extent is TS
// tensor-memory columns
constexpr uint32_t C_PSS = 0, C_PTS = 64, C_ACCA = 128, C_ACCB = 144, C_U = 160, C_QH = 176, C_QL = 240;

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* __restrict__ U, const float* __restrict__ V, int mode,
                                                            float* __restrict__ Pss, float* __restrict__ Pts,
                                                            float* __restrict__ ACCa, float* __restrict__ ACCb, int* errflag) {
    __shared__ __align__(128) float As[M * KP];
    __shared__ __align__(128) float B1[TS * KP];
    __shared__ __align__(128) float B2[KP * TS];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        tc::mbar_init(&bar[0], 1);
        tc::mbar_init(&bar[1], 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);

    // operands -> shared memory in the canonical layouts
    for (int e = tid; e < M * KP; e += 128) {
        const int o = e / KP, a = e % KP;
        As[((o % 8) * 16 + (o / 8) * SBO_K16 + (a / 4) * LBO + (a % 4) * 4) / 4] = U[o * KP + a];
    }
    for (int e = tid; e < TS * KP; e += 128) {
        const int t = e / KP, a = e % KP;
        const float v = V[t * KP + a];
        B1[((t % 8) * 16 + (t / 8) * SBO_K16 + (a / 4) * LBO + (a % 4) * 4) / 4] = v;
        B2[((a % 8) * 16 + (a / 8) * SBO_K64 + (t / 4) * LBO + (t % 4) * 4) / 4] = v;
    }
    // own row of U -> tensor memory (A operand of P_ts); Q = 0.5 U V^T split in hi / lo -> tensor memory
    {
        uint32_t r[16];
        float u[KP];
#pragma unroll
        for (int a = 0; a < KP; ++a) {
            u[a] = U[tid * KP + a];
            r[a] = __float_as_uint(u[a]);
        }
        tc::tmem_st16(lane_base + C_U, r);
        for (int t0 = 0; t0 < TS; t0 += 16) {
            uint32_t qh[16], ql[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float p = 0.f;
#pragma unroll
                for (int a = 0; a < KP; ++a) p = fmaf(u[a], V[(t0 + j) * KP + a], p);
                const float q = 0.5f * p;
                qh[j] = __float_as_uint(q) & 0xffffe000u;
                ql[j] = __float_as_uint(q - __uint_as_float(qh[j]));
            }
            tc::tmem_st16(lane_base + C_QH + t0, qh);
            tc::tmem_st16(lane_base + C_QL + t0, ql);
        }
        tc::tmem_wait_st();
    }
    tc::fence_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();

    if (tid == 0) {
        tc::tc_fence_after_sync();
        const uint32_t idP = tc::idesc_tf32(M, TS, 0);
        if (mode & 1)
            for (int ks = 0; ks < KP / 8; ++ks)
                tc::mma_tf32_ss(tbase + C_PSS, tc::smem_desc(tc::smem_u32(As) + ks * 2 * LBO, LBO, SBO_K16),
                                tc::smem_desc(tc::smem_u32(B1) + ks * 2 * LBO, LBO, SBO_K16), idP, ks > 0);
        if (mode & 2)
            for (int ks = 0; ks < KP / 8; ++ks)
                tc::mma_tf32_ts(tbase + C_PTS, tbase + C_U + ks * 8, tc::smem_desc(tc::smem_u32(B1) + ks * 2 * LBO, LBO, SBO_K16),
                                idP, ks > 0);
        if (mode & 4) {
            const uint32_t id = tc::idesc_tf32(M, KP, 0);
            for (int half = 0; half < 2; ++half)
                for (int ks = 0; ks < TS / 8; ++ks)
                    tc::mma_tf32_ts(tbase + C_ACCA, tbase + (half ? C_QL : C_QH) + ks * 8,
                                    tc::smem_desc(tc::smem_u32(B2) + ks * 2 * LBO, LBO, SBO_K64), id, (half | ks) > 0);
        }
        if (mode & 8) {
            const uint32_t id = tc::idesc_tf32(M, KP, 1);
            for (int half = 0; half < 2; ++half)
                for (int ks = 0; ks < TS / 8; ++ks)  // MN-major: 4-column groups LBO(=128) apart, 8-step groups SBO_K16 apart
                    tc::mma_tf32_ts(tbase + C_ACCB, tbase + (half ? C_QL : C_QH) + ks * 8,
                                    (mode & 16) ? tc::smem_desc(tc::smem_u32(B1) + ks * SBO_K16, LBO, SBO_K16)
                                                : tc::smem_desc(tc::smem_u32(B1) + ks * SBO_K16, SBO_K16, LBO),
                                    id, (half | ks) > 0);
        }
        tc::mma_commit(&bar[0]);
    }
    tc::mbar_wait(&bar[0], 0, errflag, 1);
    tc::tc_fence_after_sync();

    uint32_t r32[32];
    uint32_t r16[16];
    if (mode & 1)
        for (int c = 0; c < TS; c += 32) {
            tc::tmem_ld32(lane_base + C_PSS + c, r32);
            tc::tmem_wait_ld();
            for (int j = 0; j < 32; ++j) Pss[tid * TS + c + j] = __uint_as_float(r32[j]);
        }
    if (mode & 2)
        for (int c = 0; c < TS; c += 32) {
            tc::tmem_ld32(lane_base + C_PTS + c, r32);
            tc::tmem_wait_ld();
            for (int j = 0; j < 32; ++j) Pts[tid * TS + c + j] = __uint_as_float(r32[j]);
        }
    if (mode & 4) {
        tc::tmem_ld16(lane_base + C_ACCA, r16);
        tc::tmem_wait_ld();
        for (int j = 0; j < 16; ++j) ACCa[tid * KP + j] = __uint_as_float(r16[j]);
    }
    if (mode & 8) {
        tc::tmem_ld16(lane_base + C_ACCB, r16);
        tc::tmem_wait_ld();
        for (int j = 0; j < 16; ++j) ACCb[tid * KP + j] = __uint_as_float(r16[j]);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tbase);
}


// Cycle counts of the building blocks (one CTA, thread 0 issues):  out[0] one N=64 MMA + commit + wait,
// out[1] 96 back-to-back N=64 K=8 MMAs (A in tensor memory) + commit + wait, out[2] 96 N=16 MMAs,
// out[3] 96 N=32 MMAs, out[4] 16 x (tcgen05.ld x32 + wait), out[5] 16 x (tcgen05.st x32 + wait),
// out[6] 96 N=64 MMAs with A from shared memory, out[7] 96 N=16 MMAs alternating between two accumulators.
// bias[0..127]: ACC[lane][0] after `reps` x 8 accumulating K-steps of the same (Qhi,V) product (RZ accumulation test).
__global__ void __launch_bounds__(128) umma_timing_kernel(const float* __restrict__ U, const float* __restrict__ V, int reps,
                                                          long long* __restrict__ out, float* __restrict__ bias, int* errflag) {
    __shared__ __align__(128) float As[M * KP];
    __shared__ __align__(128) float B1[TS * KP];
    __shared__ __align__(128) float B2[2 * KP * TS];  // 32 rows: the N = 32 timing reads rows 16..31 (zeros)
    __shared__ __align__(8) uint64_t bar[1];
    __shared__ __align__(8) uint64_t scratch_bar[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        tc::mbar_init(&scratch_bar[0], 1);
        tc::mbar_init(&scratch_bar[1], 1);
        tc::mbar_init(&bar[0], 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    for (int e = tid; e < 2 * KP * TS; e += 128) B2[e] = 0.f;
    __syncthreads();
    for (int e = tid; e < M * KP; e += 128) {
        const int o = e / KP, a = e % KP;
        As[((o % 8) * 16 + (o / 8) * SBO_K16 + (a / 4) * LBO + (a % 4) * 4) / 4] = U[o * KP + a];
    }
    for (int e = tid; e < TS * KP; e += 128) {
        const int t = e / KP, a = e % KP;
        const float v = V[t * KP + a];
        B1[((t % 8) * 16 + (t / 8) * SBO_K16 + (a / 4) * LBO + (a % 4) * 4) / 4] = v;
        B2[((a % 8) * 16 + (a / 8) * SBO_K64 + (t / 4) * LBO + (t % 4) * 4) / 4] = v;
    }
    {
        uint32_t r[16];
#pragma unroll
        for (int a = 0; a < KP; ++a) r[a] = __float_as_uint(U[tid * KP + a]);
        tc::tmem_st16(lane_base + C_U, r);
        for (int t0 = 0; t0 < TS; t0 += 16) {  // "Q" = any full-precision values: U row repeated
            tc::tmem_st16(lane_base + C_QH + t0, r);
        }
        tc::tmem_wait_st();
    }
    tc::fence_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    uint32_t phase = 0;
    // issued from an elect.sync region of warp 0 (no waterfall loops), unrolled by 8
    auto timed = [&](int which, int nmma, int N, bool ss, bool alt) {
        long long t0 = 0;
        if (warp == 0) {
            tc::tc_fence_after_sync();
            const uint32_t id = tc::idesc_tf32(M, N, 0);
            const uint64_t da = tc::smem_desc(tc::smem_u32(As), LBO, SBO_K16), db1 = tc::smem_desc(tc::smem_u32(B1), LBO, SBO_K16);
            const uint64_t db2 = tc::smem_desc(tc::smem_u32(B2), LBO, SBO_K64);
            t0 = clock64();
            if (tc::elect_one()) {
                for (int i0 = 0; i0 < nmma; i0 += 8) {
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii) {
                        if (i0 + ii < nmma) {
                            const uint32_t d = tbase + ((alt && (ii & 1)) ? 64 : 0);
                            if (ss)
                                tc::mma_tf32_ss(d, da, db1, id, 1);
                            else if (N == 64)
                                tc::mma_tf32_ts(d, tbase + C_U, db1, id, 1);
                            else
                                tc::mma_tf32_ts(d, tbase + C_QH + ii * 8, db2 + ii * 16, id, 1);
                        }
                    }
                }
                tc::mma_commit(&bar[0]);
            }
            __syncwarp();
        }
        tc::mbar_wait(&bar[0], phase, errflag, 2);
        phase ^= 1;
        tc::tc_fence_after_sync();
        if (tid == 0) out[which] = clock64() - t0;
        tc::tc_fence_before_sync();
        __syncthreads();
    };
    timed(0, 1, 64, false, false);
    timed(0, 1, 64, false, false);
    timed(1, 96, 64, false, false);
    timed(2, 96, 16, false, false);
    timed(3, 96, 32, false, false);
    timed(6, 96, 64, true, false);
    timed(7, 96, 16, false, true);
    // Burst experiments (reps has bit 20 set): how long the issuing thread is held by the short MMA bursts of the KL pass -
    // mode 0: 16 small MMAs (8 x N=32 + 8 x N=16, chained, A from tensor memory) + 2 commits; 1: + 1 commit; 2: no commit;
    // 3: 6 x N=64 (A from tensor memory) + commit; 4: alternating 3 and 0; 5: as 4 with two of the six N=64 MMAs reading A from
    // shared memory.  Prints the cycles per burst spent issuing and the cycles per burst until completion.
    if (reps & (1 << 20)) {
        for (int mode = 0; mode < 6; ++mode) {
            long long t0 = 0, t1 = 0;
            if (warp == 0) {
                tc::tc_fence_after_sync();
                const uint32_t id64 = tc::idesc_tf32(M, 64, 0), id32 = tc::idesc_tf32(M, 32, 0), id16 = tc::idesc_tf32(M, 16, 0);
                const uint64_t da = tc::smem_desc(tc::smem_u32(As), LBO, SBO_K16), db1 = tc::smem_desc(tc::smem_u32(B1), LBO, SBO_K16);
                const uint64_t db2 = tc::smem_desc(tc::smem_u32(B2), LBO, SBO_K64);
                t0 = clock64();
                if (tc::elect_one()) {
                    for (int burst = 0; burst < 16; ++burst) {
                        const bool big = mode == 3 || (mode >= 4 && (burst & 1) == 0);
                        if (big) {
                            const uint32_t d = tbase + 320 + (burst & 2) * 32;
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                if (mode == 5)
                                    tc::mma_tf32_ss(d, da + ks * 16, db1 + ks * 16, id64, ks > 0);
                                else
                                    tc::mma_tf32_ts(d, tbase + C_U + ks * 8, db1 + ks * 16, id64, ks > 0);
                                tc::mma_tf32_ts(d, tbase + C_U + ks * 8, db1 + ks * 16, id64, 1);
                                tc::mma_tf32_ts(d, tbase + C_U + ks * 8, db1 + ks * 16, id64, 1);
                            }
                            tc::mma_commit(&scratch_bar[0]);
                        } else {
                            const uint32_t d = tbase + 448 + (burst & 2) * 16;
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks) tc::mma_tf32_ts(d, tbase + C_QH + ks * 8, db2 + ks * 16, id32, ks > 0);
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks) tc::mma_tf32_ts(d, tbase + C_QL + ks * 8, db2 + ks * 16, id16, 1);
                            if (mode != 2) tc::mma_commit(&scratch_bar[0]);
                            if (mode == 0 || mode >= 4) tc::mma_commit(&scratch_bar[1]);
                        }
                    }
                    t1 = clock64();
                    tc::mma_commit(&bar[0]);
                }
                __syncwarp();
            }
            tc::mbar_wait(&bar[0], phase, errflag, 4);
            phase ^= 1;
            tc::tc_fence_after_sync();
            const long long t2 = clock64();
            if (warp == 0) {
                const long long ti = __shfl_sync(0xffffffffu, t1, 0) | 0;  // only the elected lane holds t1
                long long tmax = t1;
                for (int off = 16; off > 0; off >>= 1) tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
                (void)ti;
                if (tid == 0) printf("burst mode %d: issue %lld clk / burst, complete %lld clk / burst\n", mode, (tmax - t0) / 16, (t2 - t0) / 16);
            }
            tc::tc_fence_before_sync();
            __syncthreads();
        }
    }
    // Tensor-memory load / store throughput with all four warps (one per 32-lane quarter) streaming, alone and while the
    // tensor pipe runs a long series of N = 64 MMAs (reps bit 20): clk per 4 x 32 x 32 x 4 B = 16 KB.
    if (reps & (1 << 20)) {
        for (int mode = 0; mode < 4; ++mode) {  // 0: loads, 1: stores, 2: loads under MMA, 3: stores under MMA
            __syncthreads();
            if ((mode & 2) && warp == 0) {
                tc::tc_fence_after_sync();
                const uint32_t id64 = tc::idesc_tf32(M, 64, 0);
                const uint64_t db1 = tc::smem_desc(tc::smem_u32(B1), LBO, SBO_K16);
                if (tc::elect_one()) {
                    for (int i = 0; i < 64; ++i) tc::mma_tf32_ts(tbase + 320, tbase + C_U, db1, id64, 1);
                    tc::mma_commit(&bar[0]);
                }
                __syncwarp();
            }
            uint32_t r32[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) r32[j] = tid + j;
            uint32_t sink = 0;
            const long long t0 = clock64();
            for (int i = 0; i < 32; ++i) {
                if (mode & 1) {
                    tc::tmem_st32(lane_base + 384 + (i & 1) * 32, r32);
                } else {
                    tc::tmem_ld32(lane_base + 384 + (i & 1) * 32, r32);
                    if ((i & 7) == 7) {
                        tc::tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) sink ^= r32[j];
                    }
                }
            }
            if (mode & 1) tc::tmem_wait_st(); else tc::tmem_wait_ld();
            const long long t1 = clock64();
            if (sink == 0x9e3779b9u) out[0] = 1;
            if (mode & 2) {
                tc::mbar_wait(&bar[0], phase, errflag, 5);
                phase ^= 1;
                tc::tc_fence_after_sync();
            }
            const long long t2 = clock64();
            if (tid == 0) printf("tmem mode %d: %lld clk per 16 KB (4 warps x x32), mma series done after %lld clk (2048+ alone)\n", mode, (t1 - t0) / 32, t2 - t0);
            tc::tc_fence_before_sync();
            __syncthreads();
        }
    }
    {
        uint32_t r32[32];
        long long t0 = clock64();
        for (int i = 0; i < 16; ++i) {
            tc::tmem_ld32(lane_base + C_QH + (i & 1) * 32, r32);
            tc::tmem_wait_ld();
        }
        long long t1 = clock64();
        for (int i = 0; i < 16; ++i) {
            tc::tmem_st32(lane_base + C_QL + (i & 1) * 32, r32);
            tc::tmem_wait_st();
        }
        long long t2 = clock64();
        if (tid == 0) {
            out[4] = t1 - t0;
            out[5] = t2 - t1;
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    // RZ test: reps x 8 K-steps of Qhi(:, 0..63) * V accumulated in one tensor-memory accumulator (N = 16)
    if (tid == 0) {
        tc::tc_fence_after_sync();
        const uint32_t id = tc::idesc_tf32(M, KP, 0);
        for (int i = 0; i < (reps & 0xfffff) * 8; ++i)
            tc::mma_tf32_ts(tbase + C_ACCA, tbase + C_QH + (i & 7) * 8, tc::smem_desc(tc::smem_u32(B2) + (i & 7) * 2 * LBO, LBO, SBO_K64), id,
                            i > 0);
        tc::mma_commit(&bar[0]);
    }
    tc::mbar_wait(&bar[0], phase, errflag, 3);
    tc::tc_fence_after_sync();
    {
        uint32_t r16[16];
        tc::tmem_ld16(lane_base + C_ACCA, r16);
        tc::tmem_wait_ld();
        for (int j = 0; j < 16; ++j) bias[tid * KP + j] = __uint_as_float(r16[j]);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tbase);
}

// Dense tcgen05.mma kind::tf32 throughput (the denominator of the Float32 tensor-core kernels): every SM runs one CTA whose
// elected thread issues `iters` back-to-back M = 128, N = 256, K = 8 instructions with both operands in shared memory
// (un-swizzled K-major tiles: 4 KB + 8 KB read per instruction = 96 B/clk, below the shared-memory bandwidth) into two
// alternating tensor-memory accumulators, then one commit.
__global__ void __launch_bounds__(128) umma_peak_kernel(int iters, int* errflag) {
    __shared__ __align__(128) float As[128 * 8];
    __shared__ __align__(128) float Bs[256 * 8];
    __shared__ __align__(8) uint64_t bar[1];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 128 * 8; e += 128) As[e] = 0.f;
    for (int e = tid; e < 256 * 8; e += 128) Bs[e] = 0.f;
    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        tc::mbar_init(&bar[0], 1);
        tc::mbar_fence_init();
    }
    tc::fence_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if (warp == 0) {
        const uint32_t id = tc::idesc_tf32(128, 256, 0);
        const uint64_t da = tc::smem_desc(tc::smem_u32(As), 128, 256), db = tc::smem_desc(tc::smem_u32(Bs), 128, 256);
        if (tc::elect_one()) {
            for (int i0 = 0; i0 < iters; i0 += 8) {
#pragma unroll
                for (int ii = 0; ii < 8; ++ii) tc::mma_tf32_ss(tbase + ((ii & 1) ? 256 : 0), da, db, id, (i0 | (ii >> 1)) != 0);
            }
            tc::mma_commit(&bar[0]);
        }
        __syncwarp();
    }
    tc::mbar_wait(&bar[0], 0, errflag, 9);
    tc::tc_fence_after_sync();
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tbase);
}

}  // namespace

// TFLOP/s of dense tcgen05.mma kind::tf32 (2 * 128 * 256 * 8 flops per instruction) over all SMs, best of 5 launches
cudaError_t umma_peak(double* tflops, cudaStream_t s) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int* derr = nullptr;
    cudaError_t e = cudaMalloc(&derr, sizeof(int));
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(derr, 0, sizeof(int), s);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 1 << 15;
    float best = 1e30f;
    for (int rep = 0; rep < 7 && e == cudaSuccess; ++rep) {
        cudaEventRecord(e0, s);
        umma_peak_kernel<<<sms, 128, 0, s>>>(iters, derr);
        cudaEventRecord(e1, s);
        e = cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(derr);
    if (e != cudaSuccess) return e;
    *tflops = 2.0 * 128.0 * 256.0 * 8.0 * (double)iters * sms / (best * 1e-3) / 1e12;
    return cudaSuccess;
}

// U: 128 x 16 row-major, V: 64 x 16 row-major (host); outputs row-major 128 x 64 / 128 x 16 (host)
cudaError_t umma_selftest(const float* U, const float* V, int mode, float* Pss, float* Pts, float* ACCa, float* ACCb, int* err,
                          cudaStream_t s) {
    float* d = nullptr;
    int* derr = nullptr;
    const size_t nU = M * KP, nV = TS * KP, nP = M * TS, nA = M * KP;
    cudaError_t e;
    if ((e = cudaMalloc(&d, (nU + nV + 2 * nP + 2 * nA) * sizeof(float))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&derr, sizeof(int))) != cudaSuccess) {
        cudaFree(d);
        return e;
    }
    float *dU = d, *dV = dU + nU, *dPss = dV + nV, *dPts = dPss + nP, *dAa = dPts + nP, *dAb = dAa + nA;
    cudaMemsetAsync(d, 0, (nU + nV + 2 * nP + 2 * nA) * sizeof(float), s);
    cudaMemsetAsync(derr, 0, sizeof(int), s);
    cudaMemcpyAsync(dU, U, nU * sizeof(float), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(dV, V, nV * sizeof(float), cudaMemcpyHostToDevice, s);
    umma_selftest_kernel<<<1, 128, 0, s>>>(dU, dV, mode, dPss, dPts, dAa, dAb, derr);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) {
        cudaMemcpy(Pss, dPss, nP * sizeof(float), cudaMemcpyDeviceToHost);
        cudaMemcpy(Pts, dPts, nP * sizeof(float), cudaMemcpyDeviceToHost);
        cudaMemcpy(ACCa, dAa, nA * sizeof(float), cudaMemcpyDeviceToHost);
        cudaMemcpy(ACCb, dAb, nA * sizeof(float), cudaMemcpyDeviceToHost);
        cudaMemcpy(err, derr, sizeof(int), cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    cudaFree(derr);
    return e;
}

cudaError_t umma_timing(const float* U, const float* V, int reps, long long* out8, float* bias, cudaStream_t s) {
    float* d = nullptr;
    long long* dout = nullptr;
    const size_t nU = M * KP, nV = TS * KP, nA = M * KP;
    cudaError_t e;
    if ((e = cudaMalloc(&d, (nU + nV + nA) * sizeof(float))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&dout, 8 * sizeof(long long) + sizeof(int))) != cudaSuccess) {
        cudaFree(d);
        return e;
    }
    cudaMemsetAsync(dout, 0, 8 * sizeof(long long) + sizeof(int), s);
    cudaMemcpyAsync(d, U, nU * sizeof(float), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d + nU, V, nV * sizeof(float), cudaMemcpyHostToDevice, s);
    umma_timing_kernel<<<1, 128, 0, s>>>(d, d + nU, reps, dout, d + nU + nV, reinterpret_cast<int*>(dout + 8));
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) {
        cudaMemcpy(out8, dout, 8 * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaMemcpy(bias, d + nU + nV, nA * sizeof(float), cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    cudaFree(dout);
    return e;
}

}  // namespace nmfk
