// Float32 half-update of the tiled KL engine on the 5th-generation tensor cores (tcgen05 / tensor memory).
//
// Replaces, for Float32 data without NaN, the loop body of NMFk.NMFmultiplicative
// (/root/reference/src/NMFkMultiplicative.jl:67,70) exactly like tiled_pass_kernel (kl_tiled.cuh) does:
//   ACC[o,a] = sum_t (D[o,t] / (U[o,:] . V[t,:])) V[t,a]      then  U[o,a] <- U[o,a] ACC[o,a] / den[a]
// (H-update: D = X^T, U = H^T, V = W;  W-update: D = X, U = W, V = H^T).
//
// One CTA = 128 own indices (the 128 lanes of tensor memory) x one slice of the reduction range x a GROUP of
// RB restarts that share every X tile.  The reduction range is walked in chunks of TS = 64 steps; a "unit" is
// one (chunk, restart) pair:
//   MMA#1  P[128 x 64]  = U_r[128 x k] V_r[64 x k]^T      A = U (hi, lo) in tensor memory, B = V chunk in smem
//   quot.  Q = X_tile ./ P  in registers (tcgen05.ld, MUFU.RCP, tcgen05.st), split Q = Qhi + Qlo (TF32 + rest)
//   MMA#2  ACC_r[128 x k] += Q[128 x 64] V_r[64 x k]      A = Q (hi, lo) in tensor memory, B = V chunk in smem
// Both products use the 3-term TF32 split (hi*hi + lo*hi + hi*lo, FP32 accumulation; never plain TF32).
// Warp roles (384 threads): warp 0 = bulk-copy producer of X tiles (cp.async.bulk + mbarrier transaction
// counts, 3 stages), warp 1 = tcgen05.mma issuer, warps 2-3 = V stagers (global -> hi/lo split -> the two
// canonical un-swizzled K-major images MMA#1 and MMA#2 read), warps 4-11 = quotient warps (lane quarter =
// warp % 4, column half = (warp - 4) / 4).  P/Q tiles are double-buffered in tensor memory so the tensor pipe
// works on unit u+1 / u+2 while the quotient warps divide unit u.
// tcgen05.mma accumulates with round-toward-zero (measured: -0.47 ulp per chained instruction, tools/umma_selftest.py
// --timing), so numerators are NOT chained across units: MMA#2 of every unit starts a fresh tensor-memory
// accumulator (8 chained K-steps) that the quotient warps add into FP32 registers with round-to-nearest.
// Tensor-memory map (512 columns): [0,256) two P/Q buffers (64 P->Qhi + 64 Qlo), then two per-unit numerator
// buffers (NST*N2 columns each; k <= 16 stacks Qhi*[Vhi;Vlo] into one N = 32 instruction), then U hi | U lo
// (K8 each) per restart of the group.
#include <algorithm>

#include "kl_tiled_args.h"
#include "tc_ptx.cuh"

namespace nmfk {
namespace {

constexpr int TC_M = 128, TC_TS = 64, TC_THREADS = 384;
constexpr int TC_NXS = 3, TC_NVB = 3;
constexpr uint32_t TC_LBO = 128;

template <int K8, int N2>
struct TcCfg {
    static constexpr int NST = N2 == 16 ? 2 : 1;                      // MMA#2: Qhi * [Vhi ; Vlo] stacked along N
    static constexpr int ACOLS = NST * N2;                            // columns of one per-unit numerator buffer
    static constexpr int ABASE = 256, UBASE = 256 + 2 * ACOLS;        // tensor-memory columns
    static constexpr int PERB = 2 * K8;                               // U hi | U lo per restart
    static constexpr int RBT = (512 - UBASE) / PERB;
    static constexpr int RB = RBT < 4 ? RBT : 4;                      // restarts per CTA (they share the X tiles)
    static constexpr uint32_t SBO1 = (K8 / 4) * 128;                  // MMA#1 B: rows = steps, K extent = K8
    static constexpr uint32_t SBO2 = (TC_TS / 4) * 128;               // MMA#2 B: rows = columns a, K extent = TS
    static constexpr uint32_t B1_BYTES = TC_TS * K8 * 4;
    static constexpr uint32_t B2_BYTES = N2 * TC_TS * 4;
    static constexpr uint32_t V_BYTES = 2 * B1_BYTES + 2 * B2_BYTES;  // B1 hi | B1 lo | B2 hi | B2 lo
    static constexpr uint32_t X_BYTES = TC_TS * TC_M * 4;
    static constexpr size_t SMEM = (size_t)TC_NXS * X_BYTES + (size_t)TC_NVB * V_BYTES + 32 * 8 + 64;
};

__device__ __forceinline__ float rcp_fast(float p) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
    return r;
}

template <int K8, int N2>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_pass_kernel(const TiledPassArgs a, int* errflag) {
    using C = TcCfg<K8, N2>;
    constexpr int RB = C::RB;
    constexpr int NC = N2 / 2;  // numerator columns per quotient thread (the two column-half warps split them)
    extern __shared__ __align__(1024) unsigned char smem[];
    float* Xs = reinterpret_cast<float*>(smem);                         // [NXS][TS][M]
    unsigned char* Vs = smem + (size_t)TC_NXS * C::X_BYTES;              // [NVB][V_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + (size_t)TC_NVB * C::V_BYTES);
    uint64_t* x_full = bars;                 // [NXS]  X tile landed (bulk-copy transaction count)
    uint64_t* x_empty = x_full + TC_NXS;     // [NXS]  quotient warps are done with the tile
    uint64_t* v_full = x_empty + TC_NXS;     // [NVB]  V images staged
    uint64_t* v_empty = v_full + TC_NVB;     // [NVB]  MMA#1 and MMA#2 of the unit have read them
    uint64_t* p_full = v_empty + TC_NVB;     // [2]    MMA#1 done: P readable
    uint64_t* q_full = p_full + 2;           // [2]    Q written (and the numerator buffer of unit u-2 drained)
    uint64_t* a_full = q_full + 2;           // [2]    MMA#2 done: the unit's numerators readable
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
    int* s_act = reinterpret_cast<int*>(tmem_slot + 1);  // [RB] active restarts of the group, then their count

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ngroups = (a.R + RB - 1) / RB;
    const int g = blockIdx.x % ngroups;
    const int rest = blockIdx.x / ngroups;
    const int ob = rest % a.nblocks;
    const int slice = rest / a.nblocks;
    const int o0 = ob * TC_M;
    const int k = a.k;
    const int t_begin = (int)(((long long)a.nred * slice) / a.S);
    const int t_end = (int)(((long long)a.nred * (slice + 1)) / a.S);
    const int nchunks = (t_end - t_begin + TC_TS - 1) / TC_TS;

    if (tid == 0) {
        int nact = 0;
        for (int b = 0; b < RB; ++b) {
            const int r = g * RB + b;
            if (r < a.R && a.st[r].stop == 0) s_act[nact++] = r;  // finished restarts are frozen
        }
        s_act[RB] = nact;
    }
    __syncthreads();
    const int nact = s_act[RB];
    if (nact == 0 || nchunks <= 0) return;
    const int total = nchunks * nact;

    if (warp == 0) tc::tmem_alloc<512>(tmem_slot);
    if (tid == 32) {
        for (int i = 0; i < TC_NXS; ++i) {
            tc::mbar_init(&x_full[i], 1);
            tc::mbar_init(&x_empty[i], 8);
        }
        for (int i = 0; i < TC_NVB; ++i) {
            tc::mbar_init(&v_full[i], 2);
            tc::mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&p_full[i], 1);
            tc::mbar_init(&q_full[i], 8);
            tc::mbar_init(&a_full[i], 1);
        }
        tc::mbar_fence_init();
    }
    // padding rows / columns of the V images stay zero for the whole kernel; so do the own indices past the
    // edge of X in the X stages (the bulk copies of an edge tile are shorter than 128 indices)
    for (uint32_t e = tid; e < TC_NVB * C::V_BYTES / 16; e += TC_THREADS) reinterpret_cast<uint4*>(Vs)[e] = make_uint4(0, 0, 0, 0);
    if (o0 + TC_M > a.nown)
        for (uint32_t e = tid; e < TC_NXS * C::X_BYTES / 16; e += TC_THREADS) reinterpret_cast<uint4*>(Xs)[e] = make_uint4(0, 0, 0, 0);
    tc::fence_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tbase = *tmem_slot;

    const float* D = static_cast<const float*>(a.D);
    const float* Vg = static_cast<const float*>(a.V);
    float* Ug = static_cast<float*>(a.U);

    // own factor rows -> tensor memory (A operand of MMA#1), split hi / lo; quotient warps, restart b by column half.
    // Own indices past the edge get U = 1 (finite P, their X is 0, their rows are never stored).
    if (warp >= 4) {
        const int lq = warp & 3, ch = (warp - 4) >> 2;
        const int o = o0 + lq * 32 + lane;
        const bool valid = o < a.nown;
        const uint32_t lane_base = tbase + ((uint32_t)(lq * 32) << 16);
        for (int b = ch; b < nact; b += 2) {
            const float* U = Ug + (long long)s_act[b] * a.u_rstride;
            uint32_t hi[K8], lo[K8];
#pragma unroll
            for (int c = 0; c < K8; ++c) {
                const float u = !valid ? 1.f : (c < k ? U[(long long)o * a.su_o + (long long)c * a.su_a] : 0.f);
                hi[c] = __float_as_uint(u) & 0xffffe000u;
                lo[c] = __float_as_uint(u - __uint_as_float(hi[c]));
            }
            const uint32_t col = C::UBASE + b * C::PERB;
#pragma unroll
            for (int c = 0; c < K8; c += 8) {
                tc::tmem_st8(lane_base + col + c, hi + c);
                tc::tmem_st8(lane_base + col + K8 + c, lo + c);
            }
        }
        tc::tmem_wait_st();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();

    if (warp == 0) {
        // ===== X producer: one bulk copy per step row (128 consecutive own indices = 512 bytes) =====
        if (lane == 0) {
            const uint32_t row_bytes = (uint32_t)min(TC_M, a.nown - o0) * 4u;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % TC_NXS;
                if (c >= TC_NXS) tc::mbar_wait(&x_empty[s], (uint32_t)((c / TC_NXS - 1) & 1), errflag, 10);
                const int t0 = t_begin + c * TC_TS;
                const int cnt = min(TC_TS, t_end - t0);
                tc::mbar_arrive_expect_tx(&x_full[s], (uint32_t)cnt * row_bytes);
                float* dst = Xs + (size_t)s * TC_TS * TC_M;
                const float* src = D + (long long)o0 + (long long)t0 * a.nown;
                for (int j = 0; j < cnt; ++j) tc::bulk_g2s(dst + j * TC_M, src + (long long)j * a.nown, row_bytes, &x_full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (one thread); descriptors are built once, the unrolled issue loops only add constants =====
        if (lane == 0) {
            constexpr uint32_t idP = tc::idesc_tf32(TC_M, TC_TS, 0);
            constexpr uint32_t idA1 = tc::idesc_tf32(TC_M, C::ACOLS, 0);  // Qhi * [Vhi ; Vlo] (or Qhi * Vhi when not stacked)
            constexpr uint32_t idA2 = tc::idesc_tf32(TC_M, N2, 0);
            const uint32_t vs0 = tc::smem_u32(Vs);
            const uint64_t d1 = tc::smem_desc(vs0, TC_LBO, C::SBO1);                     // MMA#1 image of buffer 0, hi
            const uint64_t d2 = tc::smem_desc(vs0 + 2 * C::B1_BYTES, TC_LBO, C::SBO2);   // MMA#2 image of buffer 0, hi
            constexpr uint64_t KSTEP = (2 * TC_LBO) >> 4;                                // one K-step = two 16-byte chunks
            auto mma1 = [&](int u, int b) {
                const int vb = u % TC_NVB;
                tc::mbar_wait(&v_full[vb], (uint32_t)((u / TC_NVB) & 1), errflag, 20);
                tc::tc_fence_after_sync();
                const uint32_t d = tbase + (uint32_t)(u & 1) * 128;
                const uint32_t uh = tbase + C::UBASE + b * C::PERB, ul = uh + K8;
                const uint64_t bh = d1 + (uint64_t)((vb * C::V_BYTES) >> 4), bl = bh + (C::B1_BYTES >> 4);
#pragma unroll
                for (int ks = 0; ks < K8 / 8; ++ks) {
                    tc::mma_tf32_ts(d, ul + ks * 8, bh + ks * KSTEP, idP, ks > 0);
                    tc::mma_tf32_ts(d, uh + ks * 8, bl + ks * KSTEP, idP, 1);
                    tc::mma_tf32_ts(d, uh + ks * 8, bh + ks * KSTEP, idP, 1);
                }
                tc::mma_commit(&p_full[u & 1]);
            };
            auto mma2 = [&](int u) {
                const int vb = u % TC_NVB;
                tc::mbar_wait(&q_full[u & 1], (uint32_t)((u >> 1) & 1), errflag, 21);
                tc::tc_fence_after_sync();
                const uint32_t d = tbase + C::ABASE + (uint32_t)(u & 1) * C::ACOLS;
                const uint32_t qh = tbase + (uint32_t)(u & 1) * 128, ql = qh + 64;
                const uint64_t bh = d2 + (uint64_t)((vb * C::V_BYTES) >> 4), bl = bh + (C::B2_BYTES >> 4);
#pragma unroll
                for (int ks = 0; ks < TC_TS / 8; ++ks) tc::mma_tf32_ts(d, qh + ks * 8, bh + ks * KSTEP, idA1, ks > 0);
                if (C::NST == 1) {
#pragma unroll
                    for (int ks = 0; ks < TC_TS / 8; ++ks) tc::mma_tf32_ts(d, qh + ks * 8, bl + ks * KSTEP, idA2, 1);
                }
#pragma unroll
                for (int ks = 0; ks < TC_TS / 8; ++ks) tc::mma_tf32_ts(d, ql + ks * 8, bh + ks * KSTEP, idA2, 1);
                tc::mma_commit(&v_empty[vb]);
                tc::mma_commit(&a_full[u & 1]);
            };
            // unit u = c * nact + b; MMA#1 runs two units ahead of MMA#2
            int b1 = 0;  // restart index of the next MMA#1
            auto next_b = [&](int b) { return b + 1 == nact ? 0 : b + 1; };
            mma1(0, b1);
            b1 = next_b(b1);
            if (total > 1) {
                mma1(1, b1);
                b1 = next_b(b1);
            }
            for (int u = 0; u < total; ++u) {
                mma2(u);
                if (u + 2 < total) {
                    mma1(u + 2, b1);
                    b1 = next_b(b1);
                }
            }
        }
        __syncwarp();
    } else if (warp < 4) {
        // ===== V stagers: 4 x 4 blocks (4 steps x 4 columns) of the unit's V chunk, software-prefetched =====
        const int sid = tid - 64;
        constexpr int NBLK = 16 * (K8 / 4);
        constexpr int NPT = (NBLK + 63) / 64;
        float cur[NPT][16], nxt[NPT][16];
        auto load = [&](int c, int b, float (&dst)[NPT][16]) {
            const float* V = Vg + (long long)s_act[b] * a.v_rstride;
            const int t0 = t_begin + c * TC_TS;
#pragma unroll
            for (int q = 0; q < NPT; ++q) {
                const int blk = sid + q * 64;
                const int tb = blk & 15, ab = blk >> 4;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int t = t0 + tb * 4 + i, col = ab * 4 + j;
                        dst[q][i * 4 + j] =
                            (blk < NBLK && t < t_end && col < k) ? __ldg(V + (long long)t * a.sv_t + (long long)col * a.sv_a) : 0.f;
                    }
            }
        };
        load(0, 0, cur);
        int cn = 0, bn = 0;  // (chunk, restart) of unit u + 1
        for (int u = 0; u < total; ++u) {
            if (++bn == nact) {
                bn = 0;
                ++cn;
            }
            if (u + 1 < total) load(cn, bn, nxt);
            const int vb = u % TC_NVB;
            if (u >= TC_NVB) tc::mbar_wait(&v_empty[vb], (uint32_t)((u / TC_NVB - 1) & 1), errflag, 30);
            unsigned char* base = Vs + (size_t)vb * C::V_BYTES;
#pragma unroll
            for (int q = 0; q < NPT; ++q) {
                const int blk = sid + q * 64;
                if (blk < NBLK) {
                    const int tb = blk & 15, ab = blk >> 4;
                    float h[16], l[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        h[e] = __uint_as_float(__float_as_uint(cur[q][e]) & 0xffffe000u);
                        l[e] = cur[q][e] - h[e];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {  // MMA#1 image: row = step, 16 bytes = 4 columns
                        const int t = tb * 4 + i;
                        const uint32_t off = (t & 7) * 16 + (t >> 3) * C::SBO1 + ab * TC_LBO;
                        *reinterpret_cast<float4*>(base + off) = make_float4(h[i * 4], h[i * 4 + 1], h[i * 4 + 2], h[i * 4 + 3]);
                        *reinterpret_cast<float4*>(base + C::B1_BYTES + off) =
                            make_float4(l[i * 4], l[i * 4 + 1], l[i * 4 + 2], l[i * 4 + 3]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {  // MMA#2 image: row = column a, 16 bytes = 4 steps
                        const int col = ab * 4 + j;
                        const uint32_t off = 2 * C::B1_BYTES + (col & 7) * 16 + (col >> 3) * C::SBO2 + tb * TC_LBO;
                        *reinterpret_cast<float4*>(base + off) = make_float4(h[j], h[4 + j], h[8 + j], h[12 + j]);
                        *reinterpret_cast<float4*>(base + C::B2_BYTES + off) = make_float4(l[j], l[4 + j], l[8 + j], l[12 + j]);
                    }
                }
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&v_full[vb]);
#pragma unroll
            for (int q = 0; q < NPT; ++q)
#pragma unroll
                for (int e = 0; e < 16; ++e) cur[q][e] = nxt[q][e];
        }
        __syncwarp();
    } else {
        // ===== quotient warps =====
        const int lq = warp & 3, ch = (warp - 4) >> 2;
        const int o_loc = lq * 32 + lane;
        const int o = o0 + o_loc;
        const bool valid = o < a.nown;
        const uint32_t lane_base = tbase + ((uint32_t)(lq * 32) << 16);
        const int j0 = ch * 32;
        float acc[RB][NC];
#pragma unroll
        for (int b = 0; b < RB; ++b)
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[b][c] = 0.f;
        // numerators of unit uu (restart slot bb, compile-time) -> registers, round-to-nearest adds
        auto drain = [&](int uu, int target) {
            tc::mbar_wait(&a_full[uu & 1], (uint32_t)((uu >> 1) & 1), errflag, 43);
            tc::tc_fence_after_sync();
            const uint32_t col = lane_base + C::ABASE + (uint32_t)(uu & 1) * C::ACOLS + ch * NC;
            uint32_t v0[NC], v1[NC];
#pragma unroll
            for (int c = 0; c < NC; c += 8) {
                tc::tmem_ld8(col + c, v0 + c);
                if (C::NST == 2) tc::tmem_ld8(col + N2 + c, v1 + c);
            }
            tc::tmem_wait_ld();
#pragma unroll
            for (int bb = 0; bb < RB; ++bb)
                if (bb == target) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        float v = __uint_as_float(v0[c]);
                        if (C::NST == 2) v += __uint_as_float(v1[c]);
                        acc[bb][c] += v;
                    }
                }
        };
        int u = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % TC_NXS;
            const int cnt = min(TC_TS, t_end - (t_begin + c * TC_TS));
            tc::mbar_wait(&x_full[s], (uint32_t)((c / TC_NXS) & 1), errflag, 40);
            const float* xs = Xs + (size_t)s * TC_TS * TC_M + (size_t)j0 * TC_M + o_loc;
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                if (b < nact) {
                    const uint32_t col = lane_base + (uint32_t)(u & 1) * 128 + j0;
                    tc::mbar_wait(&p_full[u & 1], (uint32_t)((u >> 1) & 1), errflag, 41);
                    tc::tc_fence_after_sync();
                    uint32_t p[32];
                    tc::tmem_ld32(col, p);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t lo[16];
                        if (cnt == TC_TS) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int jj = half * 16 + j;
                                const float q = xs[jj * TC_M] * rcp_fast(__uint_as_float(p[jj]));
                                const uint32_t h = __float_as_uint(q) & 0xffffe000u;
                                lo[j] = __float_as_uint(q - __uint_as_float(h));
                                p[jj] = h;
                            }
                        } else {  // last chunk of the slice: steps past the end contribute nothing
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int jj = half * 16 + j;
                                float q = xs[jj * TC_M] * rcp_fast(__uint_as_float(p[jj]));
                                q = (j0 + jj < cnt) ? q : 0.f;
                                const uint32_t h = __float_as_uint(q) & 0xffffe000u;
                                lo[j] = __float_as_uint(q - __uint_as_float(h));
                                p[jj] = h;
                            }
                        }
                        tc::tmem_st16(col + 64 + half * 16, lo);
                    }
                    tc::tmem_st32(col, p);
                    tc::tmem_wait_st();
                    tc::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&q_full[u & 1]);
                    // drain the numerators of the previous unit while the tensor pipe works on this one
                    if (b > 0)
                        drain(u - 1, b - 1);
                    else if (u > 0)
                        drain(u - 1, nact - 1);
                    ++u;
                }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&x_empty[s]);
        }
        drain(total - 1, nact - 1);
        // numerators -> factor update (or the slice's partial sums): this thread's NC columns of every restart
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            if (b < nact && valid) {
                const int r = s_act[b];
                if (a.partial == nullptr) {
                    float* U = Ug + (long long)r * a.u_rstride;
                    const float* den = static_cast<const float*>(a.den) + (long long)r * 32;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const int col = ch * NC + c;
                        if (col < k) {
                            const long long idx = (long long)o * a.su_o + (long long)col * a.su_a;
                            U[idx] = (U[idx] * acc[b][c]) / den[col];
                        }
                    }
                } else {
                    float* dst = static_cast<float*>(a.partial) + (((long long)slice * a.R + r) * a.nown + o) * a.ktmpl;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const int col = ch * NC + c;
                        if (col < a.ktmpl) dst[col] = acc[b][c];
                    }
                }
            }
        }
        __syncwarp();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tbase);
}

template <int K8, int N2>
cudaError_t launch_tc(const TiledPassArgs& a, int* d_errflag, cudaStream_t s) {
    using C = TcCfg<K8, N2>;
    const int ngroups = (a.R + C::RB - 1) / C::RB;
    const long long grid = (long long)a.S * a.nblocks * ngroups;
    if (grid > 2147483647ll) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(tc_pass_kernel<K8, N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
    tc_pass_kernel<K8, N2><<<(unsigned)grid, TC_THREADS, C::SMEM, s>>>(a, d_errflag);
    return cudaGetLastError();
}

}  // namespace

int tc_pass_group(int k) {
    if (k <= 8) return TcCfg<8, 16>::RB;
    if (k <= 16) return TcCfg<16, 16>::RB;
    if (k <= 24) return TcCfg<24, 32>::RB;
    return TcCfg<32, 32>::RB;
}

// bulk copies need 16-byte aligned rows of 128 own indices: nown % 4 == 0
bool tc_pass_supported(const TiledPassArgs& a) {
    return !a.has_nan && a.k >= 1 && a.k <= 32 && (a.nown % 4) == 0 && (reinterpret_cast<uintptr_t>(a.D) % 16) == 0;
}

cudaError_t launch_tc_pass(const TiledPassArgs& a, int* d_errflag, cudaStream_t s) {
    if (a.k <= 8) return launch_tc<8, 16>(a, d_errflag, s);
    if (a.k <= 16) return launch_tc<16, 16>(a, d_errflag, s);
    if (a.k <= 24) return launch_tc<24, 32>(a, d_errflag, s);
    return launch_tc<32, 32>(a, d_errflag, s);
}

}  // namespace nmfk
