// Resident KL engine, Float64, tensor-pipe formulation (DMMA: mma.sync.aligned.m8n8k4.f64).
//
// Same contract as kl_resident.cuh (one CTA = one restart of NMFmultiplicative,
// /root/reference/src/NMFkMultiplicative.jl:56-127 + NMFkExecute.jl:791-804), different hot loop.
// The ncu capture of the scalar kernel (profiles/r01_resident_scalar_k10.txt) shows the
// shared-memory LSU data pipe at 66 % and the FP64 pipe at 31 %: every element-step needs a
// broadcast read of a k-vector.  Here a WARP owns a group of 8 "own" indices and walks the
// reduction index in tiles of 8 steps; both thin products of the KL update are m8n8k4 DMMAs:
//
//   P[8 own x 8 steps]  = U[8 x k] * V[steps, :]^T           (k/4 DMMAs, A = U fragment kept in registers)
//   Q = X ./ P                                               (2 elements per lane, C-fragment layout)
//   ACC[8 own x k]     += Q[8 x 8 steps] * V[steps, :]       (2 * k/8 DMMAs)
//
// The C-fragment column a lane holds after the first product is exactly the A-fragment element
// it must supply to the second one if the second product's k-index q is bound to step 2q+s
// (s = 0,1), so Q never moves between lanes.  Per 64 element-steps a warp issues k/4 + k/4
// LDS.64 instead of 2*k/2 LDS.128 per 32, and holds ~k/2 doubles of state instead of 2k.
//
// Shared-memory layout: W rows and H^T rows with pitch = 4*odd doubles (zero padded): the
// B-fragment loads of both products then hit all 32 banks in the minimum 2 wavefronts.
#pragma once
#include "kl_resident.cuh"

namespace nmfk {

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

__host__ __device__ constexpr int dmma_max_threads(int KC) { return KC <= 3 ? 800 : 512; }
__host__ __device__ constexpr int dmma_pitch(int KC) { return (KC & 1) ? KC * 4 : KC * 4 + 4; }

struct DmmaSmem {
    size_t off_W, off_H, off_den, off_scr, off_red, off_idx, off_first, total;
    __host__ __device__ static DmmaSmem make(int n, int m, int KC, int SH, int SW) {
        DmmaSmem r;
        auto al = [](size_t x) { return (x + 15) & ~size_t(15); };
        const int pitch = dmma_pitch(KC), NA = (KC + 1) / 2;
        size_t o = 0;
        r.off_W = o;
        o = al(o + (size_t)n * pitch * 8);
        r.off_H = o;
        o = al(o + (size_t)m * pitch * 8);
        r.off_den = o;
        o = al(o + 40 * 8);
        r.off_scr = o;
        size_t scr = 0;
        if (SH > 1) scr = (size_t)SH * ((m + 7) / 8) * NA * 64;
        if (SW > 1 && (size_t)SW * ((n + 7) / 8) * NA * 64 > scr) scr = (size_t)SW * ((n + 7) / 8) * NA * 64;
        o = al(o + scr * 8);
        r.off_red = o;
        o = al(o + 40 * 8);
        r.off_idx = o;
        o = al(o + (size_t)m * 4);
        r.off_first = o;
        o = al(o + 40 * 4);
        r.total = o;
        return r;
    }
};

// den[a] = sum_t V[t][a] for a < k (runtime pitch version of factor_sums)
__device__ __forceinline__ void factor_sums_p(const double* __restrict__ V, int nred, int pitch, int k, double* den) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int a = w; a < k; a += nw) {
        double s = 0.0;
        for (int t = lane; t < nred; t += 32) s += V[(size_t)t * pitch + a];
        s = warp_sum(s);
        if (lane == 0) den[a] = s;
    }
}

// 1/p by MUFU.RCP64H + two Newton steps (relative error ~2^-52), no range check: operands outside
// the valid range (0, subnormal, Inf, NaN) turn the quotient into NaN, which the caller detects.
__device__ __forceinline__ double rcp_nr(double p) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    double e = fma(-p, r, 1.0);
    r = fma(r, e, r);
    e = fma(-p, r, 1.0);
    return fma(r, e, r);
}

// One 8-own x 8-step tile: P = U V^T (KC DMMAs), Q = X ./ P, ACC += Q V (2*NA DMMAs).
// vp1 -> V[t0+g][q], vp2 -> V[t0+2q][g]; x = this lane's X values at steps t0+2q, t0+2q+1.
template <int KC>
__device__ __forceinline__ void dmma_tile(const double (&ua)[KC], double x0, double x1, const double* __restrict__ vp1,
                                          const double* __restrict__ vp2a, const double* __restrict__ vp2b,
                                          bool last_ok, double (&acc)[(KC + 1) / 2][2]) {
    constexpr int NA = (KC + 1) / 2;
    constexpr int pitch = dmma_pitch(KC);
    double p[2] = {0.0, 0.0};
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) dmma884(p, ua[kc], vp1[kc * 4]);
    double q0 = x0 * rcp_nr(p[0]);
    double q1 = x1 * rcp_nr(p[1]);
    const double chk = q0 + q1;
    if (__any_sync(0xffffffffu, chk != chk)) {  // rare: redo with IEEE semantics (x/0 = Inf, 0/0 = NaN, ...)
        q0 = div_cold<double>(x0, p[0]);
        q1 = div_cold<double>(x1, p[1]);
    }
#pragma unroll
    for (int na = 0; na < NA; ++na) {
        const bool ok = (na * 8 + 8 <= pitch) || last_ok;
        const double b0 = ok ? vp2a[na * 8] : 0.0;
        const double b1 = ok ? vp2b[na * 8] : 0.0;
        dmma884(acc[na], q0, b0);
        dmma884(acc[na], q1, b1);
    }
}

// One half-update with DMMA tiles.  D: STEP-contiguous data (element (o,t) at D[t + o*nred]).
template <int KC, bool TRANSPOSED, bool HASNAN>
__device__ __forceinline__ void dmma_half_update(const double* __restrict__ D, int nown, int nred, int k, int S,
                                                 double* __restrict__ U, const double* __restrict__ V,
                                                 const double* __restrict__ den, double* __restrict__ scr,
                                                 bool first_iter, double lambda, const double* __restrict__ ximp,
                                                 int ldimp) {
    constexpr int NA = (KC + 1) / 2;
    constexpr int pitch = dmma_pitch(KC);
    constexpr int PF = 4;  // tiles of X kept in flight per lane (X is L2-resident but far away)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int G = (nown + 7) >> 3;
    const int tiles_total = (nred + 7) >> 3;
    const int tiles_full = nred >> 3;
    const int items = G * S;
    const bool vec_ok = ((nred & 1) == 0);
    const bool last_ok = (NA * 8 - 8 + g) < pitch;
    for (int item = warp; item < items; item += NW) {
        const int grp = item % G, slice = item / G;
        const int row = grp * 8 + g;
        const bool rvalid = row < nown;
        const int rowc = rvalid ? row : nown - 1;  // invalid rows shadow the last valid one; never stored
        const int tb = (int)(((long long)tiles_total * slice) / S), te = (int)(((long long)tiles_total * (slice + 1)) / S);
        const int tef = min(te, tiles_full);  // tiles [tb, tef) are complete
        double ua[KC];
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) ua[kc] = U[(size_t)rowc * pitch + kc * 4 + q];
        double acc[NA][2];
#pragma unroll
        for (int na = 0; na < NA; ++na) acc[na][0] = acc[na][1] = 0.0;
        const double* xp = D + (size_t)rowc * nred + (size_t)tb * 8 + 2 * q;
        const double* vp1 = V + (size_t)(tb * 8 + g) * pitch + q;
        const double* vp2 = V + (size_t)(tb * 8 + 2 * q) * pitch + g;
        auto load_x = [&](const double* ptr) -> double2 {
            if (vec_ok) return __ldg(reinterpret_cast<const double2*>(ptr));
            return make_double2(__ldg(ptr), __ldg(ptr + 1));
        };
        auto fix_nan = [&](double& x, int t) {
            if (HASNAN) {
                if (x != x)
                    x = first_iter ? lambda
                                   : ximp[TRANSPOSED ? ((size_t)t + (size_t)rowc * ldimp) : ((size_t)rowc + (size_t)t * ldimp)];
            }
        };
        double2 xq[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) xq[u] = (tb + u < tef) ? load_x(xp + u * 8) : make_double2(0.0, 0.0);
        for (int tile0 = tb; tile0 < tef; tile0 += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int tile = tile0 + u;
                if (tile < tef) {  // warp-uniform
                    double x0 = xq[u].x, x1 = xq[u].y;
                    xq[u] = (tile + PF < tef) ? load_x(xp + PF * 8) : make_double2(0.0, 0.0);
                    fix_nan(x0, tile * 8 + 2 * q);
                    fix_nan(x1, tile * 8 + 2 * q + 1);
                    dmma_tile<KC>(ua, x0, x1, vp1, vp2, vp2 + pitch, last_ok, acc);
                    xp += 8;
                    vp1 += 8 * pitch;
                    vp2 += 8 * pitch;
                }
            }
        }
        if (te > tef) {
            // the partial last tile of the reduction range: clamp the rows of V, zero X beyond the end
            const int t0 = tef * 8;
            const int ta = t0 + 2 * q;
            double x0 = (ta < nred) ? __ldg(D + (size_t)rowc * nred + ta) : 0.0;
            double x1 = (ta + 1 < nred) ? __ldg(D + (size_t)rowc * nred + ta + 1) : 0.0;
            if (ta < nred) fix_nan(x0, ta);
            if (ta + 1 < nred) fix_nan(x1, ta + 1);
            const double* t1p = V + (size_t)min(t0 + g, nred - 1) * pitch + q;
            const double* t2a = V + (size_t)min(ta, nred - 1) * pitch + g;
            const double* t2b = V + (size_t)min(ta + 1, nred - 1) * pitch + g;
            // steps beyond the end see x = 0 and a real (clamped) row of V: q = 0 / p = 0
            dmma_tile<KC>(ua, x0, x1, t1p, t2a, t2b, last_ok, acc);
        }
        if (S == 1) {
            // (U .* acc) ./ den : the rows of this group are read by this warp only
#pragma unroll
            for (int na = 0; na < NA; ++na)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int a = na * 8 + 2 * q + e;
                    if (rvalid && a < k) {
                        double* up = U + (size_t)row * pitch + a;
                        *up = div_cold<double>(*up * acc[na][e], den[a]);
                    }
                }
        } else {
            double* dst = scr + (size_t)(slice * G + grp) * (NA * 64) + lane;
#pragma unroll
            for (int na = 0; na < NA; ++na)
#pragma unroll
                for (int e = 0; e < 2; ++e) dst[(na * 2 + e) * 32] = acc[na][e];
        }
    }
    __syncthreads();
    if (S > 1) {
        for (int grp = warp; grp < G; grp += NW) {
            const int row = grp * 8 + g;
            const bool rvalid = row < nown;
#pragma unroll
            for (int na = 0; na < NA; ++na)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double s = 0.0;
                    for (int sl = 0; sl < S; ++sl) s += scr[(size_t)(sl * G + grp) * (NA * 64) + (na * 2 + e) * 32 + lane];
                    const int a = na * 8 + 2 * q + e;
                    if (rvalid && a < k) {
                        double* up = U + (size_t)row * pitch + a;
                        *up = div_cold<double>(*up * s, den[a]);
                    }
                }
        }
        __syncthreads();
    }
}

// MODE 0: sum over non-NaN of ((x-p)w)^2 and (x-p)^2 (RESTORE: substituted zeros count as 0);
// MODE 1: imputation X[inan] = (W*H)[inan] into ximp.  X layout (own = rows).
template <int KC, int MODE, bool RESTORE>
__device__ __forceinline__ double2 dmma_residual_pass(const double* __restrict__ X, int n, int m, const double* __restrict__ Ws,
                                                      const double* __restrict__ Hs, double lambda, double weight,
                                                      double* __restrict__ ximp, double* red) {
    constexpr int pitch = dmma_pitch(KC);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int G = (n + 7) >> 3, tiles = (m + 7) >> 3;
    double sw = 0.0, s1 = 0.0;
    for (int grp = warp; grp < G; grp += NW) {
        const int row = grp * 8 + g;
        const bool rvalid = row < n;
        double ua[KC];
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) ua[kc] = rvalid ? Ws[(size_t)row * pitch + kc * 4 + q] : 0.0;
        for (int tile = 0; tile < tiles; ++tile) {
            const int t0 = tile * 8;
            double p[2] = {0.0, 0.0};
            const int tg = min(t0 + g, m - 1);
            const double* vrow = Hs + (size_t)tg * pitch + q;
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) dmma884(p, ua[kc], vrow[kc * 4]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int t = t0 + 2 * q + e;
                if (!rvalid || t >= m) continue;
                const double xr = __ldg(X + (size_t)row + (size_t)t * n);
                if (MODE == 0) {
                    if (xr != xr) continue;
                    double x = xr;
                    if (RESTORE && x == lambda) x = 0.0;
                    const double d = x - p[e];
                    s1 = fma(d, d, s1);
                    const double dw = d * weight;
                    sw = fma(dw, dw, sw);
                } else {
                    if (xr != xr) ximp[(size_t)row + (size_t)t * n] = p[e];
                }
            }
        }
    }
    double2 out;
    out.x = out.y = 0.0;
    if (MODE == 0) {
        out.x = block_sum(sw, red);
        out.y = block_sum(s1, red);
    } else {
        __syncthreads();
    }
    return out;
}

template <int KC, bool HASNAN>
__global__ void __launch_bounds__(dmma_max_threads(KC), 1) kl_resident_dmma_kernel(const SolveArgs a) {
    constexpr int pitch = dmma_pitch(KC);
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, m = a.m, k = a.k;
    const int r = blockIdx.x;
    const int tid = threadIdx.x, NT = blockDim.x;
    UnitState* stg = a.st + r;
    if (stg->done) return;

    const DmmaSmem L = DmmaSmem::make(n, m, KC, a.SH, a.SW);
    double* Ws = reinterpret_cast<double*>(smem + L.off_W);
    double* Hs = reinterpret_cast<double*>(smem + L.off_H);
    double* den = reinterpret_cast<double*>(smem + L.off_den);
    double* scr = reinterpret_cast<double*>(smem + L.off_scr);
    double* red = reinterpret_cast<double*>(smem + L.off_red);
    int* idx = reinterpret_cast<int*>(smem + L.off_idx);
    int* first = reinterpret_cast<int*>(smem + L.off_first);

    const double* X = static_cast<const double*>(a.X);
    const double* Xt = static_cast<const double*>(a.Xt);
    double* Wg = static_cast<double*>(a.W) + (size_t)r * n * k;
    double* Hg = static_cast<double*>(a.H) + (size_t)r * k * m;
    int* canon_old = a.canon + (size_t)r * m;
    double* ximp = HASNAN ? static_cast<double*>(a.ximp) + (size_t)r * n * m : nullptr;
    const double lambda = a.lambda;

    for (int e = tid; e < n * pitch; e += NT) Ws[e] = 0.0;
    for (int e = tid; e < m * pitch; e += NT) Hs[e] = 0.0;
    __syncthreads();
    for (int e = tid; e < n * k; e += NT) Ws[(size_t)(e % n) * pitch + e / n] = Wg[e];
    for (int e = tid; e < k * m; e += NT) Hs[(size_t)(e / k) * pitch + e % k] = Hg[e];
    __syncthreads();

    int it = stg->it, bad = stg->bad, re = stg->re, inc = stg->inc, has_cons = stg->has_cons;
    double best = stg->best, obj_chk = stg->obj_chk;
    int stop = 0;
    __syncthreads();

    if (HASNAN && it > 0) dmma_residual_pass<KC, 1, false>(X, n, m, Ws, Hs, lambda, 1.0, ximp, red);

    while (true) {
        if (it >= a.maxiter) {  // :64
            stop = 1;
            break;
        }
        if (bad >= a.maxbad) {
            stop = 5;
            break;
        }
        if (re >= a.maxre) {
            stop = 3;
            break;
        }
        if (a.iter_limit > 0 && it >= a.iter_limit) break;
        ++it;
        const bool first_iter = (it == 1);
        if (!a.Hfixed) {  // :66-68
            factor_sums_p(Ws, n, pitch, k, den);
            __syncthreads();
            dmma_half_update<KC, true, HASNAN>(X, m, n, k, a.SH, Hs, Ws, den, scr, first_iter, lambda, ximp, n);
        }
        if (!a.Wfixed) {  // :69-71
            factor_sums_p(Hs, m, pitch, k, den);
            __syncthreads();
            dmma_half_update<KC, false, HASNAN>(Xt, n, m, k, a.SW, Ws, Hs, den, scr, first_iter, lambda, ximp, n);
        }
        if (HASNAN) dmma_residual_pass<KC, 1, false>(X, n, m, Ws, Hs, lambda, 1.0, ximp, red);  // :72
        if (it % a.check_every == 0) {                                                           // :73
            const double2 ob = dmma_residual_pass<KC, 0, false>(X, n, m, Ws, Hs, lambda, a.weight, nullptr, red);
            const double obj = ob.x;  // :74
            obj_chk = obj;
            if (obj < a.tol) {  // :75-78
                stop = 2;
                break;
            }
            if (obj < best) {  // :79-89
                if ((best - obj) < a.tolOF)
                    ++bad;
                else
                    bad = 0;
                best = obj;
            } else {
                ++bad;
            }
            if (bad >= a.maxbad) {  // :90-95
                ++re;
                bad = 0;
            }
            const double epsc = a.eps_clamp;  // :99-100
            for (int e = tid; e < n * pitch; e += NT) {
                if (e % pitch < k) {
                    const double v = Ws[e];
                    Ws[e] = (v != v) ? v : (v < epsc ? epsc : v);
                }
            }
            for (int e = tid; e < m * pitch; e += NT) {
                if (e % pitch < k) {
                    const double v = Hs[e];
                    Hs[e] = (v != v) ? v : (v < epsc ? epsc : v);
                }
            }
            if (tid < 36) first[tid] = INT_MAX;
            __syncthreads();
            for (int j = tid; j < m; j += NT) {  // :101-103
                const double* h = Hs + (size_t)j * pitch;
                double bv = h[0];
                int bi = 0;
                for (int c = 1; c < k; ++c) {
                    const double v = h[c];
                    const bool take = (bv != bv) ? false : ((v != v) ? true : (v < bv));
                    if (take) {
                        bv = v;
                        bi = c;
                    }
                }
                idx[j] = bi;
                atomicMin(&first[bi], j);
            }
            __syncthreads();
            int same = 1;  // :105-111
            for (int j = tid; j < m; j += NT) {
                const int c = first[idx[j]];
                if (!has_cons || canon_old[j] != c) same = 0;
                idx[j] = c;
            }
            same = __syncthreads_and(same);
            if (same)
                ++inc;
            else
                inc = 0;
            if (inc > a.stopconv) {  // :112-115
                stop = 4;
                break;
            }
            for (int j = tid; j < m; j += NT) canon_old[j] = idx[j];  // :116
            has_cons = 1;
            __syncthreads();
        }
    }

    double obj_ssq = stg->obj_ssq, obj_norm = stg->obj_norm;
    int done = 0;
    if (stop != 0) {
        const double2 ob = dmma_residual_pass<KC, 0, true>(X, n, m, Ws, Hs, lambda, a.weight, nullptr, red);
        obj_ssq = ob.x;           // NMFkMultiplicative.jl:125
        obj_norm = sqrt(ob.y);    // NMFkExecute.jl:792
        if (a.normalize == 1) {   // NMFkExecute.jl:800-804
            factor_sums_p(Hs, m, pitch, k, den);
            __syncthreads();
            for (int e = tid; e < n * pitch; e += NT)
                if (e % pitch < k) Ws[e] = Ws[e] * den[e % pitch];
            for (int e = tid; e < m * pitch; e += NT)
                if (e % pitch < k) Hs[e] = div_cold<double>(Hs[e], den[e % pitch]);
        } else if (a.normalize == 2) {  // :796-799
            factor_sums_p(Ws, n, pitch, k, den);
            __syncthreads();
            for (int e = tid; e < n * pitch; e += NT)
                if (e % pitch < k) Ws[e] = div_cold<double>(Ws[e], den[e % pitch]);
            for (int e = tid; e < m * pitch; e += NT)
                if (e % pitch < k) Hs[e] = Hs[e] * den[e % pitch];
        }
        done = 1;
        __syncthreads();
    }
    for (int e = tid; e < n * k; e += NT) Wg[e] = Ws[(size_t)(e % n) * pitch + e / n];
    for (int e = tid; e < k * m; e += NT) Hg[e] = Hs[(size_t)(e / k) * pitch + e % k];
    if (tid == 0) {
        stg->it = it;
        stg->bad = bad;
        stg->re = re;
        stg->inc = inc;
        stg->stop = stop;
        stg->has_cons = has_cons;
        stg->done = done;
        stg->best = best;
        stg->obj_chk = obj_chk;
        stg->obj_ssq = obj_ssq;
        stg->obj_norm = obj_norm;
    }
}

}  // namespace nmfk
