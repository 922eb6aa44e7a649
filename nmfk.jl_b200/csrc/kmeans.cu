// robustkmeans (/root/reference/src/NMFkCluster.jl:172-246): `repeats` independent runs of Clustering.kmeans(X, k;
// distance=CosineDist(), maxiter, tol) on the columns of X, the run with the smallest total cost wins.
// Third-party semantics restated (Clustering.jl 0.14 / 0.15 `kmeans.jl`, not vendored in the reference tree; the same
// restatement is oracle/nmfk_oracle.py::kmeans_lloyd): distances centre -> point = max(1 - <c,x> / (|c||x|), 0)
// (Distances.pairwise(CosineDist())), a point goes to the FIRST closest centre (strict '<' scan), the centres of the clusters
// whose membership changed become the arithmetic mean of their points, the run stops when the total cost changes by less than
// tol (or at maxiter).  The k-means++ seeding uses Julia's RNG in the reference and cannot be reproduced: the seeds (k point
// indices per repeat) are an input.  A centre that loses all its points is re-drawn at random by the package
// (`repick_unused_centers`); here it keeps its position and the repeat is flagged.
// One CTA per repeat - the repeats are embarrassingly parallel (1000 by default); X (d x N, a few thousand points of
// dimension <= 64 in NMFk's use: the columns of W[kopt]' or H[kopt]) is shared through L2.
#include <cfloat>

#include "nmfk_internal.h"

namespace nmfk {
namespace {

constexpr int KM_THREADS = 256;

__device__ __forceinline__ double km_wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// X: d x N column-major (double); xnorm[j] = |X[:, j]|; seeds: k x repeats (0-based point indices)
// outputs per repeat: assign (N, 1-based), costs (N), counts (k), centers (d x k), totalcost, iters, flags (bit 0 converged,
// bit 1 an empty cluster occurred)
__global__ void __launch_bounds__(KM_THREADS) kmeans_kernel(const double* __restrict__ X, const double* __restrict__ xnorm, int d, int N, int k,
                                                            const int* __restrict__ seeds, int maxiter, double tol, int* __restrict__ assign_all,
                                                            double* __restrict__ costs_all, int* __restrict__ counts_all,
                                                            double* __restrict__ centers_all, double* __restrict__ totalcost,
                                                            int* __restrict__ iters, int* __restrict__ flags) {
    extern __shared__ double ksm[];
    double* cent = ksm;                   // d x k
    double* cnorm = cent + (size_t)d * k;  // k
    double* red = cnorm + k;              // KM_THREADS / 32
    int* counts = reinterpret_cast<int*>(red + KM_THREADS / 32);  // k
    int* upd = counts + k;                                        // k : to_update
    __shared__ double s_objv;
    __shared__ int s_changed;
    const int rep = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = KM_THREADS / 32;
    int* assign = assign_all + (long long)rep * N;
    double* costs = costs_all + (long long)rep * N;
    for (int e = tid; e < d * k; e += KM_THREADS) cent[e] = X[(long long)seeds[(long long)rep * k + e / d] * d + e % d];  // copyseeds
    __syncthreads();

    auto center_norms = [&]() {
        for (int c = warp; c < k; c += NW) {
            double s = 0.0;
            for (int i = lane; i < d; i += 32) s = fma(cent[c * d + i], cent[c * d + i], s);
            s = km_wsum(s);
            if (lane == 0) cnorm[c] = sqrt(s);
        }
        __syncthreads();
    };
    // update_assignments!: closest centre (first minimum), costs, counts, to_update; returns the total cost
    auto assign_step = [&](bool is_init) -> double {
        for (int c = tid; c < k; c += KM_THREADS) {
            counts[c] = 0;
            upd[c] = is_init ? 1 : 0;
        }
        __syncthreads();
        double local = 0.0;
        for (int j = tid; j < N; j += KM_THREADS) {
            const double* x = X + (long long)j * d;
            const double xn = xnorm[j];
            int best = 0;
            double bv = 0.0;
            for (int c = 0; c < k; ++c) {
                double dot = 0.0;
                for (int i = 0; i < d; ++i) dot = fma(cent[c * d + i], x[i], dot);
                double dist = 1.0 - dot / (cnorm[c] * xn);
                dist = (dist != dist) ? dist : (dist < 0.0 ? 0.0 : dist);  // max(., 0) keeps NaN
                if (c == 0 || dist < bv) {
                    bv = dist;
                    best = c;
                }
            }
            if (is_init) {
                assign[j] = best + 1;
            } else {
                const int pa = assign[j] - 1;
                if (pa != best) {
                    assign[j] = best + 1;
                    upd[best] = 1;  // benign race: every writer stores 1
                    upd[pa] = 1;
                }
            }
            costs[j] = bv;
            atomicAdd(&counts[best], 1);
        }
        __syncthreads();
        // sum(costs) in point order blocks: fixed-shape tree -> deterministic
        for (int j = tid; j < N; j += KM_THREADS) local += costs[j];
        local = km_wsum(local);
        if (lane == 0) red[warp] = local;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < NW; ++w) s += red[w];
            s_objv = s;
        }
        __syncthreads();
        return s_objv;
    };

    center_norms();
    double objv = assign_step(true);
    int t = 0, converged = 0, empty_seen = 0;
    while (!converged && t < maxiter) {
        ++t;
        // update_centers!: clusters whose membership changed become the mean of their points
        for (int c = warp; c < k; c += NW) {
            if (!upd[c] || counts[c] == 0) continue;
            for (int i = 0; i < d; ++i) {
                double s = 0.0;
                for (int j = lane; j < N; j += 32)
                    if (assign[j] - 1 == c) s += X[(long long)j * d + i];
                s = km_wsum(s);
                if (lane == 0) cent[c * d + i] = s / (double)counts[c];
            }
        }
        if (tid == 0) {
            int e = 0;
            for (int c = 0; c < k; ++c) e |= (counts[c] == 0);
            s_changed = e;
        }
        __syncthreads();
        empty_seen |= s_changed;
        center_norms();
        const double prev = objv;
        objv = assign_step(false);
        const double change = objv - prev;
        if (!(change > tol) && (k == 1 || fabs(change) < tol)) converged = 1;
    }
    if (tid == 0) {
        totalcost[rep] = objv;
        iters[rep] = t;
        flags[rep] = converged | (empty_seen << 1);
    }
    for (int c = tid; c < k; c += KM_THREADS) counts_all[(long long)rep * k + c] = counts[c];
    for (int e = tid; e < d * k; e += KM_THREADS) centers_all[(long long)rep * d * k + e] = cent[e];
}

__global__ void col_norm_kernel(const double* __restrict__ X, int d, int N, double* __restrict__ xnorm) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    double s = 0.0;
    for (int i = 0; i < d; ++i) s = fma(X[(long long)j * d + i], X[(long long)j * d + i], s);
    xnorm[j] = sqrt(s);
}

}  // namespace

cudaError_t launch_kmeans(const double* X, double* xnorm, int d, int N, int k, int repeats, const int* seeds, int maxiter, double tol,
                          int* assign, double* costs, int* counts, double* centers, double* totalcost, int* iters, int* flags,
                          cudaStream_t s) {
    col_norm_kernel<<<(N + 255) / 256, 256, 0, s>>>(X, d, N, xnorm);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t smem = ((size_t)d * k + k + KM_THREADS / 32) * sizeof(double) + 2 * (size_t)k * sizeof(int);
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kmeans_kernel<<<repeats, KM_THREADS, smem, s>>>(X, xnorm, d, N, k, seeds, maxiter, tol, assign, costs, counts, centers, totalcost,
                                                    iters, flags);
    return cudaGetLastError();
}

}  // namespace nmfk
