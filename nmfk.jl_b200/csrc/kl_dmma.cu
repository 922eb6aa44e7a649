// Float64 DMMA instantiations of the resident engine + the host heuristic that picks the CTA
// width and the reduction slices per half-update (see kl_dmma.cuh).
#include "kl_dmma.cuh"

namespace nmfk {

namespace {

constexpr size_t kSmemLimit = 227u * 1024u;

struct DmmaPlan {
    int nthreads, SH, SW;
    size_t smem;
    bool ok;
};

// Minimise the per-iteration critical path (in 8-step tiles) of one warp: a half-update has
// G = ceil(own/8) row groups x S slices work items dealt round-robin to NW warps.
DmmaPlan plan_dmma(int n, int m, int KC) {
    const int maxthreads = dmma_max_threads(KC);
    const int GH = (m + 7) / 8, TH = (n + 7) / 8;  // H-update: own = columns, reduction = rows
    const int GW = (n + 7) / 8, TW = (m + 7) / 8;
    DmmaPlan best{0, 1, 1, 0, false};
    long long bestcost = -1;
    for (int NW = 8; NW * 32 <= maxthreads; ++NW) {
        for (int SH = 1; SH <= 4 && SH <= TH; ++SH)
            for (int SW = 1; SW <= 4 && SW <= TW; ++SW) {
                const size_t smem = DmmaSmem::make(n, m, KC, SH, SW).total;
                if (smem > kSmemLimit) continue;
                auto path = [&](int G, int T, int S) {
                    const long long rounds = ((long long)G * S + NW - 1) / NW;
                    return rounds * ((T + S - 1) / S) + (S > 1 ? 2 : 0);
                };
                const long long cost = (path(GH, TH, SH) + path(GW, TW, SW)) * 64 + NW;  // ties -> fewer warps
                if (bestcost < 0 || cost < bestcost) {
                    bestcost = cost;
                    best = DmmaPlan{NW * 32, SH, SW, smem, true};
                }
            }
    }
    return best;
}

template <int KC>
cudaError_t launch_dmma_kc(SolveArgs a, cudaStream_t s) {
    const DmmaPlan pl = plan_dmma(a.n, a.m, KC);
    if (!pl.ok) return cudaErrorInvalidConfiguration;
    a.SH = pl.SH;
    a.SW = pl.SW;
    cudaError_t e;
    if (a.has_nan) {
        e = cudaFuncSetAttribute(kl_resident_dmma_kernel<KC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess) return e;
        kl_resident_dmma_kernel<KC, true><<<a.R, pl.nthreads, pl.smem, s>>>(a);
    } else {
        e = cudaFuncSetAttribute(kl_resident_dmma_kernel<KC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess) return e;
        kl_resident_dmma_kernel<KC, false><<<a.R, pl.nthreads, pl.smem, s>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace

bool resident_dmma_fits(int n, int m, int k) {
    if (k < 1 || k > kMaxK) return false;
    return plan_dmma(n, m, (k + 3) / 4).ok;
}

cudaError_t launch_kl_resident_dmma(const SolveArgs& a, cudaStream_t s) {
    switch ((a.k + 3) / 4) {
        case 1: return launch_dmma_kc<1>(a, s);
        case 2: return launch_dmma_kc<2>(a, s);
        case 3: return launch_dmma_kc<3>(a, s);
        case 4: return launch_dmma_kc<4>(a, s);
        case 5: return launch_dmma_kc<5>(a, s);
        case 6: return launch_dmma_kc<6>(a, s);
        case 7: return launch_dmma_kc<7>(a, s);
        case 8: return launch_dmma_kc<8>(a, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace nmfk
