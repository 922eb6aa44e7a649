// Timing probe for the resident DMMA kernel: compiles kl_dmma.cuh with -DNMFK_PROBE=<n> (pieces of the hot
// loop removed; results are then WRONG - timing only) and runs K in {4, 8} on the C2 shape, 148 restarts x 100 iterations.
#include <cstdio>
#include <vector>
#include "../../nmfk.jl_b200/csrc/kl_dmma.cuh"
using namespace nmfk;

template <int K>
float run(int n, int m, int R, int iters, int NW, int SH, int SW) {
    const int KC = (K + 3) / 4;
    std::vector<double> X((size_t)n * m), W((size_t)n * K * R), H((size_t)K * m * R);
    unsigned s = 12345;
    auto rnd = [&] { s = s * 1664525u + 1013904223u; return 0.05 + (s >> 8) / 16777216.0; };
    for (auto& v : X) v = rnd();
    for (auto& v : W) v = rnd();
    for (auto& v : H) v = rnd();
    std::vector<double> Xt((size_t)n * m);
    for (int i = 0; i < n; ++i) for (int j = 0; j < m; ++j) Xt[(size_t)j + (size_t)i * m] = X[(size_t)i + (size_t)j * n];
    double *dX, *dXt, *dW, *dH; UnitState* st; int* canon;
    cudaMalloc(&dX, X.size() * 8); cudaMalloc(&dXt, X.size() * 8); cudaMalloc(&dW, W.size() * 8); cudaMalloc(&dH, H.size() * 8);
    cudaMalloc(&st, R * sizeof(UnitState)); cudaMalloc(&canon, (size_t)R * m * 4);
    cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dXt, Xt.data(), X.size() * 8, cudaMemcpyHostToDevice);
    SolveArgs a{}; a.X = dX; a.Xt = dXt; a.W = dW; a.H = dH; a.st = st; a.canon = canon; a.n = n; a.m = m; a.k = K; a.R = R;
    a.SH = SH; a.SW = SW; a.maxiter = iters; a.maxbad = 1 << 30; a.maxre = 1 << 30; a.stopconv = 1 << 30; a.check_every = 10;
    a.normalize = 1; a.lambda = 1e-32; a.tol = -1; a.tolOF = 1e-3; a.eps_clamp = 2.2e-16; a.weight = 1;
    const size_t smem = DmmaSmem::make(n, m, KC, SH, SW).total;
    cudaFuncSetAttribute(kl_resident_dmma_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        std::vector<UnitState> h(R); for (auto& u : h) { u = UnitState{}; u.best = 1e300; }
        cudaMemcpy(st, h.data(), R * sizeof(UnitState), cudaMemcpyHostToDevice);
        cudaMemcpy(dW, W.data(), W.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dH, H.data(), H.size() * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        kl_resident_dmma_kernel<K, false><<<R, NW * 32, smem>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
    cudaFree(dX); cudaFree(dXt); cudaFree(dW); cudaFree(dH); cudaFree(st); cudaFree(canon);
    return best;
}

int main(int argc, char** argv) {
    const int NW = argc > 1 ? atoi(argv[1]) : 20, SH = argc > 2 ? atoi(argv[2]) : 4, SW = argc > 3 ? atoi(argv[3]) : 1;
    const int iters = 100, R = 148;
#ifndef NMFK_PROBE
#define NMFK_PROBE 0
#endif
    const float t4 = run<4>(1000, 200, R, iters, NW, SH, SW), t8 = run<8>(1000, 200, R, iters, NW, SH, SW);
    printf("{\"probe\": %d, \"NW\": %d, \"SH\": %d, \"SW\": %d, \"k4_us_per_iter\": %.2f, \"k8_us_per_iter\": %.2f}\n", NMFK_PROBE, NW, SH, SW,
           t4 * 1e3 / iters, t8 * 1e3 / iters);
    return 0;
}
