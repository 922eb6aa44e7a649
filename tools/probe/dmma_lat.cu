// DMMA m8n8k4 dependent-issue latency and per-scheduler throughput vs. independent chains / warps.
#include <cstdio>
#define DMMA(c, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b))
template <int CH>
__global__ void lat(double* out, int iters, double a, double b) {
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) DMMA(c[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = (double)(t1 - t0) / (iters * CH); out[blockIdx.x * 2 + 1] = s; }
}
// chain like the KL tile: DMMA -> rcp seq -> DMMA(acc)
__global__ void tilechain(double* out, int iters, double a, double b) {
    double p[2], acc[2] = {0, 0};
    double x = 1.0 + threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        p[0] = p[1] = 0.0;
        DMMA(p, x, b);
        DMMA(p, a, b);
        double q[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double r0;
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(p[e]));
            const double er = fma(-p[e], r0, 1.0);
            const double t = fma(er, er, er);
            const double q0 = a * r0;
            q[e] = fma(q0, t, q0);
        }
        DMMA(acc, q[0], b);
        DMMA(acc, q[1], b);
        x = acc[0] * 1e-30 + 1.0;  // serialise tiles like an in-order warp would
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / iters; out[1] = acc[0] + acc[1]; }
}
int main() {
    double* d; cudaMalloc(&d, 4096); double h[512];
    const int iters = 4096;
#define RUN(CH, NWARPS)                                                                  \
    lat<CH><<<1, 32 * NWARPS>>>(d, iters, 1.0000001, 1e-9); lat<CH><<<1, 32 * NWARPS>>>(d, iters, 1.0000001, 1e-9); \
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);                                        \
    printf("chains=%d warps=%d: %.2f cycles per DMMA per warp -> %.2f per DMMA per scheduler\n", CH, NWARPS, h[0], h[0] / ((NWARPS + 3) / 4));
    RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(1, 4) RUN(1, 8) RUN(1, 16) RUN(2, 16) RUN(1, 20) RUN(2, 20) RUN(4, 20)
    tilechain<<<1, 32>>>(d, iters, 1.0000001, 1e-3); tilechain<<<1, 32>>>(d, iters, 1.0000001, 1e-3);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("serial KL tile chain (2 DMMA -> rcp -> 2 DMMA), one warp: %.1f cycles\n", h[0]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
