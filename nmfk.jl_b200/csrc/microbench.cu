// Micro-benchmarks for the roofline denominators MEASURED_PEAKS.json does not hold:
// FP64 DFMA, FP64 DMMA (mma.sync.m8n8k4.f64), FP32 FFMA throughput and a device copy.
// bench.py reports the KL kernels against these ("of measured, own micro-benchmark").
#include "nmfk_internal.h"

namespace nmfk {

namespace {

template <typename T, int CHAINS>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
    T acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = (T)(threadIdx.x + c);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    T s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    if (s == (T)-1.2345) out[0] = s;  // never true; keeps the chain alive
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0}, c2[2] = {0.0, 0.0}, c3[2] = {0.0, 0.0};
    const double av = a + threadIdx.x, bv = b;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[0]), "+d"(c0[1])
                     : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c1[0]), "+d"(c1[1])
                     : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c2[0]), "+d"(c2[1])
                     : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c3[0]), "+d"(c3[1])
                     : "d"(av), "d"(bv));
    }
    const double s = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
    if (s == -1.2345) out[0] = s;
}

// dependent-issue latency (cycles per instruction) of DFMA (which=0), FFMA (1), MUFU.RCP64H+DFMA pair (2)
__global__ void latency_kernel(double* out, int iters, int which, double a, double b) {
    double x = a + threadIdx.x;
    float xf = (float)a + threadIdx.x;
    const long long t0 = clock64();
    if (which == 0) {
        for (int i = 0; i < iters; ++i) x = fma(x, a, b);
    } else if (which == 1) {
        const float af = (float)a, bf = (float)b;
        for (int i = 0; i < iters; ++i) xf = fmaf(xf, af, bf);
    } else {
        for (int i = 0; i < iters; ++i) {
            double r;
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
            x = fma(r, a, b);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) {
        out[0] = (double)(t1 - t0) / iters;
        out[1] = x + xf;
    }
}

// throughput of single instructions / short sequences, 8 independent chains per thread:
// MODE 0 MUFU.RCP64H, 1 MUFU.RCP (f32), 2 one DMMA + 8 DFMA per step (equal FP64-pipe time if shared),
// 3 the reciprocal sequence of the KL hot loop (RCP64H + 4 DFMA + DMUL)
template <int MODE>
__global__ void __launch_bounds__(256) op_peak_kernel(double* out, int iters, double a, double b) {
    double x[8];
    float xf[8];
    double c0[2] = {0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        x[c] = a + 0.001 * (threadIdx.x + c);
        xf[c] = (float)x[c];
    }
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x[c]));
        } else if (MODE == 1) {
#pragma unroll
            for (int c = 0; c < 8; ++c) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(xf[c]));
        } else if (MODE == 2) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[0]), "+d"(c0[1])
                         : "d"(a), "d"(b));
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = fma(x[c], a, b);
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double r;
                asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[c]));
                double e = fma(-x[c], r, 1.0);
                r = fma(r, e, r);
                e = fma(-x[c], r, 1.0);
                r = fma(r, e, r);
                x[c] = r * b;
            }
        }
    }
    double s = c0[0] + c0[1];
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c] + (double)xf[c];
    if (s == -1.2345) out[0] = s;
}

// legacy tensor path for Float32 data: mma.sync m16n8k8 TF32 (the 3xTF32 split of the tiled f32 engine)
// and m16n8k16 BF16 for comparison; 4 independent accumulator sets per warp
template <int MODE>
__global__ void __launch_bounds__(256) hmma_peak_kernel(float* out, int iters, unsigned a, unsigned b) {
    float c[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[q][e] = 0.f;
    const unsigned a0 = a + threadIdx.x, a1 = a, a2 = a ^ 1u, a3 = a, b0 = b, b1 = b ^ 3u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[q][0]), "+f"(c[q][1]), "+f"(c[q][2]), "+f"(c[q][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[q][0]), "+f"(c[q][1]), "+f"(c[q][2]), "+f"(c[q][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) s += c[q][e];
    if (s == -1.2345f) out[0] = s;
}

__global__ void copy_kernel(const double4* __restrict__ src, double4* __restrict__ dst, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

template <typename F>
cudaError_t time_best(F launch, int reps, float* best_ms, cudaStream_t s) {
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) return e;
    *best_ms = 1e30f;
    for (int r = 0; r < reps + 2; ++r) {
        cudaEventRecord(e0, s);
        launch();
        cudaEventRecord(e1, s);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < *best_ms) *best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace

cudaError_t measure_peak(int which, double* value, cudaStream_t s) {
    cudaError_t e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float ms = 0.f;
    if (which == 0 || which == 2) {
        double* out = nullptr;
        if ((e = cudaMalloc(&out, 64)) != cudaSuccess) return e;
        const int blocks = sms * 8, threads = 256, chains = 8;
        const int iters = which == 0 ? 4096 : 16384;
        if (which == 0)
            e = time_best([&] { fma_peak_kernel<double, 8><<<blocks, threads, 0, s>>>(out, iters, 1.0000001, 1e-9); }, 5, &ms, s);
        else
            e = time_best([&] { fma_peak_kernel<float, 8><<<blocks, threads, 0, s>>>((float*)out, iters, 1.0000001f, 1e-9f); }, 5,
                          &ms, s);
        cudaFree(out);
        if (e != cudaSuccess) return e;
        const double flops = 2.0 * (double)blocks * threads * chains * iters;
        *value = flops / (ms * 1e-3) / 1e12;
        return cudaSuccess;
    }
    if (which == 1) {
        double* out = nullptr;
        if ((e = cudaMalloc(&out, 64)) != cudaSuccess) return e;
        const int blocks = sms * 8, threads = 256, iters = 4096;
        e = time_best([&] { dmma_peak_kernel<<<blocks, threads, 0, s>>>(out, iters, 1.0, 1e-9); }, 5, &ms, s);
        cudaFree(out);
        if (e != cudaSuccess) return e;
        // per warp per mma: 8*8*4 FMAs = 512 flops; 4 mma per iteration
        const double flops = 512.0 * 4.0 * iters * (double)blocks * (threads / 32);
        *value = flops / (ms * 1e-3) / 1e12;
        return cudaSuccess;
    }
    if (which == 3) {
        const long long bytes = 1ll << 30;
        double4 *a = nullptr, *b = nullptr;
        if ((e = cudaMalloc(&a, bytes)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&b, bytes)) != cudaSuccess) {
            cudaFree(a);
            return e;
        }
        cudaMemsetAsync(a, 1, bytes, s);
        const long long n4 = bytes / sizeof(double4);
        e = time_best([&] { copy_kernel<<<sms * 16, 512, 0, s>>>(a, b, n4); }, 5, &ms, s);
        cudaFree(a);
        cudaFree(b);
        if (e != cudaSuccess) return e;
        *value = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
        return cudaSuccess;
    }
    if (which >= 4 && which <= 6) {  // latencies in SM cycles: 4 DFMA, 5 FFMA, 6 RCP64H+DFMA
        double* out = nullptr;
        if ((e = cudaMalloc(&out, 64)) != cudaSuccess) return e;
        latency_kernel<<<1, 32, 0, s>>>(out, 8192, which - 4, 1.0000001, 1e-9);
        latency_kernel<<<1, 32, 0, s>>>(out, 8192, which - 4, 1.0000001, 1e-9);
        double h[2] = {0, 0};
        e = cudaMemcpyAsync(h, out, sizeof(h), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaFree(out);
        if (e != cudaSuccess) return e;
        *value = h[0];
        return cudaSuccess;
    }
    if (which >= 7 && which <= 10) {
        // 7: MUFU.RCP64H, 8: MUFU.RCP f32, 10: KL reciprocal sequence -> SM cycles per warp instruction (sequence)
        // per scheduler; 9: DMMA + DFMA mixed -> TFLOP/s of the two together
        double* out = nullptr;
        if ((e = cudaMalloc(&out, 64)) != cudaSuccess) return e;
        const int blocks = sms * 8, threads = 256, iters = 2048;
        if (which == 7)
            e = time_best([&] { op_peak_kernel<0><<<blocks, threads, 0, s>>>(out, iters, 1.5, 0.75); }, 3, &ms, s);
        else if (which == 8)
            e = time_best([&] { op_peak_kernel<1><<<blocks, threads, 0, s>>>(out, iters, 1.5, 0.75); }, 3, &ms, s);
        else if (which == 9)
            e = time_best([&] { op_peak_kernel<2><<<blocks, threads, 0, s>>>(out, iters, 1.0000001, 1e-9); }, 3, &ms, s);
        else
            e = time_best([&] { op_peak_kernel<3><<<blocks, threads, 0, s>>>(out, iters, 1.5, 0.75); }, 3, &ms, s);
        cudaFree(out);
        if (e != cudaSuccess) return e;
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
        if (which == 9) {
            const double flops = (512.0 + 8.0 * 64.0) * iters * (double)blocks * (threads / 32);
            *value = flops / (ms * 1e-3) / 1e12;
        } else {
            // warp instructions (sequences) issued per scheduler, at the maximum SM clock
            const double per_smsp = 8.0 * iters * (double)blocks * (threads / 32) / (sms * 4.0);
            *value = (ms * 1e-3) * (double)khz * 1e3 / per_smsp;
        }
        return cudaSuccess;
    }
    if (which == 11 || which == 12) {  // mma.sync TF32 m16n8k8 / BF16 m16n8k16 -> TFLOP/s
        float* out = nullptr;
        if ((e = cudaMalloc(&out, 64)) != cudaSuccess) return e;
        const int blocks = sms * 8, threads = 256, iters = 4096;
        if (which == 11)
            e = time_best([&] { hmma_peak_kernel<0><<<blocks, threads, 0, s>>>(out, iters, 0x3f800000u, 0x3f000000u); }, 5, &ms, s);
        else
            e = time_best([&] { hmma_peak_kernel<1><<<blocks, threads, 0, s>>>(out, iters, 0x3f803f80u, 0x3f003f00u); }, 5, &ms, s);
        cudaFree(out);
        if (e != cudaSuccess) return e;
        const double flops = (which == 11 ? 2048.0 : 4096.0) * 4.0 * iters * (double)blocks * (threads / 32);
        *value = flops / (ms * 1e-3) / 1e12;
        return cudaSuccess;
    }
    if (which == 13) return umma_peak(value, s);  // dense tcgen05.mma kind::tf32 TFLOP/s (tc_selftest.cu)
    return cudaErrorInvalidValue;
}

}  // namespace nmfk
