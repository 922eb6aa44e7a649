// Philox4x64-10 counter-based generator, bit-compatible with numpy.random.Philox:
// element i (0-based) of Generator(Philox(key=s)).random() is
//   to_double(philox4x64_10(counter = i/4 + 1, key = {s, 0})[i % 4])
// (numpy increments the 256-bit counter before producing each block of four outputs).
// Stands in for Julia's `rand(n,k)`, `rand(k,m)` (/root/reference/src/NMFkMultiplicative.jl:38,48):
// Julia's Xoshiro stream cannot be reproduced without Julia, and the reference accepts explicit
// Winit/Hinit, so the harness defines its own reproducible streams.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define NMFK_HD __host__ __device__ __forceinline__
#else
#define NMFK_HD inline
#endif

namespace nmfk {

NMFK_HD void mulhilo64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    lo = (uint64_t)p;
    hi = (uint64_t)(p >> 64);
#endif
}

NMFK_HD void philox4x64_10(uint64_t counter_lo, uint64_t key_lo, uint64_t out[4]) {
    const uint64_t M0 = 0xD2E7470EE14C6C93ull, M1 = 0xCA5A826395121157ull;
    const uint64_t W0 = 0x9E3779B97F4A7C15ull, W1 = 0xBB67AE8584CAA73Bull;
    uint64_t c0 = counter_lo, c1 = 0, c2 = 0, c3 = 0, k0 = key_lo, k1 = 0;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r > 0) {
            k0 += W0;
            k1 += W1;
        }
        uint64_t hi0, lo0, hi1, lo1;
        mulhilo64(M0, c0, hi0, lo0);
        mulhilo64(M1, c2, hi1, lo1);
        const uint64_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

NMFK_HD double philox_to_double(uint64_t v) { return (double)(v >> 11) * (1.0 / 9007199254740992.0); }

}  // namespace nmfk
