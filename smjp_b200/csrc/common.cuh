// Shared device helpers: math traits per precision, Philox4x32-10, block reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smjp {

// ------------------------------------------------------------------------------------------------
// Math traits.  Everything is phrased in base 2 so that the fp32 path maps onto the SFU
// (ex2.approx / lg2.approx) with no extra multiplies: H(tau) = 2^(k*log2(tau) - k*log2(lambda)).
// ------------------------------------------------------------------------------------------------
template <typename R>
struct Math;

template <>
struct Math<double> {
    static __device__ __forceinline__ double exp2(double x) { return ::exp2(x); }
    static __device__ __forceinline__ double log2(double x) { return ::log2(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double pow(double x, double y) { return ::pow(x, y); }
};

template <>
struct Math<float> {
    static __device__ __forceinline__ float exp2(float x) {
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    static __device__ __forceinline__ float log2(float x) {
        float y;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    static __device__ __forceinline__ float rcp(float x) {
        float y;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    static __device__ __forceinline__ float pow(float x, float y) { return exp2(y * log2(x)); }
};

#define SMJP_LOG2E 1.4426950408889634074

// ------------------------------------------------------------------------------------------------
// Philox4x32-10, counter = (r >> 1, a, b, chain), key = (seed_lo, seed_hi ^ tag).
// Bit-exact twin of oracle/philox.py.
// ------------------------------------------------------------------------------------------------
enum : uint32_t {
    TAG_CANDIDATES = 0x1001,
    TAG_BACKWARD = 0x1002,
    TAG_PRIOR = 0x1003,
    TAG_PF_PROPAGATE = 0x1004,
    TAG_PF_RESAMPLE = 0x1005,
    TAG_DATA = 0x1006,
};

struct Philox4 {
    uint32_t x[4];
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r > 0) {
            k0 += W0;
            k1 += W1;
        }
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
    }
    Philox4 o;
    o.x[0] = c0;
    o.x[1] = c1;
    o.x[2] = c2;
    o.x[3] = c3;
    return o;
}

__host__ __device__ __forceinline__ double uniform53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// uniform number r of stream (seed, tag, a, b, chain)
__host__ __device__ __forceinline__ double philox_u01(uint64_t seed, uint32_t tag, uint64_t r,
                                                      uint32_t a, uint32_t b, uint32_t chain) {
    Philox4 o = philox4x32_10((uint32_t)(r >> 1), a, b, chain, (uint32_t)seed,
                              (uint32_t)(seed >> 32) ^ tag);
    return (r & 1) ? uniform53(o.x[2], o.x[3]) : uniform53(o.x[0], o.x[1]);
}

// ------------------------------------------------------------------------------------------------
// Warp helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

}  // namespace smjp
