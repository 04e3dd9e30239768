// Per-precision launcher of ffbs_kernel: picks the <K, MAXT, MINB> instantiation.  Included by
// ffbs_f64.cu / ffbs_f32.cu so that the two precisions compile in parallel.
#pragma once
#include "ffbs.cuh"
#include "ffbs_item.cuh"
#include "ffbs_fast.cuh"
#include "ffbs_fastk.cuh"

#ifndef FFBS_FAST_MINB_F64_128
#define FFBS_FAST_MINB_F64_128 5  // resident CTAs per SM the fp64 128-thread fast kernel is compiled for (96 registers)
#endif

namespace smjp {

struct FfbsLaunch {
    int K;
    int threads;
    size_t smem;
    int grid;           // in: requested (0 = query only); out of query: resident CTAs per SM
    cudaStream_t stream;
};

template <typename Real, int K, int B, int MINB>
cudaError_t ffbs_do(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    auto kern = ffbs_kernel<Real, K, B, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return e;
    if (query) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L.grid, kern, B, L.smem);
    kern<<<L.grid, B, L.smem, L.stream>>>(a);
    return cudaGetLastError();
}

// block sizes the kernel is instantiated for
inline int ffbs_round_threads(int t) { return t <= 32 ? 32 : t <= 64 ? 64 : t <= 96 ? 96 : t <= 128 ? 128 : t <= 256 ? 256 : 512; }

template <typename Real, int K>
cudaError_t ffbs_pick(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    // resident CTAs per SM the register budget is tuned for: fp32 needs about half the registers of
    // fp64, K >= 6 about twice those of K <= 5 (K*K hazards per grid point are kept in flight)
    constexpr int S = (sizeof(Real) == 8 ? 1 : 2) * (K <= 5 ? 2 : 1);  // 1, 2 or 4
    // (fp64, K <= 3 at 128 threads: 5 CTAs/SM = 96 registers measured 9 % faster than 4 CTAs/SM = 128 and
    //  5 % faster than 6 CTAs/SM = 80)
    switch (L.threads) {
        case 32: return ffbs_do<Real, K, 32, 8 * S>(a, L, query);
        case 64: return ffbs_do<Real, K, 64, 4 * S>(a, L, query);
        case 96: return ffbs_do<Real, K, 96, sizeof(Real) == 8 ? 6 : 10>(a, L, query);
        case 128: return ffbs_do<Real, K, 128, (sizeof(Real) == 8 && K <= 3) ? 5 : 2 * S>(a, L, query);
        case 256: return ffbs_do<Real, K, 256, S>(a, L, query);
        case 512: return ffbs_do<Real, K, 512, (S + 1) / 2>(a, L, query);
    }
    return cudaErrorInvalidValue;
}

// ---- K <= 3, scratch mode: ffbs_fast.cuh (deferred normalisation, split barriers, state ring, WAYS pairs in flight) ----
template <typename Real, int K, int B, int MINB, int WAYS, int CPB = 1, int PHASE = 0>
cudaError_t ffbs_fast_do(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    auto kern = ffbs_fast_kernel<Real, K, B, MINB, WAYS, CPB, PHASE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return e;
    // the driver's default carve-out can stop short of what the registers allow: ask for all of it
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (query) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L.grid, kern, B * CPB, L.smem);
    kern<<<L.grid, B * CPB, L.smem, L.stream>>>(a);
    return cudaGetLastError();
}

#ifndef FAST_BWD_MINB
#define FAST_BWD_MINB 16  // resident warps per SM the backward-only kernel is compiled for (<= 128 registers)
#endif
// (threads per chain, ways, chains per CTA) the kernel is instantiated for
inline bool ffbs_fast_has_variant(int threads, int ways, int cpb) {
    if (threads == 32) return (ways == 2 && (cpb == 1 || cpb == 2 || cpb == 5 || cpb == 9)) || (cpb == 1 && (ways == 1 || ways == 3 || ways == 4));
    if (threads == 64) return cpb == 1 && (ways == 1 || ways == 2);
    return cpb == 1 && ways == 1 && (threads == 96 || threads == 128 || threads == 256);
}
template <typename Real, int K>
cudaError_t ffbs_fast_pick(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    constexpr bool D = sizeof(Real) == 8;
    const int w = a.fast_ways, cpb = a.fast_cpb;
    switch (L.threads) {
        case 32:
            if (w == 2 && cpb == 9) return ffbs_fast_do<Real, K, 32, 1, 2, 9>(a, L, query);   // 9 chains/SM, <= 224 registers
            if (w == 2 && cpb == 5) return ffbs_fast_do<Real, K, 32, 2, 2, 5>(a, L, query);   // 10 chains/SM, <= 200
            if (w == 2 && cpb == 2) return ffbs_fast_do<Real, K, 32, 5, 2, 2>(a, L, query);   // 10 chains/SM, <= 200
            if (w == 4) return ffbs_fast_do<Real, K, 32, 7, 4>(a, L, query);
            if (w == 3) return ffbs_fast_do<Real, K, 32, 8, 3>(a, L, query);
            // split kernels (FfbsArgs::fast_phase): forward as the fused kernel, backward compiled for twice the warps
            if (w == 2 && cpb == 1 && a.fast_phase == 1) return ffbs_fast_do<Real, K, 32, 8, 2, 1, 1>(a, L, query);
            if (w == 2 && cpb == 1 && a.fast_phase == 2) return ffbs_fast_do<Real, K, 32, FAST_BWD_MINB, 2, 1, 2>(a, L, query);
            if (w == 2) return ffbs_fast_do<Real, K, 32, 8, 2>(a, L, query);                  // 8 chains/SM, <= 256
            return ffbs_fast_do<Real, K, 32, D ? 16 : 32, 1>(a, L, query);
        case 64:
            if (w == 2) return ffbs_fast_do<Real, K, 64, 7, 2>(a, L, query);
            return ffbs_fast_do<Real, K, 64, D ? 10 : 16, 1>(a, L, query);
        case 96: return ffbs_fast_do<Real, K, 96, D ? 7 : 10, 1>(a, L, query);
        case 128: return ffbs_fast_do<Real, K, 128, D ? 5 : 8, 1>(a, L, query);
        case 256: return ffbs_fast_do<Real, K, 256, D ? 2 : 4, 1>(a, L, query);
    }
    return cudaErrorInvalidValue;
}
inline int ffbs_fast_round_threads(int t) { return t <= 32 ? 32 : t <= 64 ? 64 : t <= 96 ? 96 : t <= 128 ? 128 : 256; }

// resident warps per SM the fp32 kernel is compiled for.  Measured on the K = 10 shard (16384 chains), kernel ms:
// 12 (145 registers): 8.49; 14 (122, still 16 resident): 7.37; 16 (120): 7.55; 18 (112): 9.36; 20 (96): 10.84
#ifndef FASTK_MINB32
#define FASTK_MINB32 14
#endif
// ---- 4 <= K <= 10, scratch mode: ffbs_fastk.cuh (one warp per chain, lane = (grid point, source state)) ----
template <typename Real, int K, int WAYS>
cudaError_t ffbs_fastk_do(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    constexpr int MINB = sizeof(Real) == 8 ? 8 : FASTK_MINB32;  // <= 255 / <= 144 registers
    auto kern = ffbs_fastk_kernel<Real, K, WAYS, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (query) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L.grid, kern, 32, L.smem);
    kern<<<L.grid, 32, L.smem, L.stream>>>(a);
    return cudaGetLastError();
}
template <typename Real, int K>
cudaError_t ffbs_fastk_pick(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    if (a.fast_ways == 1) return ffbs_fastk_do<Real, K, 1>(a, L, query);
    return ffbs_fastk_do<Real, K, 2>(a, L, query);
}

// ---- K >= 4: item-parallel kernel, block size a multiple of K (two sizes per K) ----
constexpr int ffbs_item_base_threads(int K) {
    return K == 3 ? 96 : K == 4 ? 128 : K == 5 ? 160 : K == 6 ? 192 : K == 7 ? 224 : K == 8 ? 128 : K == 9 ? 288 : 160;
}
inline int ffbs_item_threads(int K, long long n_max) {
    const int b1 = ffbs_item_base_threads(K);
    return (n_max * K > 32LL * b1) ? 2 * b1 : b1;  // keep a full row within ~32 items per thread
}

template <typename Real, int K, int B>
cudaError_t ffbs_item_do(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    // register budget: ~64 per thread in fp32, ~128 in fp64
    constexpr int fit = 1024 / B < 1 ? 1 : 1024 / B;
    constexpr int MINB = sizeof(Real) == 8 ? (512 / B < 1 ? 1 : 512 / B) : fit;
    auto kern = ffbs_item_kernel<Real, K, B, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return e;
    if (query) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L.grid, kern, B, L.smem);
    kern<<<L.grid, B, L.smem, L.stream>>>(a);
    return cudaGetLastError();
}

template <typename Real, int K>
cudaError_t ffbs_item_pick(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    constexpr int B1 = ffbs_item_base_threads(K);
    if (L.threads == B1) return ffbs_item_do<Real, K, B1>(a, L, query);
    if (L.threads == 2 * B1) return ffbs_item_do<Real, K, 2 * B1>(a, L, query);
    return cudaErrorInvalidValue;
}

// K is split over two translation units per precision (compile time): lo = 2..5, hi = 6..10
template <typename Real>
cudaError_t ffbs_dispatch_lo(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    switch (L.K) {
        case 2: return a.use_fast ? ffbs_fast_pick<Real, 2>(a, L, query) : ffbs_pick<Real, 2>(a, L, query);
        case 3: return a.use_item ? ffbs_item_pick<Real, 3>(a, L, query)
                       : a.use_fast ? ffbs_fast_pick<Real, 3>(a, L, query) : ffbs_pick<Real, 3>(a, L, query);
        case 4: return a.use_fast ? ffbs_fastk_pick<Real, 4>(a, L, query) : ffbs_item_pick<Real, 4>(a, L, query);
        case 5: return a.use_fast ? ffbs_fastk_pick<Real, 5>(a, L, query) : ffbs_item_pick<Real, 5>(a, L, query);
    }
    return cudaErrorInvalidValue;
}

template <typename Real>
cudaError_t ffbs_dispatch_hi(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    switch (L.K) {
        case 6: return a.use_fast ? ffbs_fastk_pick<Real, 6>(a, L, query) : ffbs_item_pick<Real, 6>(a, L, query);
        case 7: return a.use_fast ? ffbs_fastk_pick<Real, 7>(a, L, query) : ffbs_item_pick<Real, 7>(a, L, query);
        case 8: return a.use_fast ? ffbs_fastk_pick<Real, 8>(a, L, query) : ffbs_item_pick<Real, 8>(a, L, query);
        case 9: return a.use_fast ? ffbs_fastk_pick<Real, 9>(a, L, query) : ffbs_item_pick<Real, 9>(a, L, query);
        case 10: return a.use_fast ? ffbs_fastk_pick<Real, 10>(a, L, query) : ffbs_item_pick<Real, 10>(a, L, query);
    }
    return cudaErrorInvalidValue;
}

cudaError_t ffbs_dispatch_f64_lo(const FfbsArgs& a, FfbsLaunch& L, bool query);
cudaError_t ffbs_dispatch_f64_hi(const FfbsArgs& a, FfbsLaunch& L, bool query);
cudaError_t ffbs_dispatch_f32_lo(const FfbsArgs& a, FfbsLaunch& L, bool query);
cudaError_t ffbs_dispatch_f32_hi(const FfbsArgs& a, FfbsLaunch& L, bool query);

}  // namespace smjp
