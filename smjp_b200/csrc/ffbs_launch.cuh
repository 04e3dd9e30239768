// Per-precision launcher of ffbs_kernel: picks the <K, MAXT, MINB> instantiation.  Included by
// ffbs_f64.cu / ffbs_f32.cu so that the two precisions compile in parallel.
#pragma once
#include "ffbs.cuh"

namespace smjp {

struct FfbsLaunch {
    int K;
    int threads;
    size_t smem;
    int grid;           // in: requested (0 = query only); out of query: resident CTAs per SM
    cudaStream_t stream;
};

#define SMJP_FOR_EACH_K(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10)

template <typename Real, int K, int MAXT, int MINB>
cudaError_t ffbs_do(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    auto kern = ffbs_kernel<Real, K, MAXT, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return e;
    if (query) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L.grid, kern, L.threads, L.smem);
    kern<<<L.grid, L.threads, L.smem, L.stream>>>(a);
    return cudaGetLastError();
}

template <typename Real, int K>
cudaError_t ffbs_pick(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    // resident CTAs per SM the register budget is tuned for: fp32 needs about half the registers of
    // fp64, K >= 6 about twice those of K <= 5 (K*K hazards per grid point are kept in flight)
    constexpr int S = (sizeof(Real) == 8 ? 1 : 2);
    constexpr int D = (K <= 5 ? 2 : 1);
    if (L.threads <= 128) return ffbs_do<Real, K, 128, 2 * S * D>(a, L, query);
    if (L.threads <= 256) return ffbs_do<Real, K, 256, S * D>(a, L, query);
    return ffbs_do<Real, K, 512, (S * D) / 2 < 1 ? 1 : (S * D) / 2>(a, L, query);
}

template <typename Real>
cudaError_t ffbs_dispatch(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    switch (L.K) {
#define X(k) case k: return ffbs_pick<Real, k>(a, L, query);
        SMJP_FOR_EACH_K(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

cudaError_t ffbs_dispatch_f64(const FfbsArgs& a, FfbsLaunch& L, bool query);
cudaError_t ffbs_dispatch_f32(const FfbsArgs& a, FfbsLaunch& L, bool query);

}  // namespace smjp
