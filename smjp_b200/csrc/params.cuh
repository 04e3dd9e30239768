// Parameter step of the Gibbs sampler: per-transition sufficient statistics and the
// Weibull shape estimate of the reference, on the device.
//
// Replaces (reference, paths relative to its root):
//   filter_T_by_state            pkg/raoteh.py:136-151   holding times T[j+1]-T[j] of the consecutive path
//                                                         entries (V[j], V[j+1]) == (v, v')
//   esimtate_weibull_shape       pkg/raoteh.py:153-167   grid search over k = grid, 2 grid, ... < grid_max of
//                                                         g(k) = sum T^k ln T / sum T^k - 1/k - mean(ln T),
//                                                         returning the FIRST k with the smallest |g(k)|
//   esimtate_weibull_shape_mat   pkg/raoteh.py:119-134   the K x K matrix of those (empty lists are left to the
//                                                         host, which draws U(0.6, 3) as the reference does)
//
// g is strictly increasing in k (d/dk of the first term is a weighted variance of ln T, and -1/k increases),
// so the grid point with the smallest |g| is one of the two that bracket the root: the kernel finds the
// bracket by bisection over the grid index (13 evaluations instead of 4999) and compares the two |g| values,
// the lower index winning ties -- the reference's argmin.  One CTA per (group, v, v'); a group is one chain,
// or all chains pooled (the batched sampler shares its parameters across chains).  Sums run over the path
// entries in a fixed order and reduce in a fixed tree, so results do not depend on scheduling.
#pragma once
#include "../../include/smjp_b200.h"
#include "common.cuh"

namespace smjp {

struct ShapeMleArgs {
    int C, K, pooled;
    const int64_t* p_offs;
    const int32_t* path_v;
    const double* path_t;
    const int32_t* path_len;
    double grid, grid_max;
    double* khat;     // [G, K, K]   NaN when the transition never occurs
    double* gval;     // [G, K, K]   g(khat)
    int32_t* count;   // [G, K, K]   number of holding times
    double* sum_log;  // [G, K, K]   sum ln T
    double* sum_t;    // [G, K, K]   sum T
};

constexpr int MLE_THREADS = 256;

// block-wide sum of two doubles in a fixed order; every thread receives the totals
__device__ __forceinline__ void block_sum2(double& x, double& y, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    x = warp_sum(x);
    y = warp_sum(y);
    __syncthreads();
    if (lane == 0) { red[2 * wid] = x; red[2 * wid + 1] = y; }
    __syncthreads();
    double sx = 0, sy = 0;
    for (int w = 0; w < MLE_THREADS / 32; ++w) { sx += red[2 * w]; sy += red[2 * w + 1]; }
    x = sx;
    y = sy;
}

__global__ void __launch_bounds__(MLE_THREADS) shape_mle_kernel(const ShapeMleArgs a) {
    __shared__ double red[2 * MLE_THREADS / 32];
    const int K = a.K;
    const int g = blockIdx.x / (K * K), pair = blockIdx.x - g * K * K;
    const int v0 = pair / K, v1 = pair - v0 * K;
    const int c_lo = a.pooled ? 0 : g, c_hi = a.pooled ? a.C : g + 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, NW = MLE_THREADS / 32;

    // visit(f): f(ln T, T) for every holding time of the transition v0 -> v1, warps over chains, lanes over entries
    auto visit = [&](auto&& f) {
        for (int c = c_lo + wid; c < c_hi; c += NW) {
            const int64_t p0 = a.p_offs[c];
            const int len = a.path_len[c];
            for (int p = lane; p + 1 < len; p += 32) {
                if (a.path_v[p0 + p] == v0 && a.path_v[p0 + p + 1] == v1) {
                    const double T = a.path_t[p0 + p + 1] - a.path_t[p0 + p];
                    f(log(T), T);
                }
            }
        }
    };
    double cnt = 0, slog = 0;
    visit([&](double lt, double) { cnt += 1.0; slog += lt; });
    double st = 0, dummy = 0;
    visit([&](double, double T) { st += T; });
    block_sum2(cnt, slog, red);
    block_sum2(st, dummy, red);
    const int64_t o = blockIdx.x;
    if (threadIdx.x == 0) {
        a.count[o] = (int)cnt;
        a.sum_log[o] = slog;
        a.sum_t[o] = st;
    }
    if (cnt == 0) {
        if (threadIdx.x == 0) { a.khat[o] = nan(""); a.gval[o] = nan(""); }
        return;
    }
    const double mean_log = slog / cnt;
    // numpy.arange(grid, grid_max, grid): M = ceil((grid_max - grid) / grid) points k_m = grid + m * grid
    const int M = (int)ceil((a.grid_max - a.grid) / a.grid);
    auto gfun = [&](int m) {
        const double k = a.grid + (double)m * a.grid;
        double s0 = 0, s1 = 0;
        visit([&](double lt, double) {
            const double tk = exp(k * lt);  // T^k
            s0 += tk;
            s1 += tk * lt;
        });
        block_sum2(s0, s1, red);
        return s1 / s0 - 1.0 / k - mean_log;
    };
    // first grid index whose g is >= 0 (M when there is none)
    int lo = 0, hi = M;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (gfun(mid) >= 0.0) hi = mid; else lo = mid + 1;
    }
    int best = lo < M ? lo : M - 1;
    double gb = gfun(best);
    if (lo > 0 && lo < M) {
        const double gl = gfun(lo - 1);
        if (fabs(gl) <= fabs(gb)) { best = lo - 1; gb = gl; }
    }
    if (threadIdx.x == 0) {
        a.khat[o] = a.grid + (double)best * a.grid;
        a.gval[o] = gb;
    }
}

}  // namespace smjp
