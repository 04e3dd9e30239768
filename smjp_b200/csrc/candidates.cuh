// Candidate (thinned-event) sampler, prior trajectory sampler, observation simulator.
//
// Replaces (reference):
//   sample_smjp_event_times        pkg/smjp_utils.py:194-210
//   PoissonProcess.sample          pkg/smjp_utils.py:61-99   N ~ Poisson((Omega-1) H_vs(tau)) per target
//   PoissonProcess.weibull_mean    pkg/smjp_utils.py:135-145
//   WeibullDistribution.sample     pkg/distributions.py:292-306 (unit-scale draw)
//   sample_smjp_trajectory_prior   pkg/smjp_utils.py:216-254
//   sample_data_posterior          pkg/smjp_utils.py:350-364 + smjpEmission.sample :559-563
//
// Randomness is counter-based (Philox, common.cuh), so the count pass and the fill pass regenerate
// the same draws and no RNG state or per-sojourn buffers ever touch HBM.  oracle/smjp_oracle.py
// (candidates, prior_path, simulate_observations) is the bit-level CPU twin of this file.
#pragma once
#include "../../include/smjp_b200.h"
#include "common.cuh"

namespace smjp {

#define SMJP_POISSON_CHUNK 32.0

__device__ __forceinline__ int poisson_inv(double mu, double u) {
    double p = exp(-mu), cdf = p;
    int k = 0;
    while (u > cdf && k < 1000) {
        ++k;
        p *= mu / k;
        cdf += p;
    }
    return k;
}

struct CandArgs {
    int C;
    const int64_t* p_offs;  // path windows
    const int32_t* path_v;
    const double* path_t;
    const int32_t* path_len;
    const double* model;  // shape[K*K], c2[K*K], emis[K*Kobs], scale[K*K]
    int K, Kobs;
    double omega;
    uint64_t seed;
    uint32_t sweep, chain_base;
    int mode;
    int32_t* soff;    // per sojourn: offset of its first grid point inside the chain's W (same windows as path)
    int64_t* w_len;   // per chain grid length (count pass) / w_offs (fill pass)
    const int64_t* w_offs;
    double* W;
    // speculative fill (smjp_sweep_dev): guard -> {n_max, total} of this sweep on the device; a grid that would
    // not fit the caller's W (guard_total points) is not written
    const int64_t* guard;
    int64_t guard_total;
    int R;  // item slots per sojourn = the largest number of target CLASSES of a row (smjp_set_model; 1 when shape == 1)
};

// Thinned events of ONE (sojourn q, target class s) item of chain c.  target_count draws their number N (m Poisson
// chunks) and leaves tk = tau^k; target_emit writes the N positions (unsorted).  x^k is taken as 2^(k log2 x): two
// library calls of ~40 instructions instead of pow()'s ~250, within 1e-15 of it (the kernels are bound by exactly
// these fp64 chains).  A class of equal-shape targets has intensity (Omega-1) tau^k sum lambda^-k (csum).
__device__ __forceinline__ int target_count(const CandArgs& a, int c, int q, int v, int s, double l2tau, int& m,
                                            double& k, double& tk) {
    const int K = a.K;
    v = min(max(v, 0), K - 1);  // a state index outside [0, K) (caller's paths) never indexes the model out of bounds
    const double* shape = a.model;
    const double* csum = a.model + 4 * K * K + K * a.Kobs;  // class sum of lambda^-k at the representative (smjp_set_model)
    m = 1; k = 1.0; tk = 0.0;
    if (s < 0) return 0;  // this row has fewer classes than slots
    k = shape[v * K + s];
    tk = exp2(k * l2tau);
    const double mu = (a.omega - 1.0) * tk * csum[v * K + s];
    m = (int)ceil(mu / SMJP_POISSON_CHUNK);
    if (m < 1) m = 1;
    const uint32_t aa = (uint32_t)(q * K + s);
    int N = 0;
    for (int cc = 0; cc < m; ++cc)
        N += poisson_inv(mu / m, philox_u01(a.seed, TAG_CANDIDATES, (uint64_t)cc, aa, a.sweep,
                                            a.chain_base + (uint32_t)c));
    return N;
}
__device__ __forceinline__ void target_emit(const CandArgs& a, int c, int q, int s, int N, int m, double k, double tk,
                                            double tau, double t0, double* out) {
    const double F = (a.mode == SMJP_CAND_REFERENCE) ? -expm1(-tk) : 0.0;
    const double rk = 1.0 / k;
    const uint32_t aa = (uint32_t)(q * a.K + s);
    for (int d = 0; d < N; ++d) {
        const double u = philox_u01(a.seed, TAG_CANDIDATES, (uint64_t)(m + d), aa, a.sweep,
                                    a.chain_base + (uint32_t)c);
        const double x = (a.mode == SMJP_CAND_REFERENCE) ? exp2(rk * log2(-log1p(-u * F))) : tau * exp2(rk * log2(u));
        out[d] = t0 + fmin(x, tau);
    }
}

// One WARP per chain, one LANE per (sojourn, target) item -- the serial chain of a lane is one Poisson draw and its
// events, not K of them (a thread per sojourn spent its time in K back-to-back fp64 pow/exp chains).  Items are
// ordered sojourn-major, so the exclusive prefix of the counts + the number of earlier sojourns (one grid point each,
// the jump time) is the item's position in the chain's grid.
// A sojourn has R item slots, slot r = the r-th target CLASS of its row (replist, smjp_set_model): models whose targets
// share shapes (cfg4: shape == 1, one class per row) do not spend K lanes per sojourn on one Poisson draw.
struct CandItem {
    int q, r, s, v;  // sojourn, slot, target (class representative; -1: the row has fewer classes), state
    double t0, tau, l2tau;
    bool valid;
};
__device__ __forceinline__ CandItem cand_item(const CandArgs& a, int64_t p0, int nsoj, int i) {
    CandItem it;
    it.q = i / a.R;
    it.r = i - it.q * a.R;
    it.valid = it.q < nsoj;
    it.s = -1; it.v = 0; it.t0 = 0.0; it.tau = 0.0; it.l2tau = 0.0;
    if (it.valid) {
        const int K = a.K;
        it.v = min(max(a.path_v[p0 + it.q], 0), K - 1);  // a state index outside [0, K) never indexes the model out of bounds
        it.s = (int)a.model[5 * K * K + K * a.Kobs + it.v * K + it.r];
        it.t0 = a.path_t[p0 + it.q];
        it.tau = a.path_t[p0 + it.q + 1] - it.t0;
        it.l2tau = log2(it.tau);
    }
    return it;
}

__global__ void cand_count_kernel(const CandArgs a) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, NW = blockDim.x >> 5;
    for (int c = blockIdx.x * NW + wid; c < a.C; c += gridDim.x * NW) {
        const int64_t p0 = a.p_offs[c];
        const int plen = a.path_len[c];
        const int nsoj = plen - 1;
        const int items = nsoj * a.R;
        int carry = 0;
        for (int base = 0; base < items; base += 32) {
            const CandItem it = cand_item(a, p0, nsoj, base + lane);
            int m;
            double k, tk;
            const int cnt = it.valid ? target_count(a, c, it.q, it.v, it.s, it.l2tau, m, k, tk) : 0;
            const int incl = warp_inclusive_scan(cnt, lane);
            if (it.valid && it.r == 0) a.soff[p0 + it.q] = carry + incl - cnt + it.q;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) {
            if (plen >= 2) a.soff[p0 + nsoj] = carry + nsoj;                 // the end point's slot
            a.w_len[c] = (plen >= 2) ? (int64_t)carry + nsoj + 1 : 0;  // + the end point
        }
    }
}

// single CTA: exclusive scan of the chain grid lengths -> w_offs[C+1]; stats = {n_max, total}
__global__ void chain_scan_kernel(int C, const int64_t* len, int64_t* offs, int64_t* stats) {
    __shared__ long long s_warp[33];
    __shared__ long long s_carry;
    __shared__ long long s_max;
    const int tid = threadIdx.x, B = blockDim.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { s_carry = 0; s_max = 0; }
    __syncthreads();
    long long mymax = 0;
    for (int base = 0; base < C; base += B) {
        const int c = base + tid;
        long long v = c < C ? (long long)len[c] : 0;
        mymax = v > mymax ? v : mymax;
        long long incl = warp_inclusive_scan(v, lane);
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        long long off = s_carry;
        for (int w = 0; w < wid; ++w) off += s_warp[w];
        if (c < C) offs[c] = off + incl - v;
        __syncthreads();
        if (tid == B - 1) s_carry = off + incl;
        __syncthreads();
    }
    atomicMax(&s_max, mymax);
    __syncthreads();
    if (tid == 0) {
        offs[C] = s_carry;
        stats[0] = s_max - 1;  // n_max (intervals)
        stats[1] = s_carry;    // total grid points
    }
}

// warp per chain.  Phase 1, lane per (sojourn, target): W[soff[q]] = T[q], the item's events behind it, unsorted.
// Phase 2, lane per sojourn: insertion sort of the sojourn's few events (they came target by target).
__global__ void cand_fill_kernel(const CandArgs a) {
    if (a.guard && a.guard[1] > a.guard_total) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, NW = blockDim.x >> 5;
    for (int c = blockIdx.x * NW + wid; c < a.C; c += gridDim.x * NW) {
        const int64_t p0 = a.p_offs[c], w0 = a.w_offs[c];
        const int plen = a.path_len[c];
        const int nsoj = plen - 1;
        const int items = nsoj * a.R;
        double* Wc = a.W + w0;
        int carry = 0;
        for (int base = 0; base < items; base += 32) {
            const CandItem it = cand_item(a, p0, nsoj, base + lane);
            // the count is redrawn (same Philox counters) rather than kept in HBM; the events go straight to their slot
            int m = 1;
            double k = 1.0, tk = 0.0;
            const int cnt = it.valid ? target_count(a, c, it.q, it.v, it.s, it.l2tau, m, k, tk) : 0;
            const int incl = warp_inclusive_scan(cnt, lane);
            const int pos = carry + incl - cnt + it.q;  // slot of T[q] if s == 0; events of the item start at pos + 1
            if (it.valid) {
                if (it.r == 0) Wc[pos] = it.t0;
                if (cnt > 0) target_emit(a, c, it.q, it.s, cnt, m, k, tk, it.tau, it.t0, Wc + pos + 1);
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0 && plen >= 2) Wc[carry + nsoj] = a.path_t[p0 + plen - 1];
        __syncwarp();
        for (int q = lane; q < nsoj; q += 32) {
            const int lo = a.soff[p0 + q] + 1, hi = a.soff[p0 + q + 1];
            for (int i = lo + 1; i < hi; ++i) {
                const double val = Wc[i];
                int pos = i;
                while (pos > lo && Wc[pos - 1] > val) {
                    Wc[pos] = Wc[pos - 1];
                    --pos;
                }
                Wc[pos] = val;
            }
        }
    }
}

// Host paths in pinned (device-mapped) memory -> device staging, entries in use only: warp per chain, coalesced reads
// over PCIe.  (smjp_sweep_host: the windows of a sweep's output are grid-sized, the paths a fraction of that.)
__global__ void path_gather_kernel(int C, const int64_t* p_offs, const int32_t* path_len, const int32_t* src_v,
                                   const double* src_t, int32_t* dst_v, double* dst_t) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, NW = blockDim.x >> 5;
    for (int c = blockIdx.x * NW + wid; c < C; c += gridDim.x * NW) {
        const int64_t p0 = p_offs[c];
        const int plen = path_len[c];
        for (int q = lane; q < plen; q += 32) {
            dst_t[p0 + q] = src_t[p0 + q];
            dst_v[p0 + q] = src_v[p0 + q];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// prior trajectories: thread per chain (sequential competing risks)
// ------------------------------------------------------------------------------------------------
struct PriorArgs {
    int C;
    const int64_t* p_offs;  // capacity windows
    const double* model;
    const double* pi0;  // device [K]
    int K, Kobs;
    double t_end;
    int honour_scale;
    uint64_t seed;
    uint32_t chain_base;
    int32_t* path_v;
    double* path_t;
    int32_t* path_len;
    int32_t* status;
};

__global__ void prior_kernel(const PriorArgs a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.C) return;
    const int K = a.K;
    const double* shape = a.model;
    const double* scale = a.model + 2 * K * K + K * a.Kobs;
    const int64_t p0 = a.p_offs[c];
    const int cap = (int)(a.p_offs[c + 1] - p0);
    const uint32_t chain = a.chain_base + (uint32_t)c;
    // v0 ~ pi0 by inverse CDF on the normalised vector
    double tot = 0;
    for (int s = 0; s < K; ++s) tot += a.pi0[s];
    const double u0 = philox_u01(a.seed, TAG_PRIOR, 0, 0u, 0u, chain);
    int v = K - 1;
    {
        double cdf = 0;
        int lastpos = 0;
        bool found = false;
        for (int s = 0; s < K; ++s) {
            const double p = a.pi0[s] / tot;
            if (p > 0) lastpos = s;
            cdf += p;
            if (!found && cdf > u0) { v = s; found = true; }
        }
        if (!found) v = lastpos;
    }
    double w = 0.0;
    int len = 0, vprev = v;
    bool overflow = false;
    if (cap < 2) overflow = true;
    if (!overflow) {
        a.path_v[p0] = v;
        a.path_t[p0] = 0.0;
        len = 1;
    }
    int step = 0;
    while (!overflow && w < a.t_end) {
        double best = INFINITY;
        int arg = 0;
        for (int s = 0; s < K; ++s) {
            const double u = philox_u01(a.seed, TAG_PRIOR, (uint64_t)(1 + (int64_t)step * K + s), 0u, 0u, chain);
            double h = pow(-log1p(-u), 1.0 / shape[v * K + s]);
            if (a.honour_scale) h *= scale[v * K + s];
            if (h < best) { best = h; arg = s; }
        }
        w += best;
        vprev = v;
        v = arg;
        if (len >= cap) { overflow = true; break; }
        a.path_v[p0 + len] = v;
        a.path_t[p0 + len] = w;
        ++len;
        ++step;
    }
    if (overflow) {
        a.status[c] = SMJP_ST_OVERFLOW;
        a.path_len[c] = 0;
        return;
    }
    // smjp_utils.py:251-252: last entry becomes (v_prev, t_end)
    a.path_v[p0 + len - 1] = vprev;
    a.path_t[p0 + len - 1] = a.t_end;
    a.path_len[c] = len;
    a.status[c] = 0;
}

// ------------------------------------------------------------------------------------------------
// observations: thread per (chain, observation)
// ------------------------------------------------------------------------------------------------
struct ObsSimArgs {
    int C;
    const int64_t* p_offs;
    const int32_t* path_v;
    const double* path_t;
    const int32_t* path_len;
    const int64_t* obs_offs;
    const double* obs_t;
    int32_t* obs_y;
    const double* model;
    int K, Kobs;
    uint64_t seed;
    uint32_t chain_base;
};

__global__ void obs_sim_kernel(const ObsSimArgs a) {
    const int K = a.K, Kobs = a.Kobs;
    const double* emis = a.model + 2 * K * K;
    for (int c = blockIdx.x; c < a.C; c += gridDim.x) {
        const int64_t p0 = a.p_offs[c], o0 = a.obs_offs[c];
        const int plen = a.path_len[c];
        const int D = (int)(a.obs_offs[c + 1] - o0);
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const double t = a.obs_t[o0 + d];
            int lo = 0, hi = plen;  // first q with T[q] > t
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (a.path_t[p0 + mid] <= t) lo = mid + 1; else hi = mid;
            }
            int q = lo - 1;
            q = q < 0 ? 0 : (q > plen - 2 ? plen - 2 : q);
            const int v = a.path_v[p0 + q];
            const double u = philox_u01(a.seed, TAG_DATA, (uint64_t)d, 0u, 0u, a.chain_base + (uint32_t)c);
            double cdf = 0;
            int y = -1, lastpos = 0;
            for (int x = 0; x < Kobs; ++x) {
                const double p = emis[v * Kobs + x];
                if (p > 0) lastpos = x;
                cdf += p;
                if (y < 0 && cdf > u) y = x;
            }
            a.obs_y[o0 + d] = y >= 0 ? y : lastpos;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Longest-first processing order (the persistent CTAs pull chains from a queue; work per chain is
// ~n^2, so starting the long grids first trims the tail of the launch).  Counting sort by n into
// 1024 bins, single CTA.
// ------------------------------------------------------------------------------------------------
__global__ void chain_order_kernel(int C, const int64_t* w_offs, int n_max, int32_t* order) {
    __shared__ int hist[1024];
    __shared__ int start[1024];
    const int tid = threadIdx.x;
    hist[tid] = 0;
    __syncthreads();
    const long long denom = n_max > 0 ? n_max : 1;
    for (int c = tid; c < C; c += 1024) {
        const long long n = w_offs[c + 1] - w_offs[c] - 1;
        int b = (int)((n < 0 ? 0 : n) * 1023 / denom);
        b = b > 1023 ? 1023 : b;
        atomicAdd(&hist[1023 - b], 1);  // bin 0 = longest
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int b = 0; b < 1024; ++b) {
            start[b] = acc;
            acc += hist[b];
        }
    }
    __syncthreads();
    for (int c = tid; c < C; c += 1024) {
        const long long n = w_offs[c + 1] - w_offs[c] - 1;
        int b = (int)((n < 0 ? 0 : n) * 1023 / denom);
        b = b > 1023 ? 1023 : b;
        order[atomicAdd(&start[1023 - b], 1)] = c;
    }
}

// chains whose fp32 pass ended without a path (row sum under/overflow in single precision): list for the fp64 rescue
__global__ void collect_failed_kernel(int C, const int32_t* status, int32_t* list, int32_t* count) {
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        if (status[c] & SMJP_ST_DEGENERATE) list[atomicAdd(&s_n, 1)] = c;
    __syncthreads();
    if (threadIdx.x == 0) *count = s_n;
}

}  // namespace smjp
