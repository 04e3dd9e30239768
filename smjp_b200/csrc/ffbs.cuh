// FFBS over the augmented (v, l) state space of a semi-Markov jump process: one CTA per chain.
//
// Replaces (reference, paths relative to its root):
//   likelihood_and_decoding_hmm   pkg/fast_hmm.py:202-335   forward filter, rows normalised
//   compute_transition_matrix     pkg/fast_hmm.py:343-354   never materialised here
//   smjpTransitionFunction        pkg/smjp_utils.py:382-416 jump A_vv'/B_v, thin 1 - A_v/B_v = 1 - 1/Omega
//   smjpEmission.__call__         pkg/smjp_utils.py:506-550 + PoissonProcess.likelihood :104-133
//   backward_sampling_hmm         pkg/fast_hmm.py:104-191   inverse CDF in state-major order
//   smjp_ffbs tail                pkg/smjp_utils.py:319-330 path compaction (+ :338-348 endpoint)
//   compute_evaluation_chain_metrics pkg/mcmc_utils.py:40-56
//
// Arithmetic: with tau = W[i+1] - W[j],
//   E_vs = (1/tau) (tau/lambda_vs)^k_vs = 2^((k_vs - 1) log2 tau - k_vs log2 lambda_vs)     [= H_vs / tau]
//   A_v = sum_s k_vs E_vs (hazard of leaving v),   H_v = tau sum_s E_vs (cumulative hazard)
//   e_i(v,j)      = obs_i(v) * [Omega A_v]^{not final} * exp(-Omega (H_v(tau) - H_v(tau_prev)))
//   alpha_i(v,j)  ~ e_i(v,j) (1 - 1/Omega) alpha_{i-1}(v,j)                      j < i  (thinned)
//   alpha_i(v',i) ~ e_i(v',i) J_i(v'),  J_i(v') = sum_{v,j<i} alpha_{i-1}(v,j) A_vv' / (Omega A_v)
// The Omega A_v of the emission cancels the 1/(Omega A_v) of the jump term, so the forward pass needs
// no division at all: with g = obs_i(v) surv base, alpha~_i = g Omega A_v and the contribution to
// J_{i+1}(s) is g k_vs E_vs.  The K*K powers computed for (i, j) serve e_i(v,j) AND J_{i+1}: B_i is
// generated in-kernel and never stored.  State carried per live (v,j): the unnormalised message and
// H_v(tau) -- both in shared memory, indexed so that consecutive threads hit consecutive banks.
// Normalisation is deferred by one step (the next step multiplies by 1/Z_i), so each grid step costs exactly
// one __syncthreads().  Message rows: exported on request as the reference's normalised alpha (scaled and
// streamed one step late); otherwise per-CTA scratch rows hold g = alpha / (Omega A_v), which is what the
// backward pass needs (one hazard per entry instead of K).  Exact live window: grid points whose states have
// underflowed to exact zeros are skipped in the forward pass, the row stores and the backward row scans
// (see "Live window" in the kernel); window_eps > 0 turns that into threshold pruning (opt-in).
// Siblings: ffbs_item.cuh (4 <= K <= 10, thread per state), ffbs_big.cuh (K <= 64, run-time K),
// ffbs_warp.cuh (K <= 3, one warp per chain, window in a shared-memory ring).
#pragma once
#include "../../include/smjp_b200.h"
#include <type_traits>

#include "common.cuh"

namespace smjp {

// shared-memory byte offsets of ffbs_kernel (see ffbs_layout)
struct FfbsLayout {
    int wd, ubuf, scan, sta, sth, red, par, jlo, total;
};

struct FfbsArgs {
    int C;
    const int64_t* w_offs;
    const double* W;
    const int64_t* obs_offs;
    const double* obs_t;
    const int32_t* obs_y;
    const double* uniforms;  // may be null -> Philox
    uint64_t seed;
    uint32_t sweep;
    uint32_t chain_base;  // global index of chain 0 (multi-GPU sharding keeps streams distinct)
    void* messages;       // export buffer or null
    const int64_t* msg_offs;
    void* msg_scratch;  // per-CTA message scratch when not exporting
    int64_t msg_scratch_stride;
    void* obsf_scratch;  // per-CTA observation factors [ncap][K]
    int64_t obsf_stride; // ffbs_warp.cuh: elements per warp region of obsf_scratch
    double* logZ;
    int32_t* aug_idx;
    int32_t* path_v;
    double* path_t;
    int32_t* path_len;
    double* time_in_state;
    int32_t* jumps;
    int32_t* status;
    const int32_t* order;  // optional chain processing order (longest first)
    const int32_t* C_dev;  // optional: number of entries of `order` to process, read on the device
    unsigned long long* live_pairs;  // optional device counter: grid pairs (i, j) actually processed (live window)
    FfbsLayout lay;                  // shared-memory byte offsets of ffbs_kernel (ffbs_layout)
    double window_eps;               // 0: exact live window (only exact zeros are dropped); else relative threshold
    unsigned int* queue;   // work counter, zeroed before launch
    const double* model;   // device: shape[K*K], c2[K*K] (= k*log2(lambda)), emis[K*Kobs], scale[K*K]
    int Kobs;
    int K;                    // run-time K (ffbs_big.cuh only; the other kernels take it as a template parameter)
    const void* big_params;   // ffbs_big.cuh: [3][K*K] reals = 1/k, (k-1) SCALE, (k log2 lambda - log2 k) SCALE
    // ffbs.cuh (K <= 3): the same three parameter sets as kernel arguments, in both precisions, so that the
    // forward pass takes them as constant-bank operands instead of shared-memory loads ([3][9])
    double pc_d[3][9];
    float pc_f[3][9];
    int use_item;     // K == 3 only: run the item-parallel kernel instead of the thread-per-grid-point one
    int use_fast;     // K <= 3, scratch mode: ffbs_fast.cuh (nslot = ring slots, lay = ffbs_fast_layout)
    int fast_ways;    // its grid pairs in flight per thread (1, 2 or 4)
    int fast_cpb;     // its chains per CTA (one warp each when > 1)
    int fast_phase;   // 0: forward + backward in one kernel; 1 / 2: forward-only / backward-only launch (scratch per chain)
    int32_t* fwd_ok;  // split kernels: [C] set by the forward kernel for chains whose backward pass may run
    unsigned int* queue2;  // work counter of the backward kernel
    int obs_product;  // 0: observations sharing a grid interval are SUMMED (reference); 1: multiplied
    double omega, power, t_end;
    // speculative launch (smjp_sweep_dev): the kernel was sized for guard_nmax intervals and guard_total grid points
    // before this sweep's counts reached the host; guard -> {n_max, total} on the device, and a launch whose
    // geometry turns out too small returns at once (the host relaunches it with the real numbers)
    const int64_t* guard;
    int64_t guard_nmax, guard_total;
    int ncap;   // max grid intervals any chain of this launch has (smem carve-up)
    int nslot;  // grid slots per thread: ceil((ncap + 1) / blockDim.x)
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) & ~(a - 1); }  // a: power of two

// grid slots per thread: ceil((ncap + 1) / threads)
__host__ __device__ inline int ffbs_nslot(int ncap, int threads) { return (ncap + threads) / threads; }

// dynamic shared memory layout of ffbs_kernel for (Real, K, ncap, threads): byte offsets, computed on the host
// and passed as kernel arguments (constant-bank operands; carving the pointers in the kernel cost ~40
// rematerialised instructions per grid pair at 64 registers)
template <typename Real>
__host__ __device__ inline FfbsLayout ffbs_layout(int K, int ncap, int threads) {
    const size_t state = (size_t)ffbs_nslot(ncap, threads) * K * threads;
    FfbsLayout L;
    size_t b = 0;
    L.wd = (int)b;   b += align_up((size_t)(ncap + 1) * sizeof(double), 16);                  // Wd
    L.ubuf = (int)b; b += align_up((size_t)threads * sizeof(double), 16);                     // buffered backward uniforms
    L.scan = (int)b; b += align_up((size_t)(threads / 32 + 1) * sizeof(double), 16);          // scan scratch
    L.sta = (int)b;  b += align_up(state * sizeof(Real), 16);                                 // st_a / wbuf
    L.sth = (int)b;  b += align_up(state * sizeof(Real), 16);                                 // st_h / sojourn list
    L.red = (int)b;  b += align_up((size_t)2 * (threads / 32) * (K + 1) * sizeof(Real), 16);  // reduction scratch
    L.par = (int)b;  b += align_up((size_t)(4 * K * K) * sizeof(Real), 16);                   // model
    L.jlo = (int)b;  b += align_up((size_t)(ncap + 1) * sizeof(int), 16);                     // live-window start of every row
    L.total = (int)b;
    return L;
}
template <typename Real>
__host__ __device__ inline size_t ffbs_smem_bytes(int K, int ncap, int threads) {
    return (size_t)ffbs_layout<Real>(K, ncap, threads).total;
}

// item-parallel kernel (ffbs_item.cuh): a slot holds threads/K grid points for all K states
__host__ __device__ inline int ffbs_item_nslot(int K, int ncap, int threads) {
    const int jps = threads / K;
    return (ncap + jps) / jps;
}

template <typename Real>
__host__ __device__ inline size_t ffbs_item_smem_bytes(int K, int ncap, int threads) {
    const size_t state = (size_t)ffbs_item_nslot(K, ncap, threads) * threads;
    size_t b = 0;
    b += align_up((size_t)(ncap + 1) * sizeof(double), 16);
    b += align_up((size_t)threads * sizeof(double), 16);
    b += align_up((size_t)(threads / 32 + 1) * sizeof(double), 16);
    b += align_up(state * sizeof(Real), 16);
    b += align_up(state * sizeof(Real), 16);
    b += align_up((size_t)2 * (threads / 32) * (K + 1) * sizeof(Real), 16);
    b += align_up((size_t)(4 * K * K) * sizeof(Real), 16);
    b += align_up((size_t)(ncap + 1) * sizeof(int), 16);  // live-window start of every row
    return b;
}

__device__ __forceinline__ bool np_isclose(double a, double b) {
    // numpy.isclose(a, b): |a-b| <= atol + rtol*|b|, rtol=1e-5, atol=1e-8
    return fabs(a - b) <= 1e-8 + 1e-5 * fabs(b);
}

// ------------------------------------------------------------------------------------------------
// 2^x and log2 x per precision.  The fp64 kernel is bound by the FP64 pipe (64 lanes/clk/SM), so
// both are table driven to minimise FP64-pipe instructions (DSETP also issues there):
//   2^x   : 2^(i/256) table + degree-4 polynomial on |r| <= 1/512 (truncation 4e-17), 8 DP ops,
//           under/overflow resolved with integer compares on the exponent;
//   log2 x: 128-entry {1/c, log2 c} table + degree-6 polynomial on |r| <= 1/256, 8 DP ops.
// fp32: one SFU op each.
// ------------------------------------------------------------------------------------------------
template <typename Real>
struct Exp2;

// polynomial coefficients live in constant memory so that DFMA takes them as c[bank] operands
__constant__ double c_exp2[4] = {6.931471805599453e-01, 2.402265069591007e-01, 5.550410866482158e-02,
                                 9.618129107628477e-03};  // ln2^i / i!
// the same, for an argument pre-multiplied by 256 (Exp2<double>::SCALE): ln2^i / (i! 256^i), exact rescalings
__constant__ double c_exp2s[4] = {6.931471805599453e-01 / 256.0, 2.402265069591007e-01 / 65536.0,
                                  5.550410866482158e-02 / 16777216.0, 9.618129107628477e-03 / 4294967296.0};
__constant__ double c_log2[6] = {1.442695040888963e+00, -7.213475204444817e-01, 4.808983469629878e-01,
                                 -3.606737602222409e-01, 2.885390081777927e-01, -2.404491734814939e-01};  // (-1)^(i+1)/(i ln2)

// entry i of the 2^(i/256) table in the form Exp2<double>::core wants it: (i << 12) subtracted from the high word
__device__ __forceinline__ double exp2_table_entry(int i) {
    const double v = ::exp2((double)i / 256.0);
    return __hiloint2double(__double2hiint(v) - (int)((unsigned)i << 12), __double2loint(v));
}

template <>
struct Exp2<double> {
    const double* tab;    // [256] 2^(i/256), high words offset (exp2_table_entry)
    const double2* ltab;  // [128] {1/c_i, -log2(1/c_i)}, c_i = 1 + (i + 1/2)/128
    static constexpr int TAB = 256;
    static constexpr int LTAB = 128;
    // Arguments arrive PRE-SCALED by SCALE = 256 (the callers fold it into their constants): xs = 256 x.
    // 2^x for |xs| < 2^31: n = round(xs) and r = xs - n come from the add-a-magic-constant trick (three DADD:
    // the FP64 pipe has headroom, the F2I/I2F conversions it replaces are long-latency XU instructions -- 1.4 %
    // of the cfg2 sweep); the power of two is added to the exponent field
    // of the TABLE value before the last fma, so the result needs no repacking; n is clamped to
    // the normal range with two integer min/max (results saturate at 2^-1022 / 2^1023 instead of 0 / inf).
    static constexpr double SCALE = 256.0;
    static constexpr double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: the low word of xs + MAGIC is round(xs)
    template <bool CLAMP>
    __device__ __forceinline__ double core(double xs) const {
        const double t = xs + MAGIC;
        const int n = __double2loint(t);
        const double r = xs - (t - MAGIC);
        double q = fma(r, c_exp2s[3], c_exp2s[2]);
        q = fma(r, q, c_exp2s[1]);
        q = fma(r, q, c_exp2s[0]);
        // 2^(n/256) = 2^(n >> 8) * tab[n & 255]: the table stores each entry with (i << 12) taken off its high word
        // (exp2_table_entry), so that adding n << 12 -- the exponent in bits 20.., the index back in bits 12..19 --
        // gives the scaled value with one shift-add and no separate n >> 8
        const int nc = CLAMP ? max(min(n, 1023 * 256 + 255), -1022 * 256) : n;
        const double tv = tab[nc & 255];
        const double ts = __hiloint2double(__double2hiint(tv) + (int)((unsigned)nc << 12), __double2loint(tv));
        return fma(ts, r * q, ts);
    }
    __device__ __forceinline__ double operator()(double xs) const { return core<true>(xs); }
    // any xs <= 0 including -inf (survival factors): exact 0 below 2^-1022.  No clamp is needed: whatever
    // the core makes of an argument below the normal range (a wrapped exponent, NaN) is replaced by the select
    __device__ __forceinline__ double neg(double xs) const {
        const double y = core<false>(xs);
        return xs < -1022.0 * SCALE ? 0.0 : y;
    }
    __device__ __forceinline__ double log2(double x) const {
        const int hi = __double2hiint(x);
        if (hi < 0x00100000 || hi >= 0x7ff00000) return ::log2(x);  // zero / denormal / negative / inf / NaN
        const double2 tv = ltab[(hi >> 13) & 127];
        const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
        const double r = fma(m, tv.x, -1.0);  // |r| <= 2^-8
        double p = fma(r, c_log2[5], c_log2[4]);
        p = fma(r, p, c_log2[3]);
        p = fma(r, p, c_log2[2]);
        p = fma(r, p, c_log2[1]);
        p = fma(r, p, c_log2[0]);
        return fma(r, p, (double)((hi >> 20) - 1023) + tv.y);
    }
};

template <>
struct Exp2<float> {
    const double* tab;
    const double2* ltab;
    static constexpr int TAB = 0;
    static constexpr int LTAB = 0;
    static constexpr float SCALE = 1.0f;
    __device__ __forceinline__ float operator()(float x) const { return Math<float>::exp2(x); }
    __device__ __forceinline__ float neg(float x) const { return Math<float>::exp2(x); }
    __device__ __forceinline__ float log2(float x) const { return Math<float>::log2(x); }
};

// N consecutive reals from shared memory; 128-bit (or 64-bit) loads when N allows (p must be N*sizeof(Real) aligned)
template <typename Real, int SPL>
__device__ __forceinline__ void lds_vec(const Real* p, Real (&o)[SPL]) {
    if constexpr (sizeof(Real) == 4 && SPL == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        o[0] = t.x; o[1] = t.y;
    } else if constexpr (sizeof(Real) == 4 && SPL == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else if constexpr (sizeof(Real) == 8 && (SPL == 2 || SPL == 4)) {
#pragma unroll
        for (int k = 0; k < SPL; k += 2) {
            const double2 t = *reinterpret_cast<const double2*>(p + k);
            o[k] = t.x; o[k + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SPL; ++k) o[k] = p[k];
    }
}

// Warp reduction of N (a power of two <= 16) values at once.  Plain butterflies cost 5*N shuffles; here
// the value index is folded into the lane index: at lane distance D each lane keeps one half of its
// values and trades the other half, so the live count halves per level (N/2 + N/4 + ... shuffles), and
// the single remaining value finishes with ordinary butterflies.  On return lane L holds, in v[0], the
// warp total of value number (L >> (5 - log2 N)).
template <typename T, int N, int D>
struct WarpHalve {
    static __device__ __forceinline__ void run(T* v, int lane) {
        constexpr int H = N / 2;
        const bool up = (lane & D) != 0;
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const T send = up ? v[k] : v[k + H];
            const T keep = up ? v[k + H] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, D);
        }
        WarpHalve<T, H, D / 2>::run(v, lane);
    }
};
template <typename T, int D>
struct WarpHalve<T, 1, D> {
    static __device__ __forceinline__ void run(T* v, int) {
#pragma unroll
        for (int d = D; d >= 1; d >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], d);
    }
};
template <typename T>
struct WarpHalve<T, 1, 0> {
    static __device__ __forceinline__ void run(T*, int) {}
};
constexpr int pow2_ceil(int x) { return x <= 1 ? 1 : x <= 2 ? 2 : x <= 4 ? 4 : x <= 8 ? 8 : 16; }
constexpr int log2_of(int x) { return x <= 1 ? 0 : x <= 2 ? 1 : x <= 4 ? 2 : x <= 8 ? 3 : 4; }

template <typename Real> struct ScanAcc { typedef double type; };
template <> struct ScanAcc<float> { typedef float type; };

// hazard parameter set w (0: 1/k, 1: (k-1) SCALE, 2: (k log2 lambda - log2 k) SCALE), entry t -- a constant-bank
// operand when t is a compile-time constant (the unrolled forward loops)
#define PC(w, t) (sizeof(Real) == 8 ? (Real)a.pc_d[w][t] : (Real)a.pc_f[w][t])

template <typename Real, int K, int B, int MINB>
__global__ void __launch_bounds__(B, MINB) ffbs_kernel(const FfbsArgs a) {
    using M = Math<Real>;
    using Acc = typename ScanAcc<Real>::type;
    constexpr int NW = B / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.guard && (a.guard[0] > a.guard_nmax || a.guard[1] > a.guard_total)) return;  // mis-speculated geometry
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ncap = a.ncap, Kobs = a.Kobs;
    const size_t state = (size_t)a.nslot * K * B;

    double* Wd = (double*)(smem_raw + a.lay.wd);
    double* ubuf = (double*)(smem_raw + a.lay.ubuf);  // backward uniforms of steps (u_hi - B, u_hi]
    Acc* scan_s = (Acc*)(smem_raw + a.lay.scan);      // [NW + 1] warp totals
    Real* st_a = (Real*)(smem_raw + a.lay.sta);  // forward: unnormalised message [slot][v][B]; backward: weights [K][len]
    Real* st_h = (Real*)(smem_raw + a.lay.sth);  // forward: H_v(tau) [slot][v][B]; backward: sojourn list (int pairs)
    Real* red = (Real*)(smem_raw + a.lay.red);
    Real* srk = (Real*)(smem_raw + a.lay.par);   // 1 / k_vs
    Real* skm1 = srk + K * K;    // (k_vs - 1) * SCALE
    Real* sc2p = skm1 + K * K;   // (k_vs log2(lambda_vs) - log2(k_vs)) * SCALE: 2^(((k-1) L - c2p)/SCALE) = k E = A_vs
    Real* sk = sc2p + K * K;     // k_vs
    int* jlo_arr = (int*)(smem_raw + a.lay.jlo);  // [ncap + 1] first live grid point of every row (scratch mode)
    int* soj = (int*)st_h;       // [2][ncap+1]: state, grid index of each sojourn (backward pass)

    __shared__ int s_chain;
    __shared__ int s_sel, s_selv;
    __shared__ int s_found[2];  // an owner claimed the draw; double-buffered by draw parity (reset one draw ahead)
    __shared__ int s_jmin[2 * NW];  // per-warp smallest live grid point of the oldest slot, double-buffered
    __shared__ int s_flag;
    __shared__ double s_tab[Exp2<Real>::TAB ? Exp2<Real>::TAB : 1];
    __shared__ double2 s_ltab[Exp2<Real>::LTAB ? Exp2<Real>::LTAB : 1];

    for (int t = tid; t < K * K; t += B) {
        const double k = a.model[t];
        sk[t] = (Real)k;
        srk[t] = (Real)(1.0 / k);
        skm1[t] = (Real)((k - 1.0) * (double)Exp2<Real>::SCALE);  // exponent arguments carry Exp2's scale
        sc2p[t] = (Real)((a.model[K * K + t] - ::log2(k)) * (double)Exp2<Real>::SCALE);
    }
    for (int t = tid; t < Exp2<Real>::TAB; t += B) s_tab[t] = exp2_table_entry(t);
    for (int t = tid; t < Exp2<Real>::LTAB; t += B) {
        const double rc = 1.0 / (1.0 + ((double)t + 0.5) / 128.0);
        s_ltab[t] = make_double2(rc, -::log2(rc));
    }
    Exp2<Real> ex2;
    ex2.tab = s_tab;
    ex2.ltab = s_ltab;
    const double* emis = a.model + 2 * K * K;

    const Real thin = (Real)(1.0 - 1.0 / a.omega);
    const Real omega_r = (Real)a.omega;
    const Real om_l2e = (Real)(a.omega * SMJP_LOG2E * (double)Exp2<Real>::SCALE);
    const Real power = (Real)a.power;
    Real* obsf = (Real*)a.obsf_scratch + (int64_t)blockIdx.x * ncap * K;

    while (true) {
        __syncthreads();
        if (tid == 0) {
            unsigned int q = atomicAdd(a.queue, 1u);
            const unsigned limit = a.C_dev ? (unsigned)*a.C_dev : (unsigned)a.C;  // device-side count: fp64 rescue list
            s_chain = (q < limit) ? (a.order ? a.order[q] : (int)q) : -1;
            s_flag = 0;
        }
        __syncthreads();
        const int c = s_chain;
        if (c < 0) break;
        const int64_t w0 = a.w_offs[c];
        const int n = (int)(a.w_offs[c + 1] - w0) - 1;
        if (n < 1 || n > ncap) {
            if (tid == 0) {
                a.status[c] = (n < 1) ? SMJP_ST_GRID_NOT_INCREASING : SMJP_ST_OVERFLOW;
                a.path_len[c] = 0;
            }
            continue;
        }
        // ---- stage the grid, validate it (smjp_utils.py:398-399) ----
        for (int t = tid; t <= n; t += B) Wd[t] = a.W[w0 + t];
        __syncthreads();
        {
            int bad = 0;
            for (int t = tid; t < n; t += B) bad |= !(Wd[t + 1] > Wd[t]);
            if (bad) atomicOr(&s_flag, 1);
        }
        // ---- observation factor per interval: (sum_{x in [W[i],W[i+1]]} P(x|v))^power ----
        {
            const int64_t o0 = a.obs_offs[c];
            const int D = (int)(a.obs_offs[c + 1] - o0);
            const double* ot = a.obs_t + o0;
            const int32_t* oy = a.obs_y + o0;
            for (int i = tid; i < n; i += B) {
                const double wc = Wd[i], wn = Wd[i + 1];
                int lo = 0, hi = D;  // first d with ot[d] >= wc
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (ot[mid] < wc) lo = mid + 1; else hi = mid;
                }
                Real f[K];
#pragma unroll
                for (int v = 0; v < K; ++v) f[v] = a.obs_product ? (Real)1 : (Real)0;
                int cnt = 0;
                for (int d = lo; d < D && ot[d] <= wn; ++d, ++cnt) {
                    int y = oy[d];
                    if ((unsigned)y >= (unsigned)Kobs) {  // never index the emission table out of bounds
                        y = 0;
                        atomicOr(&s_flag, 2);
                    }
#pragma unroll
                    for (int v = 0; v < K; ++v) {
                        const Real pe = (Real)emis[v * Kobs + y];
                        f[v] = a.obs_product ? f[v] * pe : f[v] + pe;
                    }
                }
#pragma unroll
                for (int v = 0; v < K; ++v)
                    obsf[i * K + v] = (cnt == 0 || power == (Real)0) ? (Real)1
                                      : (power == (Real)1 ? f[v] : M::pow(f[v], power));
            }
        }
        __syncthreads();
        if (s_flag) {
            if (tid == 0) {
                a.status[c] = ((s_flag & 1) ? SMJP_ST_GRID_NOT_INCREASING : 0) | ((s_flag & 2) ? SMJP_ST_BAD_INPUT : 0);
                a.path_len[c] = 0;
            }
            continue;
        }

        const bool exporting = a.messages != nullptr;  // exported rows are the reference's alpha, deferred one step
        Real* msg = exporting ? (Real*)a.messages + a.msg_offs[c]
                              : (Real*)a.msg_scratch + (int64_t)blockIdx.x * a.msg_scratch_stride;

        // =========================== forward ===========================
        Real invZ_prev = (Real)1;
        Real J[K], of[K];
#pragma unroll
        for (int v = 0; v < K; ++v) {
            J[v] = (Real)(1.0 / K);  // uniform prior on (v,0): smjp_utils.py:286-289
            of[v] = obsf[v];
        }
        bool degenerate = false;
        // intervals whose right end "is" t_end drop the Poisson-event term (smjp_utils.py:544): with an
        // increasing grid these are the last ones, from i_final on (normally only n-1)
        int i_final = n;
        while (i_final > 0 && np_isclose(a.t_end, Wd[i_final])) --i_final;
        Real* row_base = msg;  // start of row i-1 (row r holds K*(r+1) values)
        const Real* obsf_next = obsf + K;
        // Live window.  A state (v, W[j]) loses at least the thinning factor 1 - 1/Omega per grid step, so old
        // grid points underflow to EXACT zeros -- on cfg2 56 % of the fp64 and 85 % of the fp32 message entries
        // -- and a zero stays zero.  j_lo is the first grid point with a non-zero entry: pairs, row stores and
        // row loads below it are skipped, which changes no result bit.  Exported messages keep j_lo = 0.
        int j_lo = 0;
        unsigned long long live = 0;
        const Real weps = (Real)a.window_eps;
        for (int i = 0; i < n; ++i) {
            const double wn = Wd[i + 1];
            const bool nonfinal = i < i_final;
            const int m_lo = j_lo / B, t_lo = j_lo - m_lo * B;  // slot / thread of the first live grid point
            if (tid == 0) jlo_arr[i] = j_lo;
            live += (unsigned long long)(i - j_lo + 1);
            const bool judge = !exporting && (i & 3) == 0;  // the window start is re-examined every fourth step
            if (judge) {
                // smallest grid point of the oldest slot that still carries mass after step i-1: above zero
                // (exact mode), or above window_eps times that row's sum (pruned mode, SURVEY.md 7 hard part 2)
                int mylive = 0x7fffffff;
                const int j0 = m_lo * B + tid;
                if (j0 >= j_lo && j0 < i) {
                    const Real thr = weps > (Real)0 ? weps / invZ_prev : (Real)0;  // eps * Z_{i-1}
                    const Real* q = st_a + (size_t)m_lo * K * B + tid;
                    bool alive = false;
#pragma unroll
                    for (int v = 0; v < K; ++v) alive = alive || q[v * B] > thr;
                    if (alive) mylive = j0;
                }
                mylive = __reduce_min_sync(0xffffffffu, mylive);
                if (lane == 0) s_jmin[(i & 1) * NW + wid] = mylive;
            }
            Real ofn[K];  // next step's observation factors, fetched under this step's arithmetic
#pragma unroll
            for (int v = 0; v < K; ++v) ofn[v] = (i + 1 < n) ? obsf_next[v] : (Real)1;
            obsf_next += K;
            Real Zp = 0;
            Real Jp[K];
#pragma unroll
            for (int v = 0; v < K; ++v) Jp[v] = 0;
            const int mi = i / B, ti = i - mi * B;  // slot and owner thread of the new state (v, W[i])
            Real* rowp = row_base + tid;  // row i-1: i entries per state (exported messages: m_lo == 0)
            Real* rowg = row_base + (int64_t)K * i + m_lo * B + tid;  // row i (scratch mode): i+1 entries per state
            Real* pa = st_a + (size_t)m_lo * K * B + tid;
            Real* ph = st_h + (size_t)m_lo * K * B + tid;
            const double* pw = Wd + m_lo * B + tid;
            // one grid point (i, j): all K*K hazards, emission, thinned update, jump contributions
            auto pair = [&](auto partial_tag, bool diag) {
                constexpr bool PARTIAL = decltype(partial_tag)::value;
                const Real tau = (Real)(wn - *pw);
                const Real L2 = ex2.log2(tau);
#pragma unroll
                for (int v = 0; v < K; ++v) {
                    const Real ap = pa[v * B] * invZ_prev;
                    Real hprev = ph[v * B];
                    Real base = ap * thin;
                    if (PARTIAL) {
                        if (exporting && !diag) rowp[(int64_t)v * i] = ap;
                        base = diag ? J[v] : base;
                        hprev = diag ? (Real)0 : hprev;
                    } else {
                        if (exporting) rowp[(int64_t)v * i] = ap;
                    }
                    Real kE[K];
                    Real esum = 0, dsum = 0;
#pragma unroll
                    for (int s = 0; s < K; ++s) {
                        kE[s] = ex2(PC(1, v * K + s) * L2 - PC(2, v * K + s));  // A_vs(tau)
                        esum += kE[s] * PC(0, v * K + s);
                        dsum += kE[s];
                    }
                    const Real hsum = tau * esum;  // H_v(tau)
                    const Real g = of[v] * ex2.neg(-om_l2e * (hsum - hprev)) * base;
                    // scratch rows hold g = alpha / (Omega A_v) of row i itself (row i starts K*i past row i-1):
                    // the backward jump weight alpha A_{v,v+} / (Omega A_v) is then g A_{v,v+}, one hazard per entry
                    if (!exporting) rowg[v * (i + 1)] = g;
                    Real gj = g;
                    // a final row (right end = t_end, smjp_utils.py:544-547) holds alpha = g WITHOUT the Omega A_v factor, so the
                    // 1/(Omega A_v) of the jump term (smjp_utils.py:409-412) does not cancel there (matters when a grid point lies
                    // within isclose() of t_end: two or more final rows)
                    if (!nonfinal) gj = dsum > (Real)0 ? g * M::rcp(omega_r * dsum) : (Real)0;
#pragma unroll
                    for (int s = 0; s < K; ++s) Jp[s] += gj * kE[s];
                    const Real an = g * (nonfinal ? omega_r * dsum : (Real)1);
                    pa[v * B] = an;
                    ph[v * B] = hsum;
                    Zp += an;
                }
            };
            // ONE copy of the pair code serves every slot (the kernel is 90 KB of SASS; a second, diag-free copy
            // for the full slots cost more in instruction fetch than its two selects per state save): the
            // oldest slot masks the threads below the window start, the newest those above the new grid point
#pragma unroll 1
            for (int m = m_lo; m <= mi; ++m) {
                const bool active = (m > m_lo || tid >= t_lo) && (m < mi || tid <= ti);
                if (active) pair(std::true_type(), m == mi && tid == ti);
                pa += K * B;
                ph += K * B;
                pw += B;
                rowp += B;
                rowg += B;
            }
            // block reduction of (Z, J[0..K)) -- double-buffered scratch, one barrier per step
            Real* rb = red + (i & 1) * NW * (K + 1);
            {
                constexpr int NVP = pow2_ceil(K + 1), SH = 5 - log2_of(NVP);
                Real vals[NVP];
                vals[0] = Zp;
#pragma unroll
                for (int v = 0; v < K; ++v) vals[1 + v] = Jp[v];
#pragma unroll
                for (int v = K + 1; v < NVP; ++v) vals[v] = 0;
                WarpHalve<Real, NVP, 16>::run(vals, lane);
                const int idx = lane >> SH;
                if ((lane & ((1 << SH) - 1)) == 0 && idx < K + 1) rb[wid * (K + 1) + idx] = vals[0];
            }
            __syncthreads();
            Real Z = 0;
#pragma unroll
            for (int v = 0; v < K; ++v) J[v] = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                Real part[K + 1];
                lds_vec<Real, K + 1>(rb + w * (K + 1), part);  // one or two 128-bit loads when K + 1 == 4
                Z += part[0];
#pragma unroll
                for (int v = 0; v < K; ++v) J[v] += part[1 + v];
            }
            if (!(Z > (Real)0) || !(Z < (Real)INFINITY)) {
                degenerate = true;
                break;
            }
            invZ_prev = (Real)1 / Z;
#pragma unroll
            for (int v = 0; v < K; ++v) {
                J[v] *= invZ_prev;
                of[v] = ofn[v];
            }
            if (a.logZ && tid == 0) a.logZ[w0 + i] = log((double)Z);
            if (i > 0) row_base += (int64_t)K * i;  // row i-1 done: advance to row i
            if (judge) {
                int jl = 0x7fffffff;
#pragma unroll
                for (int w = 0; w < NW; ++w) jl = min(jl, s_jmin[(i & 1) * NW + w]);
                j_lo = max(j_lo, min(jl, min((m_lo + 1) * B, i)));  // the states of step i itself are not judged yet
            }
        }
        if (a.live_pairs && tid == 0) atomicAdd(a.live_pairs, live);
        if (degenerate) {
            if (tid == 0) {
                a.status[c] = SMJP_ST_DEGENERATE;
                a.path_len[c] = 0;
            }
            continue;
        }
        if (exporting) {  // last row: scale and store (scratch rows were complete when their step ended)
            Real* row_last = msg + (int64_t)K * (n - 1) * n / 2;
            for (int j = tid; j < n; j += B) {
                const Real* p = st_a + (size_t)(j / B) * K * B + tid;
#pragma unroll
                for (int v = 0; v < K; ++v) row_last[(int64_t)v * n + j] = p[v * B] * invZ_prev;
            }
        }
        if (tid == 0) s_found[0] = s_found[1] = 0;
        __syncthreads();

        // =========================== backward ===========================
        Real* wbuf = st_a;  // [K][wl] weights of the current draw, each entry written and read by its own thread only
        int cur = n - 1, vplus = -1, cnt = 0, u_hi = -1;
        while (true) {
            const int len = cur + 1;
            const Real* row = msg + (int64_t)K * cur * (cur + 1) / 2;
            const int lo = exporting ? 0 : jlo_arr[cur];  // live window of this row: entries below lo are exact zeros
            const int wl = len - lo;
            {   // The next draw reads row j*-1 where W[j*] starts the sojourn found now; sojourns are a few grid
                // steps long, so rows cur-1 .. cur-3 (just below this row) cover ~7/8 of the cases: pull their
                // live parts into L2 now, one draw ahead of the HBM round trip that would otherwise be exposed.
                // One 128-byte line per thread: thread t takes line t / (3K) of segment t % (3K) = (row, state).
                const int seg = tid % (3 * K), line = tid / (3 * K);
                const int r = cur - 1 - seg / K, v = seg - (seg / K) * K;
                if (r >= 0) {
                    const int rlo = exporting ? 0 : jlo_arr[r];
                    const char* p = (const char*)(msg + (int64_t)K * r * (r + 1) / 2 + (int64_t)v * (r + 1) + rlo) + line * 128;
                    if (p < (const char*)(msg + (int64_t)K * r * (r + 1) / 2 + (int64_t)(v + 1) * (r + 1)))
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                }
            }
            if (!a.uniforms && (u_hi < 0 || cur <= u_hi - B)) {  // refill B uniforms: steps cur, cur-1, ...
                if (cur - tid >= 0)
                    ubuf[tid] = philox_u01(a.seed, TAG_BACKWARD, (uint64_t)(cur - tid), 0u, a.sweep,
                                           a.chain_base + (uint32_t)c);
                u_hi = cur;
            }
            // Categorical draw over the flat vector w[v * wl + (j - lo)] (state-major, the reference's order) without
            // materialising it: every thread owns a contiguous run of grid points, computes the K weights of each on
            // the fly (one log2 per grid point, shared by the states), and the prefix sums are K per-state scans --
            // one barrier for the warp totals, one for the verdict; the owner walks its own few weights (parked in
            // shared memory by itself: no other thread reads them) to find the entry.  (An earlier version wrote the weights to shared memory and scanned thread-chunks of the flat
            // vector: a third barrier and a store/load round trip per entry, 8 % of the kernel's stall samples.)
            const int mode = vplus < 0 ? 0 : ((!exporting && cur < i_final) ? 1 : 2);
            const double wnb = Wd[cur + 1];
            // weights of grid point j for all states (mode 0: a_{n-1} ~ alpha_{n-1}, fast_hmm.py:124-125; mode 1: the
            // scratch row holds g, w = g_cur(v,j) A_{v,vplus}(tau); mode 2: w = alpha_cur(v,j) A_{v,vplus}/A_v,
            // fast_hmm.py:153-171)
            auto weights = [&](int j, Real (&w)[K]) {
                Real rv[K];
#pragma unroll
                for (int v = 0; v < K; ++v) rv[v] = row[(int64_t)v * len + j];
                if (mode == 0) {
#pragma unroll
                    for (int v = 0; v < K; ++v) w[v] = rv[v];
                    return;
                }
                const Real L2 = ex2.log2((Real)(wnb - Wd[j]));
                if (mode == 1) {
#pragma unroll
                    for (int v = 0; v < K; ++v) w[v] = rv[v] * ex2(skm1[v * K + vplus] * L2 - sc2p[v * K + vplus]);
                    return;
                }
#pragma unroll
                for (int v = 0; v < K; ++v) {
                    Real dsum = 0, hto = 0;
#pragma unroll
                    for (int s2 = 0; s2 < K; ++s2) {
                        const Real kE = ex2(skm1[v * K + s2] * L2 - sc2p[v * K + s2]);
                        dsum += kE;
                        hto = (s2 == vplus) ? kE : hto;
                    }
                    const Real r = dsum > (Real)0 ? hto * M::rcp(dsum) : (Real)0;
                    w[v] = rv[v] * r;
                }
            };
            const int q = (wl + B - 1) / B;
            const int jb = min(lo + tid * q, len), je = min(jb + q, len);  // this thread's grid points
            Acc loc[K];
#pragma unroll
            for (int v = 0; v < K; ++v) loc[v] = 0;
            for (int j = jb; j < je; ++j) {
                Real w[K];
                weights(j, w);
#pragma unroll
                for (int v = 0; v < K; ++v) {
                    loc[v] += (Acc)w[v];
                    wbuf[v * wl + (j - lo)] = w[v];  // kept for this thread's own walk, should it own the draw
                }
            }
            Acc incl[K];
#pragma unroll
            for (int v = 0; v < K; ++v) incl[v] = warp_inclusive_scan(loc[v], lane);
            Acc* wsum = (Acc*)red;  // [NW][K] warp totals per state (the forward pass's reduction scratch is free now)
            if (lane == 31) {
#pragma unroll
                for (int v = 0; v < K; ++v) wsum[wid * K + v] = incl[v];
            }
            __syncthreads();
            if (tid == 0) s_found[(cnt + 1) & 1] = 0;  // the next draw's flag: nobody touches it during this draw
            // state totals and the state the target falls into
            Acc cum[K + 1];
            cum[0] = 0;
#pragma unroll
            for (int v = 0; v < K; ++v) {
                Acc tv = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) tv += wsum[w * K + v];
                cum[v + 1] = cum[v] + tv;
            }
            const Acc total = cum[K];
            if (!(total > (Acc)0) || !(total < (Acc)INFINITY)) {
                degenerate = true;
                break;
            }
            const double u = a.uniforms ? a.uniforms[w0 + cur] : ubuf[u_hi - cur];
            const Acc target = (Acc)u * total;
            int vs = -1;  // first state whose cumulative total exceeds the target (none: u * total rounded up to total)
#pragma unroll
            for (int v = K - 1; v >= 0; --v) vs = (target < cum[v + 1]) ? v : vs;
            if (vs >= 0) {
                Acc my_loc = 0, my_incl = 0, off = 0;
#pragma unroll
                for (int v = 0; v < K; ++v) {
                    if (v == vs) {
                        my_loc = loc[v];
                        my_incl = incl[v];
                        off = cum[v];
                    }
                }
                for (int w = 0; w < wid; ++w) off += wsum[w * K + vs];
                my_incl += off;
                Acc excl = __shfl_up_sync(0xffffffffu, my_incl, 1);
                if (lane == 0) excl = off;
                // owner: excl <= target < incl (the intervals partition [cum[vs], cum[vs+1]) exactly); runs without
                // mass never own a draw (a shuffle scan over zero-padded lanes is not monotone to the last bit)
                const bool has = je > jb && my_incl > excl && my_loc > (Acc)0;
                if (has && target >= excl && target < my_incl) {
                    Acc acc = excl;
                    int sel = -1, lastpos = jb;
                    for (int j = jb; j < je; ++j) {
                        const Acc wv = (Acc)wbuf[vs * wl + (j - lo)];
                        if (wv > (Acc)0) lastpos = j;
                        acc += wv;
                        if (acc > target) { sel = j; break; }
                    }
                    s_sel = sel >= 0 ? sel : lastpos;
                    s_selv = vs;
                    s_found[cnt & 1] = 1;
                }
            }
            __syncthreads();
            // No owner: u*total rounded up to the total, or the target fell between the inclusive value of the last
            // thread that holds weights and the warp total (in fp32 that happens about once per 1e7 draws).  The last
            // entry with mass in the flat order -- largest state, then largest grid point -- takes the draw, as the
            // reference's inverse CDF does at u -> 1.
            if (!s_found[cnt & 1]) {  // block-uniform
                int key = 0;
#pragma unroll
                for (int v = 0; v < K; ++v) key = (je > jb && loc[v] > (Acc)0) ? v * B + tid + 1 : key;
                if (key) atomicMax(&s_flag, key);
                __syncthreads();
                if (key && key == s_flag) {
                    const int vf = (key - 1) / B;
                    int lastpos = jb;
                    for (int j = jb; j < je; ++j)
                        if ((Acc)wbuf[vf * wl + (j - lo)] > (Acc)0) lastpos = j;
                    s_sel = lastpos;
                    s_selv = vf;
                }
                __syncthreads();
                if (tid == 0) s_flag = 0;
            }
            const int j = s_sel;
            const int v = s_selv;
            if (a.aug_idx)
                for (int t = j + tid; t <= cur; t += B) a.aug_idx[w0 + t] = v * n + j;
            if (tid == 0) {
                soj[cnt] = v;
                soj[(ncap + 1) + cnt] = j;
            }
            ++cnt;
            if (j == 0) break;
            vplus = v;
            cur = j - 1;
        }
        __syncthreads();
        if (degenerate) {
            if (tid == 0) {
                a.status[c] = SMJP_ST_DEGENERATE;
                a.path_len[c] = 0;
            }
            continue;
        }
        // ---- path compaction (smjp_utils.py:319-330): sojourns were found last-to-first ----
        for (int p = tid; p < cnt; p += B) {
            a.path_v[w0 + p] = soj[cnt - 1 - p];
            a.path_t[w0 + p] = Wd[soj[(ncap + 1) + cnt - 1 - p]];
        }
        const int v_last = soj[0];
        const double t_last = Wd[soj[ncap + 1]];
        const bool add_end = !np_isclose(t_last, a.t_end);  // smjp_utils.py:338-348
        const int plen = cnt + (add_end ? 1 : 0);
        if (tid == 0) {
            if (add_end) {
                a.path_v[w0 + cnt] = v_last;
                a.path_t[w0 + cnt] = a.t_end;
            }
            a.path_len[c] = plen;
            a.status[c] = 0;
        }
        // ---- chain metrics (mcmc_utils.py:40-56): one thread per state, fixed summation order ----
        if (tid < K && (a.time_in_state || a.jumps)) {
            double tsum = 0;
            int jc = 0;
            for (int p = 0; p < plen; ++p) {
                const int vs = p < cnt ? soj[cnt - 1 - p] : v_last;
                if (vs == tid) {
                    ++jc;
                    if (p + 1 < plen) {
                        const double t0 = Wd[soj[(ncap + 1) + cnt - 1 - p]];
                        const double t1 = (p + 1 < cnt) ? Wd[soj[(ncap + 1) + cnt - 2 - p]] : a.t_end;
                        tsum += t1 - t0;
                    }
                }
            }
            if (a.time_in_state) a.time_in_state[(int64_t)c * K + tid] = tsum;
            if (a.jumps) a.jumps[(int64_t)c * K + tid] = jc;
        }
    }
}

#undef PC

}  // namespace smjp
