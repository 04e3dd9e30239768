// Bootstrap particle filter for the sMJP, one CTA per chain (filter run).
//
// Replaces smjp_smc (pkg/pmcmc.py:224-310):
//   propagate  each particle (v, l) by competing conditional-Weibull risks until its next jump would
//              pass the observation time (:261-289; WeibullDistribution.sample(hold_time=h),
//              pkg/distributions.py:296-304, in the cancellation-free form
//              tau = lambda ((h/lambda)^k - log(1-u))^(1/k));
//   weight     w_p = P(y_d | v_p) (:279-281, smjpEmission.likelihood);
//   normalise  Z_d = log sum_p w_p - log N (:294-297);
//   resample   multinomial (:302-304, utils.py:7-8) or systematic (new) by inverse CDF: block-wide
//              prefix scan of the weights into a CDF, then a binary search per offspring;
//   output     the path of one particle after the last resampling + (v_last, t_final) (:307-309):
//              particle 0 as in the reference for multinomial resampling, a uniformly drawn one for
//              systematic resampling (whose offspring are ordered, index 0 would be biased).
// Deviations from the reference, both deliberate: resampled particles get copy
// semantics (the reference aliases Python lists), and each particle's clock starts at the previous
// observation time (the reference leaks t_curr from one particle into the next).
// Paths are not stored per particle: only the ancestor table anc[d][p] is.  Every draw is
// counter-based (Philox keyed by chain, observation, particle slot), so after the last step one
// thread walks the lineage of particle 0 back through anc[][] and REPLAYS its propagation to emit
// the jump times.  oracle/smjp_oracle.py::particle_filter is the CPU twin.
#pragma once
#include "../../include/smjp_b200.h"
#include <cooperative_groups.h>

#include "common.cuh"
#include "ffbs.cuh"

namespace smjp {

struct PfArgs {
    int C, N, K, Kobs;
    const int64_t* obs_offs;
    const double* obs_t;
    const int32_t* obs_y;
    const double* model;  // shape[K*K], c2[K*K], emis[K*Kobs], scale[K*K]
    const double* pi0;    // device [K]
    double t_end;
    uint64_t seed;
    uint32_t chain_base;
    int resampler;
    int max_iters;        // cap on jumps per particle per observation interval
    int family;           // SMJP_HAZARD_WEIBULL | SMJP_HAZARD_GAMMA (shape = a, scale = theta)
    int extend;           // 1: also simulate the returned path on (t_last_obs, t_end] (the reference does not)
    // per-chain workspaces, chain stride given
    int32_t* pv;          // [C][2][N] particle state (double-buffered)
    double* pl;           // [C][2][N] time of last jump
    double* cdf;          // [C][N]
    int32_t* anc;         // [C][Dmax][N] ancestors
    int64_t anc_stride;   // Dmax * N
    int32_t* lineage;     // [C][Dmax]
    int64_t lineage_stride;
    // outputs
    double* logZ;         // [total_obs] or null
    double* loglik;       // [C]
    const int64_t* p_offs;  // path capacity windows
    int32_t* path_v;
    double* path_t;
    int32_t* path_len;
    int32_t* status;
};

// ------------------------------------------------------------------------------------------------
// Gamma holding times (not in the reference, whose hazards are "weibull for now", smjp_utils.py:419-423;
// BASELINE configs[2] asks for a gamma-hazard sMJP).  tau ~ Gamma(shape a, scale theta) given tau > h is
// drawn by inverting the survival function: Q(a, tau/theta) = (1 - u) Q(a, h/theta), Q the regularised
// upper incomplete gamma function (series / continued fraction, inverse by Halley steps from a
// Wilson-Hilferty start).  Twin: oracle cond_gamma (scipy.special.gammaincc / gammainccinv).
// ------------------------------------------------------------------------------------------------
// lga = lgamma(a): a model constant, tabulated once per filter run (it was recomputed ~10 times per draw).
// Both expansions run WITHOUT a division per term (an fp64 division is ~30 instructions, the term itself 3): the
// series sum is carried as a fraction P/D, the continued fraction by its forward recurrence A_i/B_i with the
// convergence test cross-multiplied; each ends in one division.
// e: the table-driven fp64 2^x / log2 of ffbs.cuh (libdevice log + exp were a fifth of the filter's instructions);
// ax_out (optional) receives x^a e^-x / Gamma(a) = x * pdf(x), which the inverse's Newton step needs anyway.
__device__ inline double igamc(const Exp2<double>& e, double a, double x, double lga, double* ax_out = nullptr) {
    constexpr double S = Exp2<double>::SCALE, L2E = 1.4426950408889634;
    if (ax_out) *ax_out = 0.0;
    if (!(x > 0.0)) return 1.0;
    const double lax2 = fma(a, e.log2(x), -(x + lga) * L2E);  // log2 of x^a e^-x / Gamma(a)
    if (lax2 < -1020.0) return x < a + 1.0 ? 1.0 : 0.0;
    const double ax = e(S * lax2);
    if (ax_out) *ax_out = ax;
    if (x < a + 1.0) {  // P(a,x) by series: sum_n x^n / ((a+1)...(a+n)) = P_n / D_n
        double r = a, N = 1.0, P = 1.0, D = 1.0;
        for (int it = 0; it < 512; it += 2) {
            r += 1.0;
            N *= x;
            P = fma(P, r, N);
            D *= r;
            r += 1.0;
            N *= x;
            P = fma(P, r, N);
            D *= r;
            if (N <= 1e-17 * P) break;
            if ((it & 30) == 30 && D > 1e100) { N *= 1e-100; P *= 1e-100; D *= 1e-100; }  // 32 terms grow D by < 1e90
        }
        return 1.0 - (P * ax) / (D * a);
    }
    // Q(a,x) = ax / (b_0 + a_1/(b_1 + a_2/(b_2 + ...))), a_i = -i (i - a), b_i = x + 1 - a + 2 i
    double b = x + 1.0 - a;
    double A1 = 1.0, B1 = 0.0, A0 = b, B0 = 1.0;  // (A_{i-2}, B_{i-2}), (A_{i-1}, B_{i-1})
    for (int i = 1; i < 500; ++i) {
        const double an = -(double)i * ((double)i - a);
        b += 2.0;
        const double A = fma(b, A0, an * A1), B = fma(b, B0, an * B1);
        const double cross = fabs(fma(A, B0, -(A0 * B)));  // |f_i - f_{i-1}| * |B B0|
        A1 = A0; B1 = B0; A0 = A; B0 = B;
        if (cross <= 1e-16 * fabs(A * B1)) break;
        if ((i & 15) == 15 && (fabs(A0) > 1e100 || fabs(B0) > 1e100)) { A0 *= 1e-100; B0 *= 1e-100; A1 *= 1e-100; B1 *= 1e-100; }
    }
    return ax * (B0 / A0);
}

__device__ inline double igamci(const Exp2<double>& e, double a, double q, double lga) {  // x with Q(a, x) = q
    if (q >= 1.0) return 0.0;
    if (!(q > 0.0)) return INFINITY;
    const double t = -normcdfinv(q), d = 1.0 / (9.0 * a), y = 1.0 - d + t * sqrt(d);
    double x = a * y * y * y;
    if (x <= 0.0 || a < 1.0) {
        const double p = 1.0 - q;
        const double x0 = exp((log(fmax(p, 1e-300)) + lga + log(a)) / a);  // lgamma(a + 1) = lgamma(a) + log a
        if (x <= 0.0) x = x0;
        else if (p < 0.5) x = x0;
        else x = fmax(x, x0);
    }
    for (int it = 0; it < 60; ++it) {
        double ax;
        const double f = igamc(e, a, x, lga, &ax) - q;
        const double fp = -ax / x;
        if (fp == 0.0) break;
        double dx = f / fp;
        const double den = 1.0 - 0.5 * dx * ((a - 1.0) / x - 1.0);
        if (den > 0.2) dx /= den;
        double xn = x - dx;
        if (!(xn > 0.0)) xn = 0.5 * x;
        const bool done = fabs(xn - x) <= 1e-15 * xn;
        x = xn;
        if (done) break;
    }
    return x;
}

// The same inverse when the root is known to lie in [lo, hi] with Q(lo) = qlo > q > Q(hi) = qhi (the holding time
// of a particle that jumps before its next observation: an interval one observation spacing wide).  Secant start,
// Halley steps kept inside the bracket: 2-3 evaluations of Q instead of the ~6 of the global search.
__device__ inline double igamci_bracketed(const Exp2<double>& e, double a, double q, double lga, double lo, double hi,
                                          double qlo, double qhi) {
    double x = lo + (qlo - q) / (qlo - qhi) * (hi - lo);
    if (!(x > lo && x < hi)) x = 0.5 * (lo + hi);
    for (int it = 0; it < 60; ++it) {
        double ax;
        const double f = igamc(e, a, x, lga, &ax) - q;
        if (f > 0.0) lo = x; else hi = x;  // Q decreases: Q(x) > q puts the root above x
        const double fp = -ax / x;
        double xn, tol = 1e-15;
        if (fp == 0.0 || !isfinite(fp)) {
            xn = 0.5 * (lo + hi);
        } else {
            double dx = f / fp;
            const double den = 1.0 - 0.5 * dx * ((a - 1.0) / x - 1.0);
            // Halley converges cubically (Newton quadratically): after a step of relative size 1e-6 (1e-9) the next
            // one is below the rounding of x and its evaluation of Q is saved; a bisection step promises nothing
            if (den > 0.2) { dx /= den; tol = 1e-6; } else tol = 1e-9;
            xn = x - dx;
            if (!(xn > lo && xn < hi)) { xn = 0.5 * (lo + hi); tol = 1e-15; }
        }
        const bool done = fabs(xn - x) <= tol * xn || !(hi > lo);
        x = xn;
        if (done) break;
    }
    return x;
}

// Arithmetic of one filter run: table-driven fp64 2^x / log2 (ffbs.cuh: the FP64 pipe is the scarce resource;
// libdevice pow + log1p cost ~450 FP64-pipe instructions per draw, this form ~35), 1/k cached in shared memory.
struct PfMath {
    Exp2<double> ex2;
    const double* shape;  // k (Weibull) or a (gamma)              [K*K] global
    const double* c2;     // k log2(lambda)                        [K*K] global
    const double* scale;  // lambda / theta                        [K*K] global
    const double* rk;     // 1/k: shared-memory cache, or null -> divide
    const double* l2lam;  // log2(lambda): shared-memory cache, or null -> c2 / k
    const double* lga;    // gamma family: lgamma(a), shared-memory cache, or null -> lgamma()
};

// One round of competing risks for a particle in state v that has held for `hold`: K conditional holding-time
// draws, the smallest wins.  Weibull: tau_s = lambda ((hold/lambda)^k - log(1-u))^(1/k) (distributions.py:296-304
// in cancellation-free form; 1-u is exact for a 53-bit uniform) is compared in the log2 domain,
// log2 tau_s = log2 lambda + log2(e_s)/k, so that only the winner is exponentiated.  Uniform r = it*K + s of
// stream (seed, TAG_PF_PROPAGATE, slot, d, chain), two per Philox block.
//
// Gamma: the caller only needs the smallest draw if it falls before the next observation, `hnext` after the sojourn
// began.  tau_s < hnext  <=>  (1-u) Q(a, hold/theta) > Q(a, hnext/theta), so two evaluations of Q decide a target and
// the inverse is taken only for the (few) targets that do jump, inside the bracket [hold, hnext]; a round without a
// jump returns +inf.  The draws are those of the full inversion (oracle cond_gamma) wherever both are exact.
__device__ __forceinline__ double pf_round(const PfArgs& a, const PfMath& m, uint32_t chain, uint32_t d, uint32_t slot,
                                           int it, int v, double hold, double hnext, int& arg) {
    constexpr double S = Exp2<double>::SCALE;
    const int K = a.K;
    const bool gamma = a.family == SMJP_HAZARD_GAMMA;
    const double l2h = (!gamma && hold > 0.0) ? m.ex2.log2(hold) : -INFINITY;
    double best = INFINITY;
    arg = 0;
    int ncand = 0;                            // gamma: targets that jump before hnext
    double c_q = 0.0, c_qlo = 1.0, c_qhi = 0.0;  // the first of them, not yet inverted
    Philox4 blk;
    uint32_t blk_idx = 0xFFFFFFFFu;
    for (int s = 0; s < K; ++s) {
        const uint32_t r = (uint32_t)(it * K + s);
        if ((r >> 1) != blk_idx) {
            blk_idx = r >> 1;
            blk = philox4x32_10(blk_idx, slot, d, chain, (uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ TAG_PF_PROPAGATE);
        }
        const double u = (r & 1) ? uniform53(blk.x[2], blk.x[3]) : uniform53(blk.x[0], blk.x[1]);
        const int t = v * K + s;
        double key;  // gamma: tau itself; Weibull: log2 tau
        if (gamma) {
            const double k = m.shape[t], lam = m.scale[t];
            const double lga = m.lga ? m.lga[t] : lgamma(k);
            const double xlo = hold / lam, xhi = hnext / lam;
            const double qlo = igamc(m.ex2, k, xlo, lga), qhi = igamc(m.ex2, k, xhi, lga);
            const double q = (1.0 - u) * qlo;
            key = INFINITY;
            if (q > qhi) {
                // the inverse is deferred: nearly always ONE target jumps, and it is then inverted after the loop,
                // where the jumping lanes of the warp run the iteration together instead of target by target
                if (ncand == 0) {
                    ncand = 1; arg = s;
                    c_q = q; c_qlo = qlo; c_qhi = qhi;
                } else {
                    if (ncand == 1) {
                        const int t1 = v * K + arg;
                        const double k1 = m.shape[t1], lam1 = m.scale[t1];
                        best = lam1 * igamci_bracketed(m.ex2, k1, c_q, m.lga ? m.lga[t1] : lgamma(k1), hold / lam1, hnext / lam1, c_qlo, c_qhi);
                    }
                    ncand = 2;
                    key = lam * igamci_bracketed(m.ex2, k, q, lga, xlo, xhi, qlo, qhi);
                }
            }
        } else {
            const double k = m.shape[t];
            const double hk = l2h == -INFINITY ? 0.0 : m.ex2(S * (k * l2h - m.c2[t]));  // (hold/lambda)^k
            const double e = hk - m.ex2.log2(1.0 - u) * 0.6931471805599453;
            const double rk = m.rk ? m.rk[t] : 1.0 / k;
            const double l2lam = m.l2lam ? m.l2lam[t] : m.c2[t] / k;
            key = e > 0.0 ? l2lam + rk * m.ex2.log2(e) : -INFINITY;
        }
        if (key < best) { best = key; arg = s; }
    }
    if (ncand == 1) {
        const int t1 = v * K + arg;
        const double k1 = m.shape[t1], lam1 = m.scale[t1];
        best = lam1 * igamci_bracketed(m.ex2, k1, c_q, m.lga ? m.lga[t1] : lgamma(k1), hold / lam1, hnext / lam1, c_qlo, c_qhi);
    }
    return gamma ? best : (best == -INFINITY ? 0.0 : m.ex2(S * best));
}

// one propagation of a particle from clock t_cur towards t_obs; optionally records jumps
__device__ __forceinline__ int pf_propagate(const PfArgs& a, const PfMath& m, uint32_t chain, uint32_t d, uint32_t slot,
                                            int& v, double& l, double t_cur, double t_obs, int32_t* out_v,
                                            double* out_t, int out_cap, int* out_n) {
    for (int it = 0; it < a.max_iters; ++it) {
        int arg;
        const double w_next = l + pf_round(a, m, chain, d, slot, it, v, t_cur - l, t_obs - l, arg);
        if (w_next >= t_obs) return 0;
        t_cur = w_next;
        l = w_next;
        v = arg;
        if (out_n) {
            if (*out_n >= out_cap) return SMJP_ST_OVERFLOW;
            out_v[*out_n] = v;
            out_t[*out_n] = l;
            ++*out_n;
        }
    }
    return SMJP_ST_OVERFLOW;  // more than max_iters jumps inside one observation interval
}

constexpr int PF_MAXW = 2048;  // (tiles x warps) of one CTA's particle slice
constexpr int PF_RK = 256;     // 1/k cache entries (K <= 16)

// One thread-block CLUSTER of CS CTAs per filter run: the N particles are split over the CTAs of the
// cluster; thread t of a CTA owns particles p_lo + t, p_lo + t + B, ... (coalesced).  Per observation:
//   1. gather the state of the ancestor chosen by the previous resampling, propagate, weight; inclusive warp
//      scans of the weights (warp-local sums to cdf[], warp totals to shared memory);
//   2. block scan of the warp totals; CTA totals are exchanged through distributed shared memory (cluster
//      barrier A), which gives every particle its global CDF value c_i;
//   3. resampling.  Systematic: offspring p descends from the first i with c_i > (u + p)/N total, i.e. particle
//      i owns the slots [S_{i-1}, S_i), S_i = #{p : target_p < c_i}: each particle computes its S_i (estimate +
//      exact predicate fix-up), takes S_{i-1} from its neighbour and writes its own index into those slots --
//      no search, no CDF round trip through memory.  Multinomial: c_i goes to cdf[], binary search per offspring.
//   cluster barrier B, next observation.  The resampled STATE is never copied: step d+1 gathers through anc[d].
// CS = 1 is the single-CTA filter.
template <int B, int CS>
__global__ void __launch_bounds__(B) pf_kernel(const PfArgs a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int NW = B / 32;
    __shared__ double s_tot[PF_MAXW];   // warp totals, (tile, warp) order
    __shared__ double s_off[PF_MAXW];   // their exclusive prefix within the CTA
    __shared__ double s_rk[PF_RK];
    __shared__ double s_l2lam[PF_RK];
    __shared__ double s_lga[PF_RK];
    __shared__ double s_tab[Exp2<double>::TAB];
    __shared__ double2 s_ltab[Exp2<double>::LTAB];
    __shared__ double s_total;          // this CTA's weight total, read by the whole cluster
    __shared__ double s_last[2];        // warp offset and warp-local sum of this CTA's LAST particle, read by the next CTA
    __shared__ int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rank = CS > 1 ? (int)cluster.block_rank() : 0;
    const int cluster_id = blockIdx.x / CS, n_clusters = gridDim.x / CS;
    const int N = a.N, K = a.K;
    const int Np = (N + CS - 1) / CS;
    const int p_lo = min(rank * Np, N), p_hi = min(p_lo + Np, N);
    const int tiles = (p_hi - p_lo + B - 1) / B, nwt = tiles * NW;
    const double* emis = a.model + 2 * K * K;
    for (int t = tid; t < Exp2<double>::TAB; t += B) s_tab[t] = exp2_table_entry(t);
    for (int t = tid; t < Exp2<double>::LTAB; t += B) {
        const double rc = 1.0 / (1.0 + ((double)t + 0.5) / 128.0);
        s_ltab[t] = make_double2(rc, -::log2(rc));
    }
    if (K * K <= PF_RK)
        for (int t = tid; t < K * K; t += B) {
            s_rk[t] = 1.0 / a.model[t];
            s_l2lam[t] = a.model[K * K + t] / a.model[t];
            s_lga[t] = lgamma(a.model[t]);
        }
    PfMath m;
    m.ex2.tab = s_tab;
    m.ex2.ltab = s_ltab;
    m.shape = a.model;
    m.c2 = a.model + K * K;
    m.scale = a.model + 2 * K * K + K * a.Kobs;
    m.rk = K * K <= PF_RK ? s_rk : nullptr;
    m.l2lam = K * K <= PF_RK ? s_l2lam : nullptr;
    m.lga = K * K <= PF_RK ? s_lga : nullptr;
    __syncthreads();
    for (int c = cluster_id; c < a.C; c += n_clusters) {
        const uint32_t chain = a.chain_base + (uint32_t)c;
        const int64_t o0 = a.obs_offs[c];
        const int D = (int)(a.obs_offs[c + 1] - o0);
        int32_t* pv = a.pv + (int64_t)c * 2 * N;
        double* pl = a.pl + (int64_t)c * 2 * N;
        double* cdf = a.cdf + (int64_t)c * N;
        int32_t* anc = a.anc + (int64_t)c * a.anc_stride;
        if (tid == 0) s_flag = 0;
        // initial states v0 ~ pi0
        double ptot = 0;
        for (int s = 0; s < K; ++s) ptot += a.pi0[s];
        for (int p = p_lo + tid; p < p_hi; p += B) {
            const double u = philox_u01(a.seed, TAG_PF_PROPAGATE, 0, (uint32_t)p, 0xFFFFFFFFu, chain);
            double acc = 0;
            int v = -1, lastpos = 0;
            for (int s = 0; s < K; ++s) {
                const double pr = a.pi0[s] / ptot;
                if (pr > 0) lastpos = s;
                acc += pr;
                if (v < 0 && acc > u) v = s;
            }
            pv[p] = v < 0 ? lastpos : v;
            pl[p] = 0.0;
        }
        __syncthreads();
        int cur = 0;
        double loglik = 0;
        bool dead = false;
        const bool systematic = a.resampler == SMJP_RESAMPLE_SYSTEMATIC;
        for (int d = 0; d < D; ++d) {
            const double t_prev = d == 0 ? 0.0 : a.obs_t[o0 + d - 1];
            const double t_obs = a.obs_t[o0 + d];
            const int y = a.obs_y[o0 + d];
            const int32_t* v_in = pv + cur * N;
            const double* l_in = pl + cur * N;
            int32_t* v_out = pv + (cur ^ 1) * N;
            double* l_out = pl + (cur ^ 1) * N;
            const int32_t* anc_prev = d > 0 ? anc + (int64_t)(d - 1) * N : nullptr;
            int32_t* anc_d = anc + (int64_t)d * N;
            // ---- 1a. gather + propagate.  Particles need different numbers of competing-risk rounds, so the loop
            // is flattened: every trip runs ONE round for each lane, and a lane whose particle has passed t_obs
            // stores it and picks up its next one (p += B) instead of idling until the slowest lane of the warp
            // is done (a warp-synchronous loop over particles was measured at 3.4 rounds per particle, the mean
            // over particles being 1.4).
            {
                int p = p_lo + tid, it = 0, v = 0;
                double l = 0.0, t_cur = t_prev;
                bool have = false;
                while (true) {
                    if (!have && p < p_hi) {
                        int src = p;
                        if (anc_prev) {
                            src = anc_prev[p];
                            // an unwritten slot can only come from CDF values that disagree in the last bit across a
                            // warp boundary; its neighbours' ancestor is then correct to within that rounding
                            if (src < 0) src = p > 0 && anc_prev[p - 1] >= 0 ? anc_prev[p - 1] : (p + 1 < N && anc_prev[p + 1] >= 0 ? anc_prev[p + 1] : p);
                        }
                        v = v_in[src];
                        l = l_in[src];
                        t_cur = t_prev;
                        it = 0;
                        have = true;
                    }
                    if (!__any_sync(0xffffffffu, have)) break;
                    if (have) {
                        int arg;
                        const double w_next = l + pf_round(a, m, chain, (uint32_t)d, (uint32_t)p, it, v, t_cur - l, t_obs - l, arg);
                        bool fin = w_next >= t_obs;
                        if (!fin) {
                            t_cur = w_next;
                            l = w_next;
                            v = arg;
                            if (++it >= a.max_iters) {  // more than max_iters jumps inside one observation interval
                                atomicOr(&s_flag, SMJP_ST_OVERFLOW);
                                fin = true;
                            }
                        }
                        if (fin) {
                            v_out[p] = v;
                            l_out[p] = l;
                            if (systematic) anc_d[p] = -1;
                            have = false;
                            p += B;
                        }
                    }
                }
            }
            // ---- 1b. weights, warp scans in particle order (each thread re-reads the states it just wrote) ----
            for (int t = 0; t < tiles; ++t) {
                const int p = p_lo + t * B + tid;
                const double w = p < p_hi ? emis[v_out[p] * a.Kobs + y] : 0.0;
                const double incl = warp_inclusive_scan(w, lane);
                if (p < p_hi) cdf[p] = incl;  // warp-local inclusive sum
                if (p == p_hi - 1) s_last[1] = incl;
                if (lane == 31) s_tot[t * NW + wid] = incl;
            }
            __syncthreads();
            // ---- 2. exclusive scan of the (tile, warp) totals by warp 0, CTA totals over the cluster ----
            if (wid == 0) {
                double carry = 0;
                for (int base = 0; base < nwt; base += 32) {
                    const double x = base + lane < nwt ? s_tot[base + lane] : 0.0;
                    const double inc = warp_inclusive_scan(x, lane);
                    if (base + lane < nwt) s_off[base + lane] = carry + (inc - x);
                    carry += __shfl_sync(0xffffffffu, inc, 31);
                }
                if (lane == 0) {
                    s_total = nwt > 0 ? s_off[nwt - 1] + s_tot[nwt - 1] : 0.0;
                    s_last[0] = nwt > 0 ? s_off[(p_hi - 1 - p_lo) / 32] : 0.0;
                }
            }
            if (CS > 1) cluster.sync();  // A: every CTA's totals are published
            else __syncthreads();
            double cta_excl = 0, total = 0, prev_excl = 0;
            if (CS > 1) {
                for (int r = 0; r < CS; ++r) {
                    const double tr = *cluster.map_shared_rank(&s_total, r);
                    if (r < rank) { prev_excl = cta_excl; cta_excl += tr; }
                    total += tr;
                }
            } else {
                total = s_total;
            }
            const double Zd = log(total) - log((double)N);
            loglik += Zd;
            if (rank == 0 && tid == 0 && a.logZ) a.logZ[o0 + d] = Zd;
            if (!(total > 0.0)) {  // every particle has zero weight (uniform over the cluster)
                dead = true;
                break;
            }
            // ---- 3. resample ----
            if (systematic) {
                const double u_sys = philox_u01(a.seed, TAG_PF_RESAMPLE, 0, 0u, (uint32_t)d, chain);
                const double n_over_total = (double)N / total;
                const bool n_pow2 = (N & (N - 1)) == 0;  // then x / N == x * (1/N) bit for bit
                const double inv_n = 1.0 / (double)N;
                auto target = [&](int q) {
                    const double x = u_sys + (double)q;
                    return (n_pow2 ? x * inv_n : x / (double)N) * total;
                };
                // S(c) = #{p in [0,N) : ((u + p)/N) total < c}: estimate, then the oracle's own predicate
                auto count_below = [&](double cv) {
                    double e = ceil(cv * n_over_total - u_sys);
                    int q = e < 0.0 ? 0 : (e > (double)N ? N : (int)e);
                    while (q > 0 && !(target(q - 1) < cv)) --q;
                    while (q < N && target(q) < cv) ++q;
                    return q;
                };
                // CDF value of the last particle of the previous CTA, by that particle's own expression
                double c_before = 0.0;
                if (CS > 1 && rank > 0) {
                    const double* pl_last = cluster.map_shared_rank(s_last, rank - 1);
                    c_before = (prev_excl + pl_last[0]) + pl_last[1];
                }
                for (int t = 0; t < tiles; ++t) {
                    const int p = p_lo + t * B + tid;
                    const bool live = p < p_hi;
                    const int idx = t * NW + wid;
                    const double ci = live ? (cta_excl + s_off[idx]) + cdf[p] : 0.0;
                    int s_hi = live ? count_below(ci) : 0;
                    if (live && p == N - 1) s_hi = N;  // inverse-CDF searches that run off the end take the last particle
                    int s_lo = __shfl_up_sync(0xffffffffu, s_hi, 1);
                    if (lane == 0) s_lo = idx > 0 ? count_below((cta_excl + s_off[idx - 1]) + s_tot[idx - 1])
                                                  : (rank > 0 ? count_below(c_before) : 0);
                    if (live)
                        for (int q = s_lo; q < s_hi; ++q) anc_d[q] = p;
                }
            } else {
                for (int t = 0; t < tiles; ++t) {
                    const int p = p_lo + t * B + tid;
                    if (p < p_hi) cdf[p] = (cta_excl + s_off[t * NW + wid]) + cdf[p];
                }
                if (CS > 1) cluster.sync();  // the whole CDF is in global memory
                else __syncthreads();
                for (int p = p_lo + tid; p < p_hi; p += B) {
                    const double pos = philox_u01(a.seed, TAG_PF_RESAMPLE, (uint64_t)p, 0u, (uint32_t)d, chain);
                    const double target = pos * total;
                    int lo = 0, hi = N;  // first index with cdf > target
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (cdf[mid] > target) hi = mid; else lo = mid + 1;
                    }
                    anc_d[p] = lo < N ? lo : N - 1;
                }
            }
            cur ^= 1;
            if (CS > 1) cluster.sync();  // B: ancestors of every slot are written; shared scratch is free again
            else __syncthreads();
        }
        // combine the per-CTA flags on rank 0 and make the last resampling visible to it
        __shared__ int s_flag_pub;
        if (tid == 0) s_flag_pub = s_flag | (dead ? SMJP_ST_DEGENERATE : 0);
        if (CS > 1) cluster.sync();
        else __syncthreads();
        int flag = 0;
        if (rank == 0 && tid == 0)
            for (int r = 0; r < CS; ++r) flag |= CS > 1 ? *cluster.map_shared_rank(&s_flag_pub, r) : s_flag_pub;
        // ---- path of particle 0: lineage through the ancestor table, then replay ----
        if (rank == 0 && tid == 0) {
            const int64_t p0 = a.p_offs[c];
            const int cap = (int)(a.p_offs[c + 1] - p0);
            int plen = 0;
            if (!(flag & SMJP_ST_DEGENERATE) && cap >= 2) {
                int32_t* line = a.lineage + (int64_t)c * a.lineage_stride;  // slot b_d of the surviving lineage
                // which offspring of the last resampling to return: multinomial offspring are i.i.d., so the
                // reference's index 0 (pmcmc.py:307) is a fair draw; systematic offspring are ordered by
                // ancestor, so the index must itself be drawn uniformly
                int b = 0;
                if (a.resampler == SMJP_RESAMPLE_SYSTEMATIC) {
                    const double uk = philox_u01(a.seed, TAG_PF_RESAMPLE, 1, 0u, (uint32_t)D, chain);
                    b = min((int)(uk * (double)N), N - 1);
                }
                for (int d = D - 1; d >= 0; --d) {
                    b = anc[(int64_t)d * N + b];
                    line[d] = b;
                }
                const int slot0 = D > 0 ? line[0] : 0;
                const double u = philox_u01(a.seed, TAG_PF_PROPAGATE, 0, (uint32_t)slot0, 0xFFFFFFFFu, chain);
                double acc = 0;
                int v = -1, lastpos = 0;
                for (int s = 0; s < K; ++s) {
                    const double pr = a.pi0[s] / ptot;
                    if (pr > 0) lastpos = s;
                    acc += pr;
                    if (v < 0 && acc > u) v = s;
                }
                v = v < 0 ? lastpos : v;
                double l = 0.0;
                int32_t* ov = a.path_v + p0;
                double* ot = a.path_t + p0;
                ov[0] = v;
                ot[0] = 0.0;
                int nout = 1;
                for (int d = 0; d < D && !(flag & SMJP_ST_OVERFLOW); ++d) {
                    const double t_prev = d == 0 ? 0.0 : a.obs_t[o0 + d - 1];
                    const int rc = pf_propagate(a, m, chain, (uint32_t)d, (uint32_t)line[d], v, l, t_prev,
                                                a.obs_t[o0 + d], ov, ot, cap - 1, &nout);
                    if (rc) flag |= rc;
                }
                if (a.extend && !(flag & SMJP_ST_OVERFLOW)) {
                    const double t_prev = D > 0 ? a.obs_t[o0 + D - 1] : 0.0;
                    const int rc = pf_propagate(a, m, chain, (uint32_t)D, (uint32_t)(D > 0 ? line[D - 1] : 0), v, l, t_prev,
                                                a.t_end, ov, ot, cap - 1, &nout);
                    if (rc) flag |= rc;
                }
                if (!(flag & SMJP_ST_OVERFLOW)) {
                    ov[nout] = v;  // pmcmc.py:307-308: (v_last, t_final)
                    ot[nout] = a.t_end;
                    plen = nout + 1;
                }
            }
            a.path_len[c] = plen;
            a.loglik[c] = (flag & SMJP_ST_DEGENERATE) ? -INFINITY : loglik;
            a.status[c] = flag;
        }
        if (CS > 1) cluster.sync();  // nobody leaves (or reuses its shared memory) while rank 0 reads the flags
        else __syncthreads();
    }
}

}  // namespace smjp
