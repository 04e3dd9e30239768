// FFBS for the shape == 1 collapse of the sMJP (a Markov jump process with self-jumps): the K-state dense
// path of SURVEY.md Appendix A.5, one warp per chain (the kernel handles K <= 128; smjp_set_model admits K <= 64).
//
// With every Weibull shape equal to 1 the hazards A_vs = 1/lambda_vs are constant, the emission of the
// augmented filter does not depend on l, and m_i(v) = sum_l alpha_i(v,l) obeys
//     m_i  ~  e_i o (m_{i-1} B),     B = (1 - 1/Omega) I + diag(1/(Omega A_v)) A        (constant K x K)
//     e_i(v) = obs_i(v) [Omega A_v]^{not final} exp(-Omega A_v (W[i+1] - W[i]))
// (reference: the same lines as ffbs.cuh -- fast_hmm.py:202-335 with smjp_utils.py:382-416, :506-550 --
// marginalised over l; hidden_markov_model.py:213-313 is its plain-HMM form).  Backward sampling draws
// v_i ~ m_i(v) B[v, v_{i+1}] and, when v_i == v_{i+1}, a Bernoulli between the thinned continuation
// (1 - 1/Omega) and the self-jump A_vv/(Omega A_v): jumps delimit the sojourns of the returned path.
// This is the path BASELINE configs[3] (K = 64, long horizon) runs on: the (v,l) kernels stop at K = 10.
// Lane s owns the SPL adjacent states SPL*s ..; B (and its transpose, and the emission table) live in shared memory; the
// message stream m_i goes to HBM (2 K sizeof(real) bytes per grid step) for the backward pass.  Grid steps
// are staged 32 at a time: lane L loads W, finds the observations and draws the Philox block of step i0 + L,
// so the serial recursion itself only sees shuffles, shared memory and the (prefetched) message row.
#pragma once
#include "../../include/smjp_b200.h"
#include "common.cuh"
#include "ffbs.cuh"

namespace smjp {

struct DenseArgs {
    int C, K, Kobs;
    const int64_t* w_offs;
    const double* W;
    const int64_t* obs_offs;
    const double* obs_t;
    const int32_t* obs_y;
    const double* u_state;  // per grid step, or null -> Philox (r = 2i)
    const double* u_jump;   // per grid step, or null -> Philox (r = 2i + 1)
    uint64_t seed;
    uint32_t sweep, chain_base;
    const double* model;    // shape[K*K], c2[K*K], emis[K*Kobs], scale[K*K]
    int obs_product;
    double omega, power, t_end;
    void* msg;              // [total_w][K] row-normalised messages (Real)
    double* logZ;           // per grid step or null
    int32_t* path_v;
    double* path_t;
    int32_t* path_len;
    double* time_in_state;
    int32_t* jumps;
    int32_t* status;
    int fwd_done;           // the message rows (and status / logZ) were produced by dense_tc.cuh: backward pass only
};

// shared-memory plan: Bs [Kr][KV] (rows = source state, zero padded), optional BT [K][KV] (B transposed, for
// the backward column reads) and emT [Kobs][KV] (emission, transposed) when they fit, rate [KV], pself [KV],
// and one message row [KV] per warp (the forward matvec reads it back as broadcast vectors)
constexpr int DENSE_WARPS = 4;
struct DensePlan {
    int KV, Kr, has_bt, has_em;
    size_t bytes;
};
inline DensePlan dense_plan(int K, int Kobs, size_t esz) {
    DensePlan p;
    p.KV = ((K + 31) / 32) * 32;
    p.Kr = (K + 3) & ~3;
    const size_t budget = 200 * 1024;
    size_t b = ((size_t)p.Kr * p.KV + (2 + DENSE_WARPS) * p.KV) * esz;
    p.has_bt = b + (size_t)K * p.KV * esz <= budget;
    if (p.has_bt) b += (size_t)K * p.KV * esz;
    p.has_em = b + (size_t)Kobs * p.KV * esz <= budget;
    if (p.has_em) b += (size_t)Kobs * p.KV * esz;
    p.bytes = b;
    return p;
}

template <typename Real, int SPL>
__global__ void __launch_bounds__(DENSE_WARPS * 32, 1) dense_smjp_kernel(const DenseArgs a, const int has_bt, const int has_em) {
    extern __shared__ __align__(16) unsigned char dsm[];
    constexpr int KV = 32 * SPL;
    constexpr unsigned FULL = 0xffffffffu;
    const int K = a.K, Kobs = a.Kobs, Kr = (K + 3) & ~3;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, NW = blockDim.x >> 5;
    Real* Bs = (Real*)dsm;                          // [Kr][KV]  B[v'][v]
    Real* rate_s = Bs + Kr * KV;                    // [KV] Omega A_v
    Real* pself_s = rate_s + KV;                    // [KV] P(self-jump | stayed in v)
    Real* mrow = pself_s + KV + wid * KV;           // [KV] this warp's current (unnormalised) message
    Real* BT = pself_s + KV + DENSE_WARPS * KV;     // [K][KV]   B[v][v_next] stored as BT[v_next][v]
    Real* emT = BT + (has_bt ? K * KV : 0);         // [Kobs][KV]
    const double* scale = a.model + 2 * K * K + K * Kobs;
    const double* emis = a.model + 2 * K * K;
    __shared__ double Av_d[128];
    for (int v = threadIdx.x; v < KV; v += blockDim.x) {
        double s = 0, ps = 0;
        if (v < K) {
            for (int t = 0; t < K; ++t) s += 1.0 / scale[v * K + t];
            const double self = (1.0 / scale[v * K + v]) / (a.omega * s);
            ps = self / ((1.0 - 1.0 / a.omega) + self);
        }
        Av_d[v] = s;
        rate_s[v] = (Real)(a.omega * s);
        pself_s[v] = (Real)ps;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < Kr * KV; t += blockDim.x) {
        const int vp = t / KV, v = t - vp * KV;
        double b = 0;
        if (vp < K && v < K) b = (vp == v ? 1.0 - 1.0 / a.omega : 0.0) + (1.0 / scale[vp * K + v]) / (a.omega * Av_d[vp]);
        Bs[t] = (Real)b;
        if (has_bt && vp < K && v < K) BT[v * KV + vp] = (Real)b;  // BT[v_next][v]; pad columns are zeroed below
    }
    if (has_bt)
        for (int t = threadIdx.x; t < K * (KV - K); t += blockDim.x) BT[(t / (KV - K)) * KV + K + t % (KV - K)] = (Real)0;
    if (has_em)
        for (int t = threadIdx.x; t < Kobs * KV; t += blockDim.x) {
            const int y = t / KV, v = t - y * KV;
            emT[t] = v < K ? (Real)emis[v * Kobs + y] : (Real)0;
        }
    __syncthreads();
    double rmin_d = 1e300;
    for (int v = 0; v < K; ++v) rmin_d = fmin(rmin_d, a.omega * Av_d[v]);
    const Real power = (Real)a.power;
    using M = Math<Real>;
    // per-lane constants of the owned states v = SPL * lane + k (adjacent: vector shared loads, one warp scan)
    const int vbase = SPL * lane;
    Real rt[SPL], fin[SPL], rd[SPL];
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const int v = vbase + k;
        rt[k] = v < K ? rate_s[v] : (Real)0;
        fin[k] = v < K ? (Real)1 : (Real)0;
        // exponent relative to the slowest state keeps fp32 away from underflow; rmin goes back into logZ
        rd[k] = v < K ? (Real)(-(double)SMJP_LOG2E * (a.omega * Av_d[v] - rmin_d)) : (Real)0;
    }

    // fp32, K <= 64: the lane's columns of B stay in registers (32 SPL^2 of them), so the recursion reads only
    // the broadcast message row from shared memory; otherwise every warp streams all of B per grid step
    constexpr bool REGB = sizeof(Real) == 4 && SPL <= 2;
    Real Breg[REGB ? KV : 1][SPL];
    if constexpr (REGB) {
#pragma unroll
        for (int vp = 0; vp < KV; ++vp) {
#pragma unroll
            for (int k = 0; k < SPL; ++k) Breg[vp][k] = vp < Kr ? Bs[vp * KV + vbase + k] : (Real)0;
        }
    }

    for (int c = blockIdx.x * NW + wid; c < a.C; c += gridDim.x * NW) {
        const int64_t w0 = a.w_offs[c];
        const int n = (int)(a.w_offs[c + 1] - w0) - 1;
        const double* W = a.W + w0;
        if (n < 1) {
            if (lane == 0) { a.status[c] = SMJP_ST_GRID_NOT_INCREASING; a.path_len[c] = 0; }
            continue;
        }
        Real* mh = (Real*)a.msg + w0 * K;
        const int64_t o0 = a.obs_offs[c];
        const int D = (int)(a.obs_offs[c + 1] - o0);
        const double* ot = a.obs_t + o0;
        const int32_t* oy = a.obs_y + o0;
        int i_final = n;
        while (i_final > 0 && np_isclose(a.t_end, W[i_final])) --i_final;
        // ---------------- forward, 32 grid steps per block: lane L stages step i0 + L ----------------
        // The row kept in shared memory and HBM is r_i = z_i m_i (normalised by the PREVIOUS step's sum only):
        // sum(r_i) is exactly the reference's normaliser z_i, and its warp reduction overlaps the next matvec.
        Real iz = 1;  // 1 / sum(r_{i-1})
        int flag = a.fwd_done ? a.status[c] : 0, dbase = 0;
        for (int i0 = 0; i0 < n && !flag && !a.fwd_done; i0 += 32) {
            const int il = i0 + lane;
            const bool live = il < n;
            const double wc_l = W[min(il, n)], wn_l = W[min(il + 1, n)];
            if (__ballot_sync(FULL, live && !(wn_l > wc_l))) { flag = SMJP_ST_GRID_NOT_INCREASING; break; }
            const Real dw_l = (Real)(wn_l - wc_l);
            // observations of the closed interval [wc, wn] (smjp_utils.py:40-48): the sorted record is read in
            // 32-wide windows from dbase (every earlier observation precedes W[i0]) and counted by shuffles
            int dlo_l = dbase, cnt_l = 0;
            if (D > 0) {
                const double wend = __shfl_sync(FULL, wn_l, 31);
                int below = 0, upto = 0;
                for (int off = dbase; off < D; off += 32) {
                    const double t_l = off + lane < D ? ot[off + lane] : (double)INFINITY;
#pragma unroll 8
                    for (int q = 0; q < 32; ++q) {
                        const double tq = __shfl_sync(FULL, t_l, q);
                        below += tq < wc_l;
                        upto += tq <= wn_l;
                    }
                    if (!(__shfl_sync(FULL, t_l, 31) <= wend)) break;
                }
                dlo_l = dbase + below;
                cnt_l = live ? upto - below : 0;
                dbase = __shfl_sync(FULL, dlo_l, 31);
            }
            const int y_l = cnt_l > 0 ? oy[dlo_l] : 0;
            Real z_l = 1;
            const int steps = min(32, n - i0);
            for (int s = 0; s < steps; ++s) {
                const int i = i0 + s;
                const Real dw = __shfl_sync(FULL, dw_l, s);
                const int cnt = __shfl_sync(FULL, cnt_l, s);
                const bool nonfinal = i < i_final;
                Real nw[SPL];
#pragma unroll
                for (int k = 0; k < SPL; ++k) nw[k] = (nonfinal ? rt[k] : fin[k]) * M::exp2(rd[k] * dw);
                if (cnt > 0) {
                    const int y0 = __shfl_sync(FULL, y_l, s);
                    const int dlo = __shfl_sync(FULL, dlo_l, s);
                    Real f[SPL];
#pragma unroll
                    for (int k = 0; k < SPL; ++k) f[k] = a.obs_product ? (Real)1 : (Real)0;
                    for (int d = 0; d < cnt; ++d) {
                        const int y = d == 0 ? y0 : oy[dlo + d];
                        Real pe[SPL];
                        if (has_em) lds_vec<Real, SPL>(emT + y * KV + vbase, pe);
#pragma unroll
                        for (int k = 0; k < SPL; ++k) {
                            if (!has_em) pe[k] = vbase + k < K ? (Real)emis[(vbase + k) * Kobs + y] : (Real)0;
                            f[k] = a.obs_product ? f[k] * pe[k] : f[k] + pe[k];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < SPL; ++k)
                        nw[k] *= power == (Real)0 ? (Real)1 : (power == (Real)1 ? f[k] : M::pow(f[k], power));
                }
                if (i == 0) {
#pragma unroll
                    for (int k = 0; k < SPL; ++k) nw[k] *= (Real)(1.0 / K);  // uniform over (v, 0): smjp_utils.py:286-289
                } else {
                    Real acc[SPL][2];
#pragma unroll
                    for (int k = 0; k < SPL; ++k) acc[k][0] = acc[k][1] = 0;
                    if constexpr (REGB) {
#pragma unroll
                        for (int vp = 0; vp < KV; vp += 4) {  // entries >= K of mrow are zero
                            Real m4[4];
                            lds_vec<Real, 4>(mrow + vp, m4);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
#pragma unroll
                                for (int k = 0; k < SPL; ++k) acc[k][q & 1] += m4[q] * Breg[vp + q][k];
                            }
                        }
                    } else {
                        const Real* bcol = Bs + vbase;
#pragma unroll 4
                        for (int vp = 0; vp < Kr; vp += 4) {  // rows >= K of Bs and entries >= K of mrow are zero
                            Real m4[4];
                            lds_vec<Real, 4>(mrow + vp, m4);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                Real b[SPL];
                                lds_vec<Real, SPL>(bcol + (vp + q) * KV, b);
#pragma unroll
                                for (int k = 0; k < SPL; ++k) acc[k][q & 1] += m4[q] * b[k];
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < SPL; ++k) nw[k] *= (acc[k][0] + acc[k][1]) * iz;
                    __syncwarp();  // every lane has read mrow before it is overwritten
                }
                {
                    if constexpr (SPL == 2 || SPL == 4) {
#pragma unroll
                        for (int k = 0; k < SPL; k += 2) {
                            if constexpr (sizeof(Real) == 4) *reinterpret_cast<float2*>(mrow + vbase + k) = make_float2(nw[k], nw[k + 1]);
                            else *reinterpret_cast<double2*>(mrow + vbase + k) = make_double2(nw[k], nw[k + 1]);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < SPL; ++k) mrow[vbase + k] = nw[k];
                    }
                }
#pragma unroll
                for (int k = 0; k < SPL; ++k)
                    if (vbase + k < K) mh[(int64_t)i * K + vbase + k] = nw[k];
                Real z = 0;
#pragma unroll
                for (int k = 0; k < SPL; ++k) z += nw[k];
                z = warp_sum(z);  // also orders the mrow stores before the next step's loads (shuffles synchronise)
                __syncwarp();
                if (!(z > (Real)0) || !(z < (Real)INFINITY)) { flag = SMJP_ST_DEGENERATE; break; }
                iz = (Real)1 / z;
                if (lane == s) z_l = z;
            }
            if (a.logZ && live && !flag) a.logZ[w0 + il] = log((double)z_l) - rmin_d * (wn_l - wc_l);
        }
        if (flag) {
            if (lane == 0) { a.status[c] = flag; a.path_len[c] = 0; }
            continue;
        }
        __syncwarp();
        // ---------------- backward: states, jump flags, path written last-to-first ----------------
        int32_t* pv = a.path_v + w0;
        double* pt = a.path_t + w0;
        const uint32_t chain = a.chain_base + (uint32_t)c;
        int v_next = -1, cnt = 0;
        Real cur[SPL], nxt[SPL];
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            cur[k] = vbase + k < K ? mh[(int64_t)(n - 1) * K + vbase + k] : (Real)0;
            nxt[k] = 0;
        }
        for (int i0 = ((n - 1) >> 5) << 5; i0 >= 0 && !flag; i0 -= 32) {
            const int il = i0 + lane;
            const bool live = il < n;
            const double wn_l = W[min(il + 1, n)];
            double us_l = 0, uj_l = 0;
            if (live) {
                if (a.u_state) {
                    us_l = a.u_state[w0 + il];
                    uj_l = a.u_jump[w0 + il];
                } else {  // one Philox block per grid step: words 0,1 -> state draw (r = 2i), 2,3 -> jump draw (r = 2i+1)
                    const Philox4 o = philox4x32_10((uint32_t)il, 0u, a.sweep, chain, (uint32_t)a.seed,
                                                    (uint32_t)(a.seed >> 32) ^ TAG_BACKWARD);
                    us_l = uniform53(o.x[0], o.x[1]);
                    uj_l = uniform53(o.x[2], o.x[3]);
                }
            }
            const Real usr_l = (Real)us_l;
            for (int s = min(31, n - 1 - i0); s >= 0; --s) {
                const int i = i0 + s;
                if (i > 0) {
#pragma unroll
                    for (int k = 0; k < SPL; ++k) nxt[k] = vbase + k < K ? mh[(int64_t)(i - 1) * K + vbase + k] : (Real)0;
                }
                // rows are read last-to-first, one per draw: only the newest are still in L2, so pull row i-8
                // towards the SM now (the draw itself does not depend on it)
                if (i >= 8 && vbase < K) asm volatile("prefetch.global.L1 [%0];" ::"l"(mh + (int64_t)(i - 8) * K + vbase));
                Real w[SPL];
#pragma unroll
                for (int k = 0; k < SPL; ++k) w[k] = cur[k];
                if (v_next >= 0) {
                    Real bt[SPL];
                    if (has_bt) lds_vec<Real, SPL>(BT + v_next * KV + vbase, bt);
#pragma unroll
                    for (int k = 0; k < SPL; ++k) {
                        if (!has_bt) bt[k] = vbase + k < K ? Bs[(vbase + k) * KV + v_next] : (Real)0;
                        w[k] *= bt[k];
                    }
                }
                Real pre[SPL];  // inclusive prefix over the lane's own states, then one warp scan of lane totals
                pre[0] = w[0];
#pragma unroll
                for (int k = 1; k < SPL; ++k) pre[k] = pre[k - 1] + w[k];
                const Real incl_lane = warp_inclusive_scan(pre[SPL - 1], lane);
                const Real excl = incl_lane - pre[SPL - 1];
                const Real target = __shfl_sync(FULL, usr_l, s) * __shfl_sync(FULL, incl_lane, 31);
                int cand_l = -1, last_l = -1;
#pragma unroll
                for (int k = SPL - 1; k >= 0; --k) {
                    if (w[k] > (Real)0) {
                        if (last_l < 0) last_l = vbase + k;
                        if (excl + pre[k] > target) cand_l = vbase + k;
                    }
                }
                const unsigned hit = __ballot_sync(FULL, cand_l >= 0);
                const unsigned pos = __ballot_sync(FULL, last_l >= 0);
                if (!pos) { flag = SMJP_ST_DEGENERATE; break; }
                const int pick = hit ? __shfl_sync(FULL, cand_l, __ffs(hit) - 1) : __shfl_sync(FULL, last_l, 31 - __clz(pos));
                // did the transition i -> i+1 jump?  (a jump at W[i+1] starts a sojourn in v_next there)
                if (v_next >= 0) {
                    bool jump = pick != v_next;
                    if (!jump) jump = __shfl_sync(FULL, uj_l, s) < (double)pself_s[pick];
                    if (jump) {
                        if (lane == s) { pv[n - cnt] = v_next; pt[n - cnt] = wn_l; }
                        ++cnt;
                    }
                }
                v_next = pick;
#pragma unroll
                for (int k = 0; k < SPL; ++k) cur[k] = nxt[k];
            }
        }
        if (flag) {
            if (lane == 0) { a.status[c] = flag; a.path_len[c] = 0; }
            continue;
        }
        if (lane == 0) { pv[n - cnt] = v_next; pt[n - cnt] = 0.0; }  // the sojourn that starts at time 0
        ++cnt;
        __syncwarp();
        // entries sit at [n-cnt+1 .. n] in time order; move them to the front (dst <= src, chunked by warp)
        const int src0 = n - cnt + 1;
        for (int base = 0; base < cnt; base += 32) {
            const int p = base + lane;
            int32_t tv = 0;
            double tt = 0;
            if (p < cnt) { tv = pv[src0 + p]; tt = pt[src0 + p]; }
            __syncwarp();
            if (p < cnt) { pv[p] = tv; pt[p] = tt; }
            __syncwarp();
        }
        const int v_last = pv[cnt - 1];
        const double t_last = pt[cnt - 1];
        const bool endp = !np_isclose(t_last, a.t_end);  // smjp_utils.py:338-348
        const int plen = cnt + (endp ? 1 : 0);
        if (lane == 0) {
            if (endp) { pv[cnt] = v_last; pt[cnt] = a.t_end; }
            a.path_len[c] = plen;
            a.status[c] = 0;
        }
        __syncwarp();
        // chain metrics (mcmc_utils.py:40-56): each lane accumulates its own states in path order
        if (a.time_in_state || a.jumps) {
            double tsum[SPL];
            int jc[SPL];
#pragma unroll
            for (int k = 0; k < SPL; ++k) { tsum[k] = 0; jc[k] = 0; }
            for (int p = 0; p < plen; ++p) {
                const int k = pv[p] - vbase;
                const double dt = p + 1 < plen ? pt[p + 1] - pt[p] : 0.0;
#pragma unroll
                for (int q = 0; q < SPL; ++q)
                    if (k == q) { ++jc[q]; tsum[q] += dt; }
            }
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                if (vbase + k >= K) continue;
                if (a.time_in_state) a.time_in_state[(int64_t)c * K + vbase + k] = tsum[k];
                if (a.jumps) a.jumps[(int64_t)c * K + vbase + k] = jc[k];
            }
        }
    }
}

}  // namespace smjp
