// FFBS kernel for larger state spaces (K >= 4): one thread per augmented STATE (v, j) instead of one
// thread per grid point j.
//
// Same arithmetic, same outputs and the same reference lines as ffbs.cuh (likelihood_and_decoding_hmm
// fast_hmm.py:202-335, backward_sampling_hmm :104-191, smjp_utils.py:319-330, mcmc_utils.py:40-56).
// With K*K hazards per grid pair the thread-per-j mapping keeps K*K transcendentals and their
// accumulators live per thread (K = 10: 255 registers and spills, 2 resident CTAs); here a thread owns
// ONE source state v = tid % K for all its grid points, so it evaluates only the K hazards A_vs of that
// row, K*(i+1) work items exist per step instead of i+1, and the register footprint is O(K).
// The block size is a multiple of K (and of 32), so v is a per-thread constant.
// Scratch rows are stored [j][v] (the work-item order: fully coalesced); exported messages keep the
// reference's state-major layout [v][j].
#pragma once
#include "ffbs.cuh"

namespace smjp {

template <typename Real, int K, int B, int MINB>
__global__ void __launch_bounds__(B, MINB) ffbs_item_kernel(const FfbsArgs a) {
    static_assert(B % K == 0 && B % 32 == 0, "block size must be a multiple of K and of the warp size");
    using M = Math<Real>;
    using Acc = typename ScanAcc<Real>::type;
    constexpr int NW = B / 32;
    constexpr int JPS = B / K;  // grid points per slot
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.guard && (a.guard[0] > a.guard_nmax || a.guard[1] > a.guard_total)) return;  // mis-speculated geometry
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int v = tid % K, jt = tid / K;
    const int ncap = a.ncap, Kobs = a.Kobs;
    const size_t state = (size_t)a.nslot * B;

    unsigned char* sp = smem_raw;
    double* Wd = (double*)sp;
    sp += align_up((size_t)(ncap + 1) * sizeof(double), 16);
    double* ubuf = (double*)sp;
    sp += align_up((size_t)B * sizeof(double), 16);
    Acc* scan_s = (Acc*)sp;
    sp += align_up((size_t)(NW + 1) * sizeof(double), 16);
    Real* st_a = (Real*)sp;  // forward: unnormalised message [slot][B]; backward: weights [K][len]
    sp += align_up(state * sizeof(Real), 16);
    Real* st_h = (Real*)sp;  // forward: H_v(tau) [slot][B]; backward: sojourn list
    sp += align_up(state * sizeof(Real), 16);
    Real* red = (Real*)sp;
    sp += align_up((size_t)2 * NW * (K + 1) * sizeof(Real), 16);
    Real* srk = (Real*)sp;
    Real* skm1 = srk + K * K;
    Real* sc2p = skm1 + K * K;
    sp += align_up((size_t)(4 * K * K) * sizeof(Real), 16);
    int* jlo_arr = (int*)sp;  // [ncap + 1] first live grid point of every row (scratch mode), see ffbs.cuh
    int* soj = (int*)st_h;

    __shared__ int s_chain;
    __shared__ int s_sel;
    __shared__ int s_found;
    __shared__ int s_jmin[2 * NW];
    __shared__ int s_flag;
    __shared__ double s_tab[Exp2<Real>::TAB ? Exp2<Real>::TAB : 1];
    __shared__ double2 s_ltab[Exp2<Real>::LTAB ? Exp2<Real>::LTAB : 1];

    for (int t = tid; t < K * K; t += B) {
        const double k = a.model[t];
        srk[t] = (Real)(1.0 / k);
        skm1[t] = (Real)((k - 1.0) * (double)Exp2<Real>::SCALE);
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
    // the K hazards of row v are read from shared memory (v is a per-thread constant)
    // (measured K=10 fp32: register copies cost occupancy and were 4 % slower than shared-memory reads)
    constexpr bool REGP = false;
    __syncthreads();
    Real r_rk[REGP ? K : 1], r_km1[REGP ? K : 1], r_c2p[REGP ? K : 1];
    if (REGP) {
#pragma unroll
        for (int s = 0; s < K; ++s) {
            r_rk[s] = srk[v * K + s];
            r_km1[s] = skm1[v * K + s];
            r_c2p[s] = sc2p[v * K + s];
        }
    }
    const Real* s_rk = srk + v * K;
    const Real* s_km1 = skm1 + v * K;
    const Real* s_c2p = sc2p + v * K;
#define P_RK(s) (REGP ? r_rk[REGP ? (s) : 0] : s_rk[s])
#define P_KM1(s) (REGP ? r_km1[REGP ? (s) : 0] : s_km1[s])
#define P_C2P(s) (REGP ? r_c2p[REGP ? (s) : 0] : s_c2p[s])

    const Real thin = (Real)(1.0 - 1.0 / a.omega);
    const Real omega_r = (Real)a.omega;
    const Real om_l2e = (Real)(a.omega * SMJP_LOG2E * (double)Exp2<Real>::SCALE);
    const Real power = (Real)a.power;
    Real* obsf = (Real*)a.obsf_scratch + (int64_t)blockIdx.x * ncap * K;
    const bool exporting = a.messages != nullptr;

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
        for (int t = tid; t <= n; t += B) Wd[t] = a.W[w0 + t];
        __syncthreads();
        {
            int bad = 0;
            for (int t = tid; t < n; t += B) bad |= !(Wd[t + 1] > Wd[t]);
            if (bad) atomicOr(&s_flag, 1);
        }
        {   // observation factor per interval (same as ffbs.cuh)
            const int64_t o0 = a.obs_offs[c];
            const int D = (int)(a.obs_offs[c + 1] - o0);
            const double* ot = a.obs_t + o0;
            const int32_t* oy = a.obs_y + o0;
            for (int i = tid; i < n; i += B) {
                const double wc = Wd[i], wn = Wd[i + 1];
                int lo = 0, hi = D;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (ot[mid] < wc) lo = mid + 1; else hi = mid;
                }
                Real f[K];
#pragma unroll
                for (int u = 0; u < K; ++u) f[u] = a.obs_product ? (Real)1 : (Real)0;
                int cnt = 0;
                for (int d = lo; d < D && ot[d] <= wn; ++d, ++cnt) {
                    int y = oy[d];
                    if ((unsigned)y >= (unsigned)Kobs) {  // never index the emission table out of bounds
                        y = 0;
                        atomicOr(&s_flag, 2);
                    }
#pragma unroll
                    for (int u = 0; u < K; ++u) {
                        const Real pe = (Real)emis[u * Kobs + y];
                        f[u] = a.obs_product ? f[u] * pe : f[u] + pe;
                    }
                }
#pragma unroll
                for (int u = 0; u < K; ++u)
                    obsf[i * K + u] = (cnt == 0 || power == (Real)0) ? (Real)1
                                      : (power == (Real)1 ? f[u] : M::pow(f[u], power));
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

        Real* msg = exporting ? (Real*)a.messages + a.msg_offs[c]
                              : (Real*)a.msg_scratch + (int64_t)blockIdx.x * a.msg_scratch_stride;

        // =========================== forward ===========================
        Real invZ_prev = (Real)1;
        Real Jv = (Real)(1.0 / K);  // inflow into (v, W[i]); a thread only ever needs its own state's
        Real of_v = obsf[v];
        bool degenerate = false;
        int i_final = n;
        while (i_final > 0 && np_isclose(a.t_end, Wd[i_final])) --i_final;
        Real* row_base = msg;  // start of row i-1
        // exact live window (ffbs.cuh): grid points below j_lo hold exact zeros for every state and are skipped
        int j_lo = 0;
        unsigned long long live = 0;
        const Real weps = (Real)a.window_eps;
        for (int i = 0; i < n; ++i) {
            const double wn = Wd[i + 1];
            const bool nonfinal = i < i_final;
            const int m_lo = j_lo / JPS, t_lo = j_lo - m_lo * JPS;
            if (tid == 0) jlo_arr[i] = j_lo;
            live += (unsigned long long)(i - j_lo + 1);
            const bool judge = !exporting && (i & 3) == 0;
            if (judge) {  // smallest grid point of the oldest slot with a state above zero (or above eps * Z_{i-1})
                int mylive = 0x7fffffff;
                const int j0 = m_lo * JPS + jt;
                if (j0 >= j_lo && j0 < i) {
                    const Real thr = weps > (Real)0 ? weps / invZ_prev : (Real)0;
                    if (st_a[(size_t)m_lo * B + tid] > thr) mylive = j0;
                }
                mylive = __reduce_min_sync(0xffffffffu, mylive);
                if (lane == 0) s_jmin[(i & 1) * NW + wid] = mylive;
            }
            const Real ofn_v = (i + 1 < n) ? obsf[(i + 1) * K + v] : (Real)1;
            Real Zp = 0;
            Real Jp[K];
#pragma unroll
            for (int u = 0; u < K; ++u) Jp[u] = 0;
            const int mi = i / JPS, ji = i - mi * JPS;  // slot of the new states (., W[i]); their jt
            Real* pa = st_a + (size_t)m_lo * B + tid;
            Real* ph = st_h + (size_t)m_lo * B + tid;
            const double* pw = Wd + m_lo * JPS + jt;
            // one augmented state (v, j): K hazards of row v, emission, thinned update, jump contributions
            auto item = [&](auto partial_tag, bool diag, int j) {
                constexpr bool PARTIAL = decltype(partial_tag)::value;
                const Real tau = (Real)(wn - *pw);
                const Real L2 = ex2.log2(tau);
                const Real ap = *pa * invZ_prev;
                Real hprev = *ph;
                Real base = ap * thin;
                if (exporting && (!PARTIAL || !diag)) row_base[(int64_t)v * i + j] = ap;  // row i-1, [v][j]
                if (PARTIAL) {
                    base = diag ? Jv : base;
                    hprev = diag ? (Real)0 : hprev;
                }
                Real kE[K];
                Real esum = 0, dsum = 0;
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    kE[s] = ex2(P_KM1(s) * L2 - P_C2P(s));  // A_vs(tau)
                    esum += kE[s] * P_RK(s);
                    dsum += kE[s];
                }
                const Real hsum = tau * esum;
                const Real g = of_v * ex2.neg(-om_l2e * (hsum - hprev)) * base;
                // scratch rows hold g = alpha / (Omega A_v) of row i itself, [j][v] (this thread's item order):
                // the backward jump weight alpha A_{v,v+} / (Omega A_v) is then g A_{v,v+}, one hazard per entry
                if (!exporting) row_base[(int64_t)K * i + (int64_t)j * K + v] = g;
                Real gj = g;
                // a final row (right end = t_end, smjp_utils.py:544-547) holds alpha = g WITHOUT the Omega A_v factor, so the
                // 1/(Omega A_v) of the jump term (smjp_utils.py:409-412) does not cancel there (matters when a grid point lies
                // within isclose() of t_end: two or more final rows)
                if (!nonfinal) gj = dsum > (Real)0 ? g * M::rcp(omega_r * dsum) : (Real)0;
#pragma unroll
                for (int s = 0; s < K; ++s) Jp[s] += gj * kE[s];
                const Real an = g * (nonfinal ? omega_r * dsum : (Real)1);
                *pa = an;
                *ph = hsum;
                Zp += an;
            };
            int j = m_lo * JPS + jt;
            int m = m_lo;
            if (m < mi) {  // oldest live slot: grid points below the window start are dead
                if (jt >= t_lo) item(std::false_type(), false, j);
                pa += B;
                ph += B;
                pw += JPS;
                j += JPS;
                ++m;
            }
            for (; m < mi; ++m) {  // full slots: j < i for every thread
                item(std::false_type(), false, j);
                pa += B;
                ph += B;
                pw += JPS;
                j += JPS;
            }
            if (jt <= ji && (mi > m_lo || jt >= t_lo)) item(std::true_type(), jt == ji, j);
            // block reduction of (Z, J[0..K))
            Real* rb = red + (i & 1) * NW * (K + 1);
            {
                constexpr int NVP = pow2_ceil(K + 1), SH = 5 - log2_of(NVP);
                Real vals[NVP];
                vals[0] = Zp;
#pragma unroll
                for (int u = 0; u < K; ++u) vals[1 + u] = Jp[u];
#pragma unroll
                for (int u = K + 1; u < NVP; ++u) vals[u] = 0;
                WarpHalve<Real, NVP, 16>::run(vals, lane);
                const int idx = lane >> SH;
                if ((lane & ((1 << SH) - 1)) == 0 && idx < K + 1) rb[wid * (K + 1) + idx] = vals[0];
            }
            __syncthreads();
            Real Z = 0;
            Jv = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                Z += rb[w * (K + 1)];
                Jv += rb[w * (K + 1) + 1 + v];
            }
            if (!(Z > (Real)0) || !(Z < (Real)INFINITY)) {
                degenerate = true;
                break;
            }
            invZ_prev = (Real)1 / Z;
            Jv *= invZ_prev;
            of_v = ofn_v;
            if (a.logZ && tid == 0) a.logZ[w0 + i] = log((double)Z);
            if (i > 0) row_base += (int64_t)K * i;
            if (judge) {
                int jl = 0x7fffffff;
#pragma unroll
                for (int w = 0; w < NW; ++w) jl = min(jl, s_jmin[(i & 1) * NW + w]);
                j_lo = max(j_lo, min(jl, min((m_lo + 1) * JPS, i)));
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
        if (exporting) {  // last row (scratch rows were complete when their step ended)
            Real* row_last = msg + (int64_t)K * (n - 1) * n / 2;
            for (int j = jt, m = 0; j < n; j += JPS, ++m)
                row_last[(int64_t)v * n + j] = st_a[(size_t)m * B + tid] * invZ_prev;
        }
        __syncthreads();

        // =========================== backward ===========================
        Real* wbuf = st_a;  // [K][len] state-major: the reference's categorical order a = v*len + j
        int cur = n - 1, vplus = -1, cnt = 0, u_hi = -1;
        while (true) {
            const int len = cur + 1;
            const Real* row = msg + (int64_t)K * cur * (cur + 1) / 2;
            const int lo = exporting ? 0 : jlo_arr[cur];  // live window of this row (entries below lo are exact zeros)
            const int wl = len - lo;
            {   // The next draw reads row j*-1 where W[j*] starts the sojourn found now; sojourns are a few grid
                // steps long, so rows cur-1 .. cur-3 cover ~7/8 of the cases: pull their live parts (one contiguous
                // [j][v] segment per scratch row) into L2 now, one draw ahead of the HBM round trip.
                const int r = cur - 1 - tid % 3, line = tid / 3;
                if (r >= 0) {
                    const int rlo = exporting ? 0 : jlo_arr[r];
                    const Real* rb0 = msg + (int64_t)K * r * (r + 1) / 2;
                    const char* p = (const char*)(rb0 + (int64_t)K * rlo) + (size_t)line * 128;
                    if (p < (const char*)(rb0 + (int64_t)K * (r + 1))) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                }
            }
            if (!a.uniforms && (u_hi < 0 || cur <= u_hi - B)) {
                if (cur - tid >= 0)
                    ubuf[tid] = philox_u01(a.seed, TAG_BACKWARD, (uint64_t)(cur - tid), 0u, a.sweep,
                                           a.chain_base + (uint32_t)c);
                u_hi = cur;
            }
            const double wn = Wd[cur + 1];
            // weights over the live window, state-major: wbuf[v * wl + (j - lo)]
            if (!exporting && (vplus < 0 || cur < i_final)) {
                // scratch row holds g, [j][v]: first draw w = g (= alpha on the last row), jump draws w = g A_{v,vplus}.
                // Unrolled so that the row loads of several entries are in flight before the first hazard.
                const Real km1 = vplus >= 0 ? s_km1[vplus] : (Real)0, c2p = vplus >= 0 ? s_c2p[vplus] : (Real)0;
#pragma unroll 4
                for (int j = lo + jt; j < len; j += JPS) {
                    Real w = row[(int64_t)j * K + v];
                    if (vplus >= 0) w *= ex2(km1 * ex2.log2((Real)(wn - Wd[j])) - c2p);
                    wbuf[v * wl + (j - lo)] = w;
                }
            } else {
                for (int j = lo + jt; j < len; j += JPS) {
                    const int64_t pos = exporting ? (int64_t)v * len + j : (int64_t)j * K + v;
                    Real w = row[pos];
                    if (vplus >= 0) {  // jump at W[cur+1] into vplus: weight alpha_cur(v,j) A_{v,vplus}/A_v
                        const Real L2 = ex2.log2((Real)(wn - Wd[j]));
                        Real dsum = 0, hto = 0;
#pragma unroll
                        for (int s = 0; s < K; ++s) {
                            const Real kE = ex2(P_KM1(s) * L2 - P_C2P(s));
                            dsum += kE;
                            hto = (s == vplus) ? kE : hto;
                        }
                        w *= dsum > (Real)0 ? hto * M::rcp(dsum) : (Real)0;
                    }
                    wbuf[v * wl + (j - lo)] = w;
                }
            }
            __syncthreads();
            const int L = K * wl;
            const int q = (L + B - 1) / B;
            const int beg = min(tid * q, L), end = min(beg + q, L);
            Acc loc = 0;
            for (int t = beg; t < end; ++t) loc += (Acc)wbuf[t];
            Acc incl = warp_inclusive_scan(loc, lane);
            if (lane == 31) scan_s[wid + 1] = incl;
            if (tid == 0) { scan_s[0] = 0; s_found = 0; }
            __syncthreads();
            Acc off = 0;
            for (int w = 1; w <= wid; ++w) off += scan_s[w];
            incl += off;
            Acc excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = off;
            Acc total = 0;
#pragma unroll
            for (int w = 1; w <= NW; ++w) total += scan_s[w];
            if (!(total > (Acc)0) || !(total < (Acc)INFINITY)) {
                degenerate = true;
                break;
            }
            const double u = a.uniforms ? a.uniforms[w0 + cur] : ubuf[u_hi - cur];
            const Acc target = (Acc)u * total;
            const bool has = end > beg && incl > excl;  // empty chunks never own a draw (see below)
            if (has && target >= excl && target < incl) {
                Acc acc = excl;
                int sel = -1, lastpos = beg;
                for (int t = beg; t < end; ++t) {
                    const Acc wv = (Acc)wbuf[t];
                    if (wv > (Acc)0) lastpos = t;
                    acc += wv;
                    if (acc > target) { sel = t; break; }
                }
                s_sel = sel >= 0 ? sel : lastpos;
                s_found = 1;
            }
            __syncthreads();
            // No owner: u*total rounded up to the total, or the target fell between the inclusive value of the
            // last thread that holds weights and the warp total (a shuffle scan over zero-padded lanes is not
            // monotone to the last bit -- in fp32 that happens about once per 1e7 draws).  The last thread with
            // mass then takes its last positive entry, as the reference's inverse CDF does at u -> 1.
            if (!s_found) {  // block-uniform
                if (has) atomicMax(&s_flag, tid + 1);
                __syncthreads();
                if (tid + 1 == s_flag) {
                    int lastpos = beg;
                    for (int t = beg; t < end; ++t) if ((Acc)wbuf[t] > (Acc)0) lastpos = t;
                    s_sel = lastpos;
                }
                __syncthreads();
                if (tid == 0) s_flag = 0;
            }
            const int idx = s_sel;
            const int vs = idx / wl, js = lo + idx - vs * wl;
            if (a.aug_idx)
                for (int t = js + tid; t <= cur; t += B) a.aug_idx[w0 + t] = vs * n + js;
            if (tid == 0) {
                soj[cnt] = vs;
                soj[(ncap + 1) + cnt] = js;
            }
            ++cnt;
            if (js == 0) break;
            vplus = vs;
            cur = js - 1;
        }
        __syncthreads();
        if (degenerate) {
            if (tid == 0) {
                a.status[c] = SMJP_ST_DEGENERATE;
                a.path_len[c] = 0;
            }
            continue;
        }
        for (int p = tid; p < cnt; p += B) {
            a.path_v[w0 + p] = soj[cnt - 1 - p];
            a.path_t[w0 + p] = Wd[soj[(ncap + 1) + cnt - 1 - p]];
        }
        const int v_last = soj[0];
        const double t_last = Wd[soj[ncap + 1]];
        const bool add_end = !np_isclose(t_last, a.t_end);
        const int plen = cnt + (add_end ? 1 : 0);
        if (tid == 0) {
            if (add_end) {
                a.path_v[w0 + cnt] = v_last;
                a.path_t[w0 + cnt] = a.t_end;
            }
            a.path_len[c] = plen;
            a.status[c] = 0;
        }
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

#undef P_RK
#undef P_KM1
#undef P_C2P

}  // namespace smjp
