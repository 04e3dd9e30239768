// FFBS kernel for LARGE state spaces with general Weibull shapes: K is a run-time value (10 < K <= 64).
//
// Same arithmetic, inputs, outputs and reference lines as ffbs.cuh (likelihood_and_decoding_hmm
// fast_hmm.py:202-335, backward_sampling_hmm :104-191, smjp_ffbs tail smjp_utils.py:319-330 / :338-348,
// chain metrics mcmc_utils.py:40-56); what changes is where the per-thread vectors live.  ffbs.cuh /
// ffbs_item.cuh keep K hazards and K inflow accumulators in registers and are instantiated per K; here
// thread t owns grid points j = t, t+B, ... and its K hazards kE[s] and K accumulators Jp[s] sit in shared
// memory columns ([s][B+1], conflict free), the hazard parameters are read from global memory through L1
// (every lane reads the same (v,s) entry), and the per-step reduction sums the columns row by row.  The chain
// state is 2 K reals per LIVE grid point (ffbs.cuh, "Live window": exact zeros are dropped; window_eps > 0 also
// drops states below that fraction of the row sum) and sits in a RING of `nslot` slots of B grid points -- as many
// as shared memory holds: K = 64 keeps a window of ~128 (fp64) / ~256 (fp32) grid points whatever the length of
// the grid (the first version held the whole grid: n <= 140 / 280).  The grid, the window start of every row and
// the sojourn list stay in global memory (L2).  A window that outgrows the ring ends the chain with
// SMJP_ST_OVERFLOW (the caller prunes: option window_log2).  Exported messages need the whole triangle and keep the
// old limit.  Scratch rows hold g (see ffbs.cuh, Backward).
#pragma once
#include "ffbs.cuh"

namespace smjp {

// shared memory for a ring of `slots` slots (ffbs_nslot(ncap, threads) slots hold the whole grid)
template <typename Real>
__host__ __device__ inline size_t ffbs_big_smem_bytes(int K, int slots, int threads) {
    const size_t state = (size_t)slots * K * threads;
    size_t b = 0;
    b += align_up((size_t)threads * sizeof(double), 16);          // backward uniforms
    b += align_up((size_t)(threads / 32 + 1) * sizeof(double), 16);  // scan totals
    b += 2 * align_up(state * sizeof(Real), 16);                  // message, H_v
    b += 2 * align_up((size_t)K * (threads + 1) * sizeof(Real), 16);  // Jp, kE columns
    b += align_up((size_t)(3 * K + 1 + threads) * sizeof(Real), 16);  // J, of, red[K+1], Zp[B]
    return b;
}

template <typename Real, int B>
__global__ void __launch_bounds__(B) ffbs_big_kernel(const FfbsArgs a) {
    using M = Math<Real>;
    using Acc = typename ScanAcc<Real>::type;
    constexpr int NW = B / 32, BP = B + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.guard && (a.guard[0] > a.guard_nmax || a.guard[1] > a.guard_total)) return;  // mis-speculated geometry
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int K = a.K, ncap = a.ncap, Kobs = a.Kobs;
    const int NS = a.nslot;  // ring slots of B grid points
    const size_t state = (size_t)NS * K * B;

    unsigned char* sp = smem_raw;
    double* ubuf = (double*)sp;
    sp += align_up((size_t)B * sizeof(double), 16);
    Acc* scan_s = (Acc*)sp;
    sp += align_up((size_t)(NW + 1) * sizeof(double), 16);
    Real* st_a = (Real*)sp;  // forward: unnormalised message, ring [slot][v][B]; backward: weights [K][window]
    sp += align_up(state * sizeof(Real), 16);
    Real* st_h = (Real*)sp;  // forward: H_v(tau), ring [slot][v][B]
    sp += align_up(state * sizeof(Real), 16);
    Real* Jp_s = (Real*)sp;  // [K][B+1] this step's inflow contributions, one column per thread
    sp += align_up((size_t)K * BP * sizeof(Real), 16);
    Real* kE_s = (Real*)sp;  // [K][B+1] hazards of the (v, j) being processed, one column per thread
    sp += align_up((size_t)K * BP * sizeof(Real), 16);
    Real* J_s = (Real*)sp;   // [K] normalised inflow into (., W[i])
    Real* of_s = J_s + K;    // [K] observation factors of the current step
    Real* red_s = of_s + K;  // [K+1] column sums, red_s[K] = Z
    Real* Zp_s = red_s + K + 1;  // [B]

    __shared__ int s_chain, s_sel, s_flag, s_found;
    __shared__ int s_jmin[2 * NW];
    __shared__ double s_tab[Exp2<Real>::TAB ? Exp2<Real>::TAB : 1];
    __shared__ double2 s_ltab[Exp2<Real>::LTAB ? Exp2<Real>::LTAB : 1];
    for (int t = tid; t < Exp2<Real>::TAB; t += B) s_tab[t] = exp2_table_entry(t);
    for (int t = tid; t < Exp2<Real>::LTAB; t += B) {
        const double rc = 1.0 / (1.0 + ((double)t + 0.5) / 128.0);
        s_ltab[t] = make_double2(rc, -::log2(rc));
    }
    Exp2<Real> ex2;
    ex2.tab = s_tab;
    ex2.ltab = s_ltab;
    // hazard parameters, prepared by the launcher in the kernel's precision: [3][K*K] = 1/k, (k-1) SCALE,
    // (k log2 lambda - log2 k) SCALE
    const Real* prk = (const Real*)a.big_params;
    const Real* pkm1 = prk + K * K;
    const Real* pc2p = pkm1 + K * K;
    const double* emis = a.model + 2 * K * K;

    const Real thin = (Real)(1.0 - 1.0 / a.omega);
    const Real omega_r = (Real)a.omega;
    const Real om_l2e = (Real)(a.omega * SMJP_LOG2E * (double)Exp2<Real>::SCALE);
    const Real power = (Real)a.power;
    // per-CTA global scratch: observation factors [ncap][K], window start of every row [ncap + 1], sojourn list
    // [2][ncap + 1]
    Real* obsf = (Real*)a.obsf_scratch + (int64_t)blockIdx.x * a.obsf_stride;
    int* jlo_arr = (int*)(obsf + (int64_t)ncap * K);
    int* soj = jlo_arr + (ncap + 1);
    const bool exporting = a.messages != nullptr;
    const Real weps = (Real)a.window_eps;

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
        const double* Wd = a.W + w0;  // the grid stays in global memory (L1 / L2)
        {
            int bad = 0;
            for (int t = tid; t < n; t += B) bad |= !(Wd[t + 1] > Wd[t]);
            if (bad) atomicOr(&s_flag, 1);
        }
        {   // observation factor per interval: (sum_{x in [W[i],W[i+1]]} P(x|v))^power
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
                int cnt = 0;
                while (lo + cnt < D && ot[lo + cnt] <= wn) ++cnt;
                for (int v = 0; v < K; ++v) {
                    Real f = a.obs_product ? (Real)1 : (Real)0;
                    for (int d = 0; d < cnt; ++d) {
                        int y = oy[lo + d];
                        if ((unsigned)y >= (unsigned)Kobs) {  // never index the emission table out of bounds
                            y = 0;
                            atomicOr(&s_flag, 2);
                        }
                        const Real pe = (Real)emis[v * Kobs + y];
                        f = a.obs_product ? f * pe : f + pe;
                    }
                    obsf[i * K + v] = (cnt == 0 || power == (Real)0) ? (Real)1 : (power == (Real)1 ? f : M::pow(f, power));
                }
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
        for (int v = tid; v < K; v += B) {
            J_s[v] = (Real)(1.0 / K);  // uniform prior on (v,0): smjp_utils.py:286-289
            of_s[v] = obsf[v];
        }
        bool degenerate = false;
        int i_final = n;
        while (i_final > 0 && np_isclose(a.t_end, Wd[i_final])) --i_final;
        Real* row_base = msg;  // start of row i-1 (row r holds K*(r+1) values)
        int j_lo = 0, ms_lo = 0;  // first live grid point (ffbs.cuh, "Live window"); ring slot of its slot
        bool overflow = false;
        unsigned long long live = 0;
        __syncthreads();
        for (int i = 0; i < n; ++i) {
            const double wn = Wd[i + 1];
            const bool nonfinal = i < i_final;
            for (int s = 0; s < K; ++s) Jp_s[s * BP + tid] = 0;
            Real Zp = 0;
            const int mi = i / B, ti = i - mi * B;
            const int m_lo = j_lo / B, t_lo = j_lo - m_lo * B;
            if (mi - m_lo >= NS) {  // the window outgrew the ring (block-uniform)
                overflow = true;
                break;
            }
            if (tid == 0) jlo_arr[i] = j_lo;
            live += (unsigned long long)(i - j_lo + 1);
            const bool judge = !exporting && (i & 3) == 0;  // the window start is re-examined every fourth step
            if (judge) {
                int mylive = 0x7fffffff;
                const int j0 = m_lo * B + tid;
                if (j0 >= j_lo && j0 < i) {
                    const Real thr = weps > (Real)0 ? weps / invZ_prev : (Real)0;
                    const Real* q = st_a + (size_t)ms_lo * K * B + tid;
                    bool alive = false;
                    for (int v = 0; v < K; ++v) alive = alive || q[v * B] > thr;
                    if (alive) mylive = j0;
                }
                mylive = __reduce_min_sync(0xffffffffu, mylive);
                if (lane == 0) s_jmin[(i & 1) * NW + wid] = mylive;
            }
            int ms = ms_lo;
            for (int m = m_lo; m <= mi; ++m, ms = (ms + 1 == NS ? 0 : ms + 1)) {
                const bool diag = m == mi && tid == ti;
                if (m == mi && tid > ti) break;
                if (m == m_lo && tid < t_lo) continue;
                const int j = m * B + tid;
                Real* pa = st_a + (size_t)ms * K * B + tid;
                Real* ph = st_h + (size_t)ms * K * B + tid;
                const Real tau = (Real)(wn - Wd[j]);
                const Real L2 = ex2.log2(tau);
                for (int v = 0; v < K; ++v) {
                    const Real ap = diag ? (Real)0 : pa[v * B] * invZ_prev;
                    const Real hprev = diag ? (Real)0 : ph[v * B];
                    const Real base = diag ? J_s[v] : ap * thin;
                    if (exporting && !diag) row_base[(int64_t)v * i + j] = ap;  // row i-1, scaled one step late
                    Real esum = 0, dsum = 0;
                    for (int s = 0; s < K; ++s) {
                        const int t = v * K + s;
                        const Real kE = ex2(pkm1[t] * L2 - pc2p[t]);  // A_vs(tau)
                        kE_s[s * BP + tid] = kE;
                        esum += kE * prk[t];
                        dsum += kE;
                    }
                    const Real hsum = tau * esum;  // H_v(tau)
                    const Real g = of_s[v] * ex2.neg(-om_l2e * (hsum - hprev)) * base;
                    if (!exporting) row_base[(int64_t)K * i + (int64_t)v * (i + 1) + j] = g;  // row i
                    Real gj = g;
                    // a final row (right end = t_end, smjp_utils.py:544-547) holds alpha = g WITHOUT the Omega A_v factor, so the
                    // 1/(Omega A_v) of the jump term (smjp_utils.py:409-412) does not cancel there (matters when a grid point lies
                    // within isclose() of t_end: two or more final rows)
                    if (!nonfinal) gj = dsum > (Real)0 ? g * M::rcp(omega_r * dsum) : (Real)0;
                    for (int s = 0; s < K; ++s) Jp_s[s * BP + tid] += gj * kE_s[s * BP + tid];
                    const Real an = g * (nonfinal ? omega_r * dsum : (Real)1);
                    pa[v * B] = an;
                    ph[v * B] = hsum;
                    Zp += an;
                }
            }
            Zp_s[tid] = Zp;
            __syncthreads();
            for (int r = tid; r <= K; r += B) {  // column sums: rows 0..K-1 = inflow into s, row K = Z
                const Real* src = r < K ? Jp_s + r * BP : Zp_s;
                Real acc = 0;
                for (int t = 0; t < B; ++t) acc += src[t];
                red_s[r] = acc;
            }
            __syncthreads();
            const Real Z = red_s[K];
            if (!(Z > (Real)0) || !(Z < (Real)INFINITY)) {
                degenerate = true;
                break;
            }
            invZ_prev = (Real)1 / Z;
            for (int v = tid; v < K; v += B) {
                J_s[v] = red_s[v] * invZ_prev;
                of_s[v] = (i + 1 < n) ? obsf[(i + 1) * K + v] : (Real)1;
            }
            if (a.logZ && tid == 0) a.logZ[w0 + i] = log((double)Z);
            if (i > 0) row_base += (int64_t)K * i;
            if (judge) {
                int jl = 0x7fffffff;
                for (int w = 0; w < NW; ++w) jl = min(jl, s_jmin[(i & 1) * NW + w]);
                const int j_new = max(j_lo, min(jl, min((m_lo + 1) * B, i)));  // the states of step i itself are not judged yet
                if (j_new / B != m_lo) ms_lo = ms_lo + 1 == NS ? 0 : ms_lo + 1;
                j_lo = j_new;
            }
            __syncthreads();
        }
        if (a.live_pairs && tid == 0) atomicAdd(a.live_pairs, live);
        if (degenerate || overflow) {  // block-uniform
            if (tid == 0) {
                a.status[c] = overflow ? SMJP_ST_OVERFLOW : SMJP_ST_DEGENERATE;
                a.path_len[c] = 0;
            }
            continue;
        }
        if (exporting) {  // last row: scale and store
            Real* row_last = msg + (int64_t)K * (n - 1) * n / 2;
            for (int j = tid; j < n; j += B) {
                const Real* p = st_a + (size_t)(j / B) * K * B + tid;
                for (int v = 0; v < K; ++v) row_last[(int64_t)v * n + j] = p[v * B] * invZ_prev;
            }
        }
        __syncthreads();

        // =========================== backward ===========================
        Real* wbuf = st_a;  // [K][len]: the reference's state-major order a = v*len + j
        int cur = n - 1, vplus = -1, cnt = 0, u_hi = -1;
        while (true) {
            const int len = cur + 1;
            const Real* row = msg + (int64_t)K * cur * (cur + 1) / 2;
            const int lo = exporting ? 0 : jlo_arr[cur];  // live window of this row: entries below lo were never written
            const int wl = len - lo;
            if (!a.uniforms && (u_hi < 0 || cur <= u_hi - B)) {
                if (cur - tid >= 0)
                    ubuf[tid] = philox_u01(a.seed, TAG_BACKWARD, (uint64_t)(cur - tid), 0u, a.sweep,
                                           a.chain_base + (uint32_t)c);
                u_hi = cur;
            }
            if (vplus < 0) {  // a_{n-1} ~ alpha_{n-1}  (fast_hmm.py:124-125); g = alpha on the last row
                for (int t = tid; t < K * wl; t += B) {
                    const int v = t / wl, jj = t - v * wl;
                    wbuf[t] = row[(int64_t)v * len + lo + jj];
                }
            } else {
                const double wn = Wd[cur + 1];
                const bool grow = !exporting && cur < i_final;  // the row holds g: one hazard per entry
                for (int j = lo + tid; j < len; j += B) {
                    const Real L2 = ex2.log2((Real)(wn - Wd[j]));
                    for (int v = 0; v < K; ++v) {
                        const Real hto = ex2(pkm1[v * K + vplus] * L2 - pc2p[v * K + vplus]);
                        Real r = hto;
                        if (!grow) {  // w = alpha_cur(v,j) A_{v,vplus}/A_v  (fast_hmm.py:153-171)
                            Real dsum = 0;
                            for (int s = 0; s < K; ++s) dsum += ex2(pkm1[v * K + s] * L2 - pc2p[v * K + s]);
                            r = dsum > (Real)0 ? hto * M::rcp(dsum) : (Real)0;
                        }
                        wbuf[v * wl + (j - lo)] = row[(int64_t)v * len + j] * r;
                    }
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
            const int v = idx / wl, j = lo + idx - v * wl;
            if (a.aug_idx)
                for (int t = j + tid; t <= cur; t += B) a.aug_idx[w0 + t] = v * n + j;
            __syncthreads();  // everyone has read s_sel and the weights before they are overwritten
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
        // ---- chain metrics (mcmc_utils.py:40-56) ----
        if (a.time_in_state || a.jumps) {
            for (int s = tid; s < K; s += B) {
                double tsum = 0;
                int jc = 0;
                for (int p = 0; p < plen; ++p) {
                    const int vs = p < cnt ? soj[cnt - 1 - p] : v_last;
                    if (vs == s) {
                        ++jc;
                        if (p + 1 < plen) {
                            const double t0 = Wd[soj[(ncap + 1) + cnt - 1 - p]];
                            const double t1 = (p + 1 < cnt) ? Wd[soj[(ncap + 1) + cnt - 2 - p]] : a.t_end;
                            tsum += t1 - t0;
                        }
                    }
                }
                if (a.time_in_state) a.time_in_state[(int64_t)c * K + s] = tsum;
                if (a.jumps) a.jumps[(int64_t)c * K + s] = jc;
            }
        }
    }
}

}  // namespace smjp
