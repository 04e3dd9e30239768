// FFBS, one WARP per chain, live window in a shared-memory ring (K <= 3, scratch mode; the code is generic in K).
//
// Same arithmetic, outputs and reference lines as ffbs.cuh.  What the exact live window of ffbs.cuh makes
// possible: on cfg2 only ~250 (fp64) / ~100 (fp32) grid points per row carry non-zero states, so the chain
// state fits a ring of R slots of 32 grid points, a chain needs one warp, a grid step ends in warp shuffles
// instead of a block-wide reduction and barrier, and 16-32 chains are resident per SM instead of 5-8.
//   lane L owns the grid points j with j % 32 == L; slot m = j / 32 lives in ring slot m % R;
//   W, the observation factors and the message rows are read from / written to global memory (L1/L2);
//   the backward uniforms are generated 32 at a time into registers and shuffled out;
//   the sojourns are written straight into the path arrays, last to first, and moved to the front at the end.
// A window that outgrows the ring (never observed on the BASELINE configurations, but nothing bounds it) sets
// SMJP_ST_OVERFLOW | SMJP_ST_DEGENERATE for the chain, and the launcher's rescue pass redoes such chains with the
// full-size kernel of ffbs.cuh -- results never depend on the ring size.
#pragma once
#include "ffbs.cuh"

namespace smjp {

#define PC(w, t) (sizeof(Real) == 8 ? (Real)a.pc_d[w][t] : (Real)a.pc_f[w][t])  // as in ffbs.cuh (K <= 3)
// hazard parameter set w, entry t: kernel-argument constants for K <= 3, shared memory beyond
#define PW(w, t) (K <= 3 ? PC(w, (K <= 3 ? (t) : 0)) : s_par[(w) * K * K + (t)])

constexpr int FFBS_WARPS = 4;  // chains in flight per CTA (independent warps)

template <typename Real>
__host__ __device__ inline size_t ffbs_warp_smem_bytes(int K, int ring, int ncap) {
    size_t per_warp = 2 * align_up((size_t)ring * K * 32 * sizeof(Real), 16)       // message, H_v
                      + align_up((size_t)(ncap + 1) * sizeof(unsigned short), 16);  // window start of every row
    return per_warp * FFBS_WARPS;
}

template <typename Real, int K>
// resident CTAs the register budget is tuned for: 7 in fp32 (72 registers), 4 in fp64 (<= 128); an explicit 1 lets
// the compiler take 107 / 161 registers and costs 17 % (measured)
__global__ void __launch_bounds__(FFBS_WARPS * 32, (sizeof(Real) == 4 && K <= 3 ? 7 : 4)) ffbs_warp_kernel(const FfbsArgs a) {
    using M = Math<Real>;
    using Acc = typename ScanAcc<Real>::type;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.guard && (a.guard[0] > a.guard_nmax || a.guard[1] > a.guard_total)) return;  // mis-speculated geometry
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ncap = a.ncap, Kobs = a.Kobs, R = a.nslot;  // nslot = ring slots here
    const size_t ring_b = align_up((size_t)R * K * 32 * sizeof(Real), 16);
    const size_t per_warp = 2 * ring_b + align_up((size_t)(ncap + 1) * sizeof(unsigned short), 16);
    unsigned char* mine = smem_raw + (size_t)wid * per_warp;
    Real* st_a = (Real*)mine;                       // [R][K][32] unnormalised message; backward: weights
    Real* st_h = (Real*)(mine + ring_b);            // [R][K][32] H_v(tau)
    unsigned short* jlo_arr = (unsigned short*)(mine + 2 * ring_b);

    __shared__ double s_tab[Exp2<Real>::TAB ? Exp2<Real>::TAB : 1];
    __shared__ double2 s_ltab[Exp2<Real>::LTAB ? Exp2<Real>::LTAB : 1];
    __shared__ Real s_par[3 * K * K];  // (k-1) SCALE, c2p SCALE for the backward's run-time target index
    for (int t = threadIdx.x; t < Exp2<Real>::TAB; t += blockDim.x) s_tab[t] = exp2_table_entry(t);
    for (int t = threadIdx.x; t < Exp2<Real>::LTAB; t += blockDim.x) {
        const double rc = 1.0 / (1.0 + ((double)t + 0.5) / 128.0);
        s_ltab[t] = make_double2(rc, -::log2(rc));
    }
    for (int t = threadIdx.x; t < K * K; t += blockDim.x) {
        const double k = a.model[t];
        s_par[t] = (Real)(1.0 / k);
        s_par[K * K + t] = (Real)((k - 1.0) * (double)Exp2<Real>::SCALE);
        s_par[2 * K * K + t] = (Real)((a.model[K * K + t] - ::log2(k)) * (double)Exp2<Real>::SCALE);
        if (K <= 3) {  // the same values the constant-bank copies hold (host-rounded)
            s_par[t] = PC(0, (K <= 3 ? t : 0));
            s_par[K * K + t] = PC(1, (K <= 3 ? t : 0));
            s_par[2 * K * K + t] = PC(2, (K <= 3 ? t : 0));
        }
    }
    __syncthreads();
    const Real* skm1 = s_par + K * K;
    const Real* sc2p = s_par + 2 * K * K;
    Exp2<Real> ex2;
    ex2.tab = s_tab;
    ex2.ltab = s_ltab;
    const double* emis = a.model + 2 * K * K;
    const Real thin = (Real)(1.0 - 1.0 / a.omega);
    const Real omega_r = (Real)a.omega;
    const Real om_l2e = (Real)(a.omega * SMJP_LOG2E * (double)Exp2<Real>::SCALE);
    const Real power = (Real)a.power;
    const Real weps = (Real)a.window_eps;
    const int gw = blockIdx.x * FFBS_WARPS + wid;  // global warp index: owns one scratch region
    Real* obsf = (Real*)a.obsf_scratch + (int64_t)gw * a.obsf_stride;  // later: the sojourn records (t, v)
    Real* msg = (Real*)a.msg_scratch + (int64_t)gw * a.msg_scratch_stride;

    while (true) {
        int c = -1;
        if (lane == 0) {
            const unsigned int q = atomicAdd(a.queue, 1u);
            const unsigned limit = a.C_dev ? (unsigned)*a.C_dev : (unsigned)a.C;
            c = (q < limit) ? (a.order ? a.order[q] : (int)q) : -1;
        }
        c = __shfl_sync(FULL, c, 0);
        if (c < 0) break;
        const int64_t w0 = a.w_offs[c];
        const int n = (int)(a.w_offs[c + 1] - w0) - 1;
        if (n < 1 || n > ncap) {
            if (lane == 0) {
                a.status[c] = (n < 1) ? SMJP_ST_GRID_NOT_INCREASING : SMJP_ST_OVERFLOW;
                a.path_len[c] = 0;
            }
            continue;
        }
        const double* Wg = a.W + w0;
        // ---- validate the grid (smjp_utils.py:398-399), observation factor per interval ----
        int bad = 0;
        {
            const int64_t o0 = a.obs_offs[c];
            const int D = (int)(a.obs_offs[c + 1] - o0);
            const double* ot = a.obs_t + o0;
            const int32_t* oy = a.obs_y + o0;
            for (int i = lane; i < n; i += 32) {
                const double wc = Wg[i], wn = Wg[i + 1];
                bad |= !(wn > wc);
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
                        bad |= 2;
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
        if (__any_sync(FULL, bad)) {
            const int grid_bad = __any_sync(FULL, bad & 1), sym_bad = __any_sync(FULL, bad & 2);
            if (lane == 0) {
                a.status[c] = (grid_bad ? SMJP_ST_GRID_NOT_INCREASING : 0) | (sym_bad ? SMJP_ST_BAD_INPUT : 0);
                a.path_len[c] = 0;
            }
            continue;
        }
        __syncwarp();

        // =========================== forward ===========================
        Real invZ_prev = (Real)1;
        Real J[K], of[K];
#pragma unroll
        for (int v = 0; v < K; ++v) {
            J[v] = (Real)(1.0 / K);  // uniform prior on (v,0): smjp_utils.py:286-289
            of[v] = obsf[v];
        }
        int flag = 0;
        int i_final = n;
        while (i_final > 0 && np_isclose(a.t_end, Wg[i_final])) --i_final;
        Real* row_i = msg;  // start of row i (row r holds K*(r+1) values, [v][j])
        int j_lo = 0;
        unsigned long long live = 0;
        for (int i = 0; i < n; ++i) {
            const double wn = Wg[i + 1];
            const bool nonfinal = i < i_final;
            Real ofn[K];
#pragma unroll
            for (int v = 0; v < K; ++v) ofn[v] = (i + 1 < n) ? obsf[(i + 1) * K + v] : (Real)1;
            const int mi = i >> 5;
            const int m_lo = j_lo >> 5;
            if ((i & 3) == 0) {  // window start: smallest grid point of the oldest slot still above zero / eps * Z
                int mylive = 0x7fffffff;
                const int j0 = (m_lo << 5) + lane;
                if (j0 >= j_lo && j0 < i) {
                    const Real thr = weps > (Real)0 ? weps / invZ_prev : (Real)0;
                    const Real* q = st_a + (size_t)(m_lo % R) * K * 32 + lane;
                    bool alive = false;
#pragma unroll
                    for (int v = 0; v < K; ++v) alive = alive || q[v * 32] > thr;
                    if (alive) mylive = j0;
                }
                mylive = __reduce_min_sync(FULL, mylive);
                j_lo = max(j_lo, min(mylive, min((m_lo + 1) << 5, i)));
            }
            if (mi - (j_lo >> 5) >= R) {  // the window has outgrown the ring: leave the chain to the full-size kernel
                flag = SMJP_ST_OVERFLOW | SMJP_ST_DEGENERATE;
                break;
            }
            if (lane == 0) jlo_arr[i] = (unsigned short)j_lo;
            live += (unsigned long long)(i - j_lo + 1);
            Real Zp = 0;
            Real Jp[K];
#pragma unroll
            for (int v = 0; v < K; ++v) Jp[v] = 0;
            int rs = (j_lo >> 5) % R;  // ring slot of the oldest live slot, advanced without a division per slot
            for (int m = j_lo >> 5; m <= mi; ++m, rs = rs + 1 == R ? 0 : rs + 1) {
                const int j = (m << 5) + lane;
                if (j < j_lo || j > i) continue;
                const bool diag = j == i;
                Real* pa = st_a + rs * (K * 32) + lane;
                Real* ph = st_h + rs * (K * 32) + lane;
                const Real tau = (Real)(wn - Wg[j]);
                const Real L2 = ex2.log2(tau);
#pragma unroll
                for (int v = 0; v < K; ++v) {
                    const Real ap = diag ? (Real)0 : pa[v * 32] * invZ_prev;
                    const Real hprev = diag ? (Real)0 : ph[v * 32];
                    const Real base = diag ? J[v] : ap * thin;
                    Real kE[K];
                    Real esum = 0, dsum = 0;
#pragma unroll
                    for (int s = 0; s < K; ++s) {
                        kE[s] = ex2(PW(1, v * K + s) * L2 - PW(2, v * K + s));  // A_vs(tau)
                        esum += kE[s] * PW(0, v * K + s);
                        dsum += kE[s];
                    }
                    const Real hsum = tau * esum;  // H_v(tau)
                    const Real g = of[v] * ex2.neg(-om_l2e * (hsum - hprev)) * base;
                    row_i[v * (i + 1) + j] = g;  // scratch rows hold g (ffbs.cuh, Backward)
                    Real gj = g;
                    // a final row (right end = t_end, smjp_utils.py:544-547) holds alpha = g WITHOUT the Omega A_v factor, so the
                    // 1/(Omega A_v) of the jump term (smjp_utils.py:409-412) does not cancel there (matters when a grid point lies
                    // within isclose() of t_end: two or more final rows)
                    if (!nonfinal) gj = dsum > (Real)0 ? g * M::rcp(omega_r * dsum) : (Real)0;
#pragma unroll
                    for (int s = 0; s < K; ++s) Jp[s] += gj * kE[s];
                    const Real an = g * (nonfinal ? omega_r * dsum : (Real)1);
                    pa[v * 32] = an;
                    ph[v * 32] = hsum;
                    Zp += an;
                }
            }
            // warp reduction of (Z, J[0..K)): lane (idx << SH) ends up with the total of value idx
            Real Z;
            {
                constexpr int NVP = pow2_ceil(K + 1), SH = 5 - log2_of(NVP);
                Real vals[NVP];
                vals[0] = Zp;
#pragma unroll
                for (int v = 0; v < K; ++v) vals[1 + v] = Jp[v];
#pragma unroll
                for (int v = K + 1; v < NVP; ++v) vals[v] = 0;
                WarpHalve<Real, NVP, 16>::run(vals, lane);
                Z = __shfl_sync(FULL, vals[0], 0);
#pragma unroll
                for (int v = 0; v < K; ++v) J[v] = __shfl_sync(FULL, vals[0], (1 + v) << SH);
            }
            if (!(Z > (Real)0) || !(Z < (Real)INFINITY)) {
                flag = SMJP_ST_DEGENERATE;
                break;
            }
            invZ_prev = (Real)1 / Z;
#pragma unroll
            for (int v = 0; v < K; ++v) {
                J[v] *= invZ_prev;
                of[v] = ofn[v];
            }
            if (a.logZ && lane == 0) a.logZ[w0 + i] = log((double)Z);
            row_i += (int64_t)K * (i + 1);
            __syncwarp();
        }
        if (a.live_pairs && lane == 0) atomicAdd(a.live_pairs, live);
        if (flag) {
            if (lane == 0) {
                a.status[c] = flag;
                a.path_len[c] = 0;
            }
            continue;
        }
        __syncwarp();

        // =========================== backward ===========================
        Real* wbuf = st_a;  // [K][wl] weights of the live window, state-major
        // sojourn records, found last to first, go to this warp's (now dead) observation-factor scratch: the
        // output arrays are only written (they may be pinned host memory)
        double* pt = (double*)obsf;      // [n + 1]
        int32_t* pv = (int32_t*)(pt + n + 1);  // [n + 1]
        int cur = n - 1, vplus = -1, cnt = 0, u_hi = -1;
        double u_l = 0.0;  // lane L holds the uniform of step u_hi - L
        while (true) {
            const int len = cur + 1;
            const Real* row = msg + (int64_t)K * cur * (cur + 1) / 2;
            const int lo = jlo_arr[cur], wl = len - lo;
            {   // next rows' live parts into L2 (ffbs.cuh): one 128-byte line per lane and (row, state) segment
#pragma unroll
                for (int seg = 0; seg < 3 * K; ++seg) {
                    const int r = cur - 1 - seg / K, v = seg % K;
                    if (r >= 0) {
                        const int rlo = jlo_arr[r];
                        const char* p = (const char*)(msg + (int64_t)K * r * (r + 1) / 2 + (int64_t)v * (r + 1) + rlo) + lane * 128;
                        if (p < (const char*)(msg + (int64_t)K * r * (r + 1) / 2 + (int64_t)(v + 1) * (r + 1)))
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                    }
                }
            }
            if (!a.uniforms && (u_hi < 0 || cur <= u_hi - 32)) {
                u_l = cur - lane >= 0 ? philox_u01(a.seed, TAG_BACKWARD, (uint64_t)(cur - lane), 0u, a.sweep,
                                                   a.chain_base + (uint32_t)c)
                                      : 0.0;
                u_hi = cur;
            }
            if (vplus < 0) {  // a_{n-1} ~ alpha_{n-1}  (fast_hmm.py:124-125); g = alpha on the last row
                for (int j = lo + lane; j < len; j += 32) {
#pragma unroll
                    for (int v = 0; v < K; ++v) wbuf[v * wl + (j - lo)] = row[(int64_t)v * len + j];
                }
            } else if (cur < i_final) {  // the row holds g: w = g_cur(v,j) A_{v,vplus}(tau)
                const double wn = Wg[cur + 1];
                for (int j = lo + lane; j < len; j += 32) {
                    Real rv[K];
#pragma unroll
                    for (int v = 0; v < K; ++v) rv[v] = row[(int64_t)v * len + j];
                    const Real L2 = ex2.log2((Real)(wn - Wg[j]));
#pragma unroll
                    for (int v = 0; v < K; ++v)
                        wbuf[v * wl + (j - lo)] = rv[v] * ex2(skm1[v * K + vplus] * L2 - sc2p[v * K + vplus]);
                }
            } else {  // final rows: w = alpha_cur(v,j) A_{v,vplus}/A_v  (fast_hmm.py:153-171)
                const double wn = Wg[cur + 1];
                for (int j = lo + lane; j < len; j += 32) {
                    const Real L2 = ex2.log2((Real)(wn - Wg[j]));
#pragma unroll
                    for (int v = 0; v < K; ++v) {
                        Real dsum = 0, hto = 0;
#pragma unroll
                        for (int s = 0; s < K; ++s) {
                            const Real kE = ex2(skm1[v * K + s] * L2 - sc2p[v * K + s]);
                            dsum += kE;
                            hto = (s == vplus) ? kE : hto;
                        }
                        const Real r = dsum > (Real)0 ? hto * M::rcp(dsum) : (Real)0;
                        wbuf[v * wl + (j - lo)] = row[(int64_t)v * len + j] * r;
                    }
                }
            }
            __syncwarp();
            // inclusive scan over the flat weight vector, contiguous chunk per lane
            const int L = K * wl;
            const int q = (L + 31) / 32;
            const int beg = min(lane * q, L), end = min(beg + q, L);
            Acc loc = 0;
            for (int t = beg; t < end; ++t) loc += (Acc)wbuf[t];
            const Acc incl = warp_inclusive_scan(loc, lane);
            Acc excl = __shfl_up_sync(FULL, incl, 1);
            if (lane == 0) excl = 0;
            // the total is the inclusive value of the LAST lane that holds entries (lanes past it carry zeros whose
            // scan values need not agree to the last bit)
            const int last_lane = (L - 1) / q;
            const Acc total = __shfl_sync(FULL, incl, last_lane);
            if (!(total > (Acc)0) || !(total < (Acc)INFINITY)) {
                flag = SMJP_ST_DEGENERATE;
                break;
            }
            const double u = a.uniforms ? a.uniforms[w0 + cur] : __shfl_sync(FULL, u_l, u_hi - cur);
            const Acc target = (Acc)u * total;
            const bool has = end > beg && incl > excl;
            int sel = -1;
            if (has && target >= excl && target < incl) {
                Acc acc = excl;
                int lastpos = beg;
                for (int t = beg; t < end; ++t) {
                    const Acc wv = (Acc)wbuf[t];
                    if (wv > (Acc)0) lastpos = t;
                    acc += wv;
                    if (acc > target) { sel = t; break; }
                }
                if (sel < 0) sel = lastpos;
            }
            unsigned own = __ballot_sync(FULL, sel >= 0);
            if (!own) {  // u*total rounded up to the total: the last lane with mass takes its last positive entry
                const unsigned hm = __ballot_sync(FULL, has);
                const int ll = 31 - __clz(hm);
                if (lane == ll) {
                    int lastpos = beg;
                    for (int t = beg; t < end; ++t) if ((Acc)wbuf[t] > (Acc)0) lastpos = t;
                    sel = lastpos;
                }
                own = 1u << ll;
            }
            const int idx = __shfl_sync(FULL, sel, __ffs(own) - 1);
            int v = 0;
#pragma unroll
            for (int u2 = 1; u2 < K; ++u2) v += idx >= u2 * wl;
            const int j = lo + idx - v * wl;
            if (a.aug_idx)
                for (int t = j + lane; t <= cur; t += 32) a.aug_idx[w0 + t] = v * n + j;
            if (lane == 0) {  // sojourns are found last to first: fill the path arrays from the end
                pv[n - cnt] = v;
                pt[n - cnt] = Wg[j];
            }
            ++cnt;
            __syncwarp();
            if (j == 0) break;
            vplus = v;
            cur = j - 1;
        }
        if (flag) {
            if (lane == 0) {
                a.status[c] = flag;
                a.path_len[c] = 0;
            }
            continue;
        }
        __syncwarp();
        // ---- path compaction (smjp_utils.py:319-330): records sit at [n-cnt+1 .. n] in time order ----
        const int src0 = n - cnt + 1;
        for (int p = lane; p < cnt; p += 32) {
            a.path_v[w0 + p] = pv[src0 + p];
            a.path_t[w0 + p] = pt[src0 + p];
        }
        const int v_last = pv[n];
        const double t_last = pt[n];
        const bool add_end = !np_isclose(t_last, a.t_end);  // smjp_utils.py:338-348
        const int plen = cnt + (add_end ? 1 : 0);
        if (lane == 0) {
            if (add_end) {
                a.path_v[w0 + cnt] = v_last;
                a.path_t[w0 + cnt] = a.t_end;
            }
            a.path_len[c] = plen;
            a.status[c] = 0;
        }
        // ---- chain metrics (mcmc_utils.py:40-56): lane s < K, fixed summation order ----
        if (lane < K && (a.time_in_state || a.jumps)) {
            double tsum = 0;
            int jc = 0;
            for (int p = 0; p < plen; ++p) {
                const int vs = p < cnt ? pv[src0 + p] : v_last;
                if (vs == lane) {
                    ++jc;
                    if (p + 1 < plen) {
                        const double t0 = pt[src0 + p];
                        const double t1 = (p + 1 < cnt) ? pt[src0 + p + 1] : a.t_end;
                        tsum += t1 - t0;
                    }
                }
            }
            if (a.time_in_state) a.time_in_state[(int64_t)c * K + lane] = tsum;
            if (a.jumps) a.jumps[(int64_t)c * K + lane] = jc;
        }
        __syncwarp();
    }
}

#undef PW
#undef PC

}  // namespace smjp
