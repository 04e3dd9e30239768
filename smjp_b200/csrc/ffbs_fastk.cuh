// FFBS, ONE WARP per chain, 4 <= K <= 10, SCRATCH MODE ONLY: the K >= 4 sibling of ffbs_fast.cuh.
//
// Same reference lines, arithmetic, scratch-row layout ([j][v], as ffbs_item.cuh) and outputs as ffbs_item.cuh (the
// export / rescue kernel for these K); organised like ffbs_fast.cuh:
//   * lane = (jt, v): JPS = 32 / K grid points x K source states per warp slot (K = 10: 30 of 32 lanes); a lane owns
//     ONE source state v for all its grid points, evaluates the K hazards A_vs of that row (its 3K parameters live in
//     registers) and keeps K inflow accumulators;
//   * normalisation deferred by RS = 8 steps, folded into the next observation factor; the per-step totals
//     (U, J[0..K)) come out of one folded warp reduction (WarpHalve) and reach the lanes that seed the new state by
//     shuffle: no shared-memory reduction, no barrier, no division on the other steps;
//   * state (message, H_v) and the window's grid times in a shared-memory RING of NS slots of JPS grid points; the
//     live-window start is re-examined every RS steps over as many slots as have died; a window that outgrows the
//     ring, or a chain whose hazard exponents leave the range of the unclamped fp64 2^x, goes to the rescue pass;
//   * WAYS slots in flight per lane; the grid and the row starts stay in L2.
// Scale bookkeeping and the log Z_i conversion: header of ffbs_fast.cuh.
#pragma once
#include "ffbs_fast.cuh"

namespace smjp {

// dynamic shared memory of ffbs_fastk_kernel (per chain = per CTA)
template <typename Real>
__host__ __device__ inline FfbsLayout ffbs_fastk_layout(int K, int ring, int ncap) {
    const int jps = 32 / K;
    const size_t state = (size_t)ring * jps * K * sizeof(Real);
    FfbsLayout L;
    size_t b = 0;
    L.wd = (int)b;   b += align_up((size_t)ring * jps * sizeof(double), 16);          // grid times of the window (ring)
    L.ubuf = (int)b; b += align_up((size_t)32 * sizeof(double), 16);                  // backward uniforms
    L.scan = (int)b;                                                                  // (unused)
    // Shared memory decides how many chains an SM holds (fp64, K = 10: 7 with a separate sojourn list and the hazard
    // tables here, 8 without): the backward pass's sojourn list reuses the H ring, dead by then, when it fits, and the
    // K x K hazard tables the backward pass reads live in the CTA's global scratch region (L1-resident, 2.4 KB).
    const size_t soj = (size_t)2 * (ncap + 1) * sizeof(int);
    L.sta = (int)b;  b += align_up(state, 16);                                        // message ring; backward: weights
    L.sth = (int)b;  b += align_up(state, 16);                                        // H ring (contiguous with sta)
    L.red = L.sth;                                                                    // backward: sojourn list
    if (soj > state) { L.red = (int)b; b += align_up(soj, 16); }
    L.par = (int)b;                                                                   // hazard tables: fp32 only
    if (sizeof(Real) == 4) b += align_up((size_t)(3 * K * K) * sizeof(Real), 16);     // (fp32 is register-bound: 16 chains either way)
    L.jlo = (int)b;
    L.total = (int)b;
    return L;
}

template <typename Real, int K, int WAYS, int MINB>
__global__ void __launch_bounds__(32, MINB) ffbs_fastk_kernel(const FfbsArgs a) {
    using M = Math<Real>;
    using Acc = typename ScanAcc<Real>::type;
    using FL = FastLog2<Real>;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int RS = FAST_RS;
    constexpr int JPS = 32 / K;   // grid points per slot
    constexpr int IT = JPS * K;   // working lanes
    static_assert(K >= 4 && K <= 16, "K");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.guard && (a.guard[0] > a.guard_nmax || a.guard[1] > a.guard_total)) return;  // mis-speculated geometry
    const int lane = threadIdx.x;
    const bool work = lane < IT;
    const int jt = work ? lane / K : JPS - 1, v = work ? lane - (lane / K) * K : K - 1;  // (idle lanes mirror the last item)
    const int ncap = a.ncap, Kobs = a.Kobs, NS = a.nslot;

    double* Wr = (double*)(smem_raw + a.lay.wd);      // [slot][JPS]
    double* ubuf = (double*)(smem_raw + a.lay.ubuf);
    Real* st_a = (Real*)(smem_raw + a.lay.sta);       // [slot][IT]
    Real* st_h = (Real*)(smem_raw + a.lay.sth);
    int* soj = (int*)(smem_raw + a.lay.red);
    const int region = blockIdx.x;
    Real* obsf = (Real*)a.obsf_scratch + (int64_t)region * a.obsf_stride;
    int* jlo_arr = (int*)(obsf + (int64_t)ncap * K);
    // hazard tables [3][K*K]: fp64 behind the window starts in global scratch, fp32 in shared memory; written once by
    // this CTA (device arithmetic, like every kernel)
    Real* srk = sizeof(Real) == 8 ? obsf + (int64_t)ncap * K + ffbs_fast_jlo_elems<Real>(ncap) : (Real*)(smem_raw + a.lay.par);
    Real* skm1 = srk + K * K;
    Real* sc2p = skm1 + K * K;

    __shared__ double s_tab[Exp2<Real>::TAB ? Exp2<Real>::TAB : 1];
    __shared__ double2 s_ltab[Exp2<Real>::LTAB ? Exp2<Real>::LTAB : 1];
    __shared__ __align__(8) unsigned long long s_obar;  // completion of the observation bulk copy (ffbs_fast.cuh)
    uint32_t obs_parity = 0;
    if (lane == 0) mbar_init(&s_obar, 1);
    for (int t = lane; t < K * K; t += 32) {
        const double k = a.model[t];
        srk[t] = (Real)(1.0 / k);
        skm1[t] = (Real)((k - 1.0) * (double)Exp2<Real>::SCALE);
        sc2p[t] = (Real)((a.model[K * K + t] - ::log2(k)) * (double)Exp2<Real>::SCALE);
    }
    for (int t = lane; t < Exp2<Real>::TAB; t += 32) s_tab[t] = exp2_table_entry(t);
    for (int t = lane; t < Exp2<Real>::LTAB; t += 32) {
        const double rc = 1.0 / (1.0 + ((double)t + 0.5) / 128.0);
        s_ltab[t] = make_double2(rc, -::log2(rc));
    }
    __syncwarp();
    Exp2<Real> ex2;
    ex2.tab = s_tab;
    ex2.ltab = s_ltab;
    const double* emis = a.model + 2 * K * K;
    Real r_rk[K], r_km1[K], r_c2p[K];  // the K hazards of this lane's row v
#pragma unroll
    for (int s = 0; s < K; ++s) {
        r_rk[s] = srk[v * K + s];
        r_km1[s] = skm1[v * K + s];
        r_c2p[s] = sc2p[v * K + s];
    }
    // range of this model's hazard exponents, as a function of L = log2 tau: |km1 L - c| <= 900 for L in [l0, l1]
    const Real om1 = (Real)(a.omega - 1.0);
    const Real c_new = (Real)(1.0 / (a.omega - 1.0));
    const Real om_l2e = (Real)(a.omega * SMJP_LOG2E * (double)Exp2<Real>::SCALE);
    const Real power = (Real)a.power;
    const Real weps = (Real)a.window_eps;
    const int64_t obs_total = a.obs_offs[a.C];

    while (true) {
        int c;
        {
            unsigned int q = 0;
            if (lane == 0) q = atomicAdd(a.queue, 1u);
            q = __shfl_sync(FULL, q, 0);
            const unsigned limit = a.C_dev ? (unsigned)*a.C_dev : (unsigned)a.C;
            c = (q < limit) ? (a.order ? a.order[q] : (int)q) : -1;
        }
        if (c < 0) break;
        __syncwarp();
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
        // ---- observations: TMA bulk copy of the chain's slice into the (idle) state ring, under the grid validation ----
        const int64_t o0 = a.obs_offs[c];
        const int D = (int)(a.obs_offs[c + 1] - o0);
        const ObsStage os = obs_stage_plan(o0, D);
        const bool staged = D > 0 && (size_t)os.bt + os.by <= 2 * (size_t)NS * IT * sizeof(Real) &&
                            os.t0 + (int64_t)(os.bt / 8) <= obs_total && os.y0 + (int64_t)(os.by / 4) <= obs_total &&
                            ((reinterpret_cast<uintptr_t>(a.obs_t) | reinterpret_cast<uintptr_t>(a.obs_y)) & 15) == 0;
        double* so_t = (double*)st_a;  // st_a and st_h are contiguous
        int32_t* so_y = (int32_t*)((char*)st_a + os.bt);
        if (staged && lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s_obar, os.bt + os.by);
            bulk_g2s(so_t, a.obs_t + os.t0, os.bt, &s_obar);
            bulk_g2s(so_y, a.obs_y + os.y0, os.by, &s_obar);
        }
        // ---- validate the grid (smjp_utils.py:398-399), smallest gap ----
        int bad = 0;
        double dmin = INFINITY;
        for (int t = lane; t < n; t += 32) {
            const double gap = Wg[t + 1] - Wg[t];
            bad |= !(gap > 0.0);
            dmin = fmin(dmin, gap);
        }
        bad = __any_sync(FULL, bad);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(FULL, dmin, o));
        // ---- observation factor per interval (as in ffbs.cuh) ----
        int badsym = 0;
        {
            if (staged) {
                mbar_wait(&s_obar, obs_parity);
                obs_parity ^= 1u;
            }
            const double* ot = staged ? so_t + os.dt : a.obs_t + o0;
            const int32_t* oy = staged ? so_y + os.dy : a.obs_y + o0;
            for (int i = lane; i < n; i += 32) {
                const double wc = Wg[i], wn = Wg[i + 1];
                int lo = 0, hi = D;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (ot[mid] < wc) lo = mid + 1; else hi = mid;
                }
                Real f[K];
#pragma unroll
                for (int s = 0; s < K; ++s) f[s] = a.obs_product ? (Real)1 : (Real)0;
                int cnt = 0;
                for (int d = lo; d < D && ot[d] <= wn; ++d, ++cnt) {
                    int y = oy[d];
                    if ((unsigned)y >= (unsigned)Kobs) {  // never index the emission table out of bounds
                        y = 0;
                        badsym = 1;
                    }
#pragma unroll
                    for (int s = 0; s < K; ++s) {
                        const Real pe = (Real)emis[s * Kobs + y];
                        f[s] = a.obs_product ? f[s] * pe : f[s] + pe;
                    }
                }
#pragma unroll
                for (int s = 0; s < K; ++s)
                    obsf[i * K + s] = (cnt == 0 || power == (Real)0) ? (Real)1
                                      : (power == (Real)1 ? f[s] : M::pow(f[s], power));
            }
        }
        __syncwarp();
        badsym = __any_sync(FULL, badsym);
        if (bad || badsym) {
            if (lane == 0) {
                a.status[c] = (bad ? SMJP_ST_GRID_NOT_INCREASING : 0) | (badsym ? SMJP_ST_BAD_INPUT : 0);
                a.path_len[c] = 0;
            }
            continue;
        }
        if (sizeof(Real) == 8) {
            const double span = Wg[n] - Wg[0];
            bool ok = dmin >= 1.0e-270 && span <= 1.0e30;
            if (ok) {
                const double l0 = ::log2(dmin), l1 = ::log2(span);
                const double inv = 1.0 / (double)Exp2<Real>::SCALE;
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    const double km1 = (double)r_km1[s] * inv, cc = (double)r_c2p[s] * inv;
                    ok = ok && fabs(km1 * l0 - cc) <= 900.0 && fabs(km1 * l1 - cc) <= 900.0;
                }
            }
            if (!__all_sync(FULL, ok)) {
                if (lane == 0) {
                    a.status[c] = SMJP_ST_OVERFLOW | SMJP_ST_DEGENERATE;
                    a.path_len[c] = 0;
                }
                continue;
            }
        }

        Real* msg = (Real*)a.msg_scratch + (int64_t)region * a.msg_scratch_stride;

        // =========================== forward ===========================
        int i_final = n;
        while (i_final > 0 && np_isclose(a.t_end, Wg[i_final])) --i_final;
        bool degenerate = false, overflow = false;
        int j_lo = 0, ms_lo = 0;  // first live grid point; ring slot of its slot m_lo = j_lo / JPS
        unsigned long long live = 0;
        Real sc = om1;
        Real ofn = obsf[v];       // observation factor of the coming step for this lane's state
        const Real* obsf_next = obsf + K + v;
        Real* row_i = msg;        // row i: K * (i + 1) values, entry (j, v) at j * K + v
        double wn_next = Wg[1];
        if (work && jt == 0) {    // state (v, W[0]): uniform prior (smjp_utils.py:286-289), H = 0
            st_a[lane] = (Real)(1.0 / K) * c_new;
            st_h[lane] = (Real)0;
            if (v == 0) Wr[0] = Wg[0];
        }
        __syncwarp();

        auto step = [&](auto final_tag, const int i) -> bool {
            constexpr bool FINAL = decltype(final_tag)::value;
            const int mi = i / JPS;            // slot of the newest state (v, W[i])
            const int m_lo = j_lo / JPS;
            if (mi - m_lo >= NS) {
                overflow = true;
                return false;
            }
            const double wn = wn_next;
            wn_next = i + 2 <= n ? Wg[i + 2] : 0.0;
            if (lane == 0) jlo_arr[i] = j_lo;
            live += (unsigned long long)(i - j_lo + 1);
            const Real oft = ofn * sc;
            ofn = (i + 1 < n) ? *obsf_next : (Real)1;
            obsf_next += K;
            Real Zp = 0;
            Real Jp[K];
#pragma unroll
            for (int s = 0; s < K; ++s) Jp[s] = 0;
            int ms = ms_lo;
            Real* rowg = row_i + (int64_t)m_lo * IT + lane;  // entry (j, v) of slot m_lo: (m_lo * JPS + jt) * K + v
            auto items = [&](auto nway_tag, const int m) {
                constexpr int NWAY = decltype(nway_tag)::value;
                bool act[NWAY];
                int so[NWAY];
                Real tau[NWAY], L2[NWAY], kE[NWAY][K], esum[NWAY], dsum[NWAY], ap[NWAY], hp[NWAY];
#pragma unroll
                for (int p = 0; p < NWAY; ++p) {
                    const int j = (m + p) * JPS + jt;
                    act[p] = work && j >= j_lo && j <= i;
                    int sp = ms + p;
                    sp = sp >= NS ? sp - NS : sp;
                    so[p] = sp * IT + (work ? lane : 0);
                    tau[p] = act[p] ? (Real)(wn - Wr[sp * JPS + jt]) : (Real)1;
                    ap[p] = act[p] ? st_a[so[p]] : (Real)0;
                    hp[p] = act[p] ? st_h[so[p]] : (Real)0;
                    esum[p] = 0;
                    dsum[p] = 0;
                }
#pragma unroll
                for (int p = 0; p < NWAY; ++p) L2[p] = FL::run(ex2, tau[p]);
#pragma unroll
                for (int s = 0; s < K; ++s) {
#pragma unroll
                    for (int p = 0; p < NWAY; ++p) {
                        kE[p][s] = FL::hazard(ex2, r_km1[s] * L2[p] - r_c2p[s]);  // A_vs(tau)
                        esum[p] += kE[p][s] * r_rk[s];
                        dsum[p] += kE[p][s];
                    }
                }
#pragma unroll
                for (int p = 0; p < NWAY; ++p) {
                    const Real hsum = tau[p] * esum[p];  // H_v(tau)
                    const Real g = (oft * ex2.neg(-om_l2e * (hsum - hp[p]))) * ap[p];
                    Real gj = g, an = g * dsum[p];
                    if (FINAL) {
                        an = g;
                        gj = dsum[p] > (Real)0 ? g * M::rcp(dsum[p]) : (Real)0;
                    }
                    if (act[p]) {
                        rowg[p * IT] = g;  // scratch row i, [j][v]: the backward jump weight is g A_{v,v+}
                        st_a[so[p]] = an;
                        st_h[so[p]] = hsum;
                    }
#pragma unroll
                    for (int s = 0; s < K; ++s) Jp[s] += gj * kE[p][s];
                    Zp += an;
                }
                rowg += NWAY * IT;
                ms += NWAY;
                ms = ms >= NS ? ms - NS : ms;
            };
            {
                int m = m_lo;
                if (WAYS > 1) {
#pragma unroll 1
                    for (; m + WAYS - 1 <= mi; m += WAYS) items(std::integral_constant<int, WAYS>(), m);
                }
#pragma unroll 1
                for (; m <= mi; ++m) items(std::integral_constant<int, 1>(), m);
            }
            // totals of (U, J[0..K)) over the warp: lane (idx << SH) ends up with value idx
            constexpr int NVP = pow2_ceil(K + 1), SH = 5 - log2_of(NVP);
            Real vals[NVP];
            vals[0] = Zp;
#pragma unroll
            for (int s = 0; s < K; ++s) vals[1 + s] = Jp[s];
#pragma unroll
            for (int s = K + 1; s < NVP; ++s) vals[s] = 0;
            WarpHalve<Real, NVP, 16>::run(vals, lane);
            const Real U = __shfl_sync(FULL, vals[0], 0);
            const Real Jv = __shfl_sync(FULL, vals[0], (1 + v) << SH);  // inflow into this lane's state
            const bool sync_step = ((i + 1) % RS) == 0 || i == n - 1;
            if (sync_step) {
                if (!(U > (Real)0) || !(U < (Real)INFINITY)) {
                    degenerate = true;
                    return false;
                }
                sc = om1 / U;  // renormalisation, folded into the next step's observation factors
                // live-window judgement on the state after step i: advance over every slot whose states are all dead
                // (exact mode: zero; pruned mode: below window_eps of the last renormalised total)
                int mm = m_lo, msx = ms_lo;
                while (true) {
                    const int j0 = mm * JPS + jt;
                    const bool alive = work && j0 >= j_lo && j0 <= i && st_a[msx * IT + lane] > weps;
                    const int jl = __reduce_min_sync(FULL, alive ? j0 : 0x7fffffff);
                    const int cap = min((mm + 1) * JPS, i + 1);
                    if (jl < cap) {
                        j_lo = max(j_lo, jl);
                        break;
                    }
                    j_lo = max(j_lo, cap);  // the whole slot is dead
                    if (cap == i + 1) break;
                    ++mm;
                    msx = msx + 1 == NS ? 0 : msx + 1;
                }
                const int m_new = j_lo / JPS;
                ms_lo += m_new - m_lo;
                ms_lo = ms_lo >= NS ? ms_lo % NS : ms_lo;
            } else {
                sc = om1;
            }
            if (a.logZ && lane == 0) a.logZ[w0 + i] = (double)U;  // raw totals; turned into log Z_i after the pass
            if (i + 1 < n) {
                const int jn = i + 1, sl = (jn / JPS) % NS;
                if (work && jt == jn - (jn / JPS) * JPS) {
                    st_a[sl * IT + lane] = Jv * c_new;
                    st_h[sl * IT + lane] = (Real)0;
                    if (v == 0) Wr[sl * JPS + jt] = wn;  // W[i + 1]
                }
            }
            __syncwarp();
            row_i += (int64_t)K * (i + 1);
            return true;
        };
        {
            int i = 0;
            bool go = true;
            for (; go && i < i_final; ++i) go = step(std::false_type(), i);
            for (; go && i < n; ++i) go = step(std::true_type(), i);
        }
        __syncwarp();
        if (a.live_pairs && lane == 0) atomicAdd(a.live_pairs, live);
        if (degenerate || overflow) {
            if (lane == 0) {
                a.status[c] = overflow ? (SMJP_ST_OVERFLOW | SMJP_ST_DEGENERATE) : SMJP_ST_DEGENERATE;
                a.path_len[c] = 0;
            }
            continue;
        }
        if (a.logZ) {  // raw totals U_i -> log Z_i (ffbs_fast.cuh), 32 entries at a time from the end
            const double l_om = log(a.omega);
            for (int hi = n - 1; hi >= 0; hi -= 32) {
                const int i = hi - lane;
                double z = 0;
                if (i >= 0) {
                    const double u = a.logZ[w0 + i];
                    const bool fin = i >= i_final, renorm = i > 0 && (i % RS) == 0;
                    z = log(u);
                    if (i == 0) z += fin ? 0.0 : l_om;
                    else z -= (renorm ? 0.0 : log(a.logZ[w0 + i - 1])) + (fin ? l_om : 0.0);
                }
                __syncwarp();
                if (i >= 0) a.logZ[w0 + i] = z;
                __syncwarp();
            }
        }

        // =========================== backward ===========================
        Real* wbuf = st_a;  // [K][wl] state-major (the reference's categorical order a = v*len + j); spans both rings
        const int wcap = 2 * NS * IT;
        int cur = n - 1, vplus = -1, cnt = 0, u_hi = -1;
        int cache_top = -1, jl_c = 0;
        double w_c = 0;
        while (true) {
            const int len = cur + 1;
            const Real* row = msg + (int64_t)K * cur * (cur + 1) / 2;
            int lo;
            double wnb;
            {
                const int d = cache_top - 1 - cur;
                if (cache_top >= 0 && d < 32) {
                    lo = __shfl_sync(FULL, jl_c, d);
                    wnb = __shfl_sync(FULL, w_c, d);
                } else {
                    lo = jlo_arr[cur];
                    wnb = Wg[cur + 1];
                }
                const int r = cur - 1 - lane;
                jl_c = r >= 0 ? jlo_arr[r] : 0;
                w_c = r >= 0 ? Wg[r + 1] : 0.0;
                cache_top = cur;
            }
            const int wl = len - lo;
            const int L = K * wl;
            if (L > wcap) {  // cannot happen while the window fits the ring; keeps the smem writes in bounds
                overflow = true;
                break;
            }
            if (!a.uniforms && (u_hi < 0 || cur <= u_hi - 32)) {
                if (cur - lane >= 0)
                    ubuf[lane] = philox_u01(a.seed, TAG_BACKWARD, (uint64_t)(cur - lane), 0u, a.sweep,
                                            a.chain_base + (uint32_t)c);
                u_hi = cur;
            }
            // weights in item order t = (j - lo) * K + v (the scratch row's own order: coalesced), stored state-major.
            // First draw on a final row: w = alpha = g; jump draws: w = g A_{v,vplus} (rows with the B_v factor) or
            // alpha A_{v,vplus} / A_v (final rows); a grid that does not end at t_end: first draw w = g A_v
            const bool hot = vplus >= 0 && cur < i_final;
            if (hot) {
                constexpr int UB = 4;
                for (int t0 = lane; t0 < L; t0 += 32 * UB) {
                    Real rv[UB];
                    double wj[UB];
                    int jj[UB], vv[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        const int t = min(t0 + 32 * u, L - 1);
                        jj[u] = t / K;
                        vv[u] = t - jj[u] * K;
                        rv[u] = row[(int64_t)lo * K + t];
                        wj[u] = Wg[lo + jj[u]];
                    }
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        const Real L2 = FL::run(ex2, (Real)(wnb - wj[u]));
                        const Real w = rv[u] * FL::hazard(ex2, skm1[vv[u] * K + vplus] * L2 - sc2p[vv[u] * K + vplus]);
                        if (t0 + 32 * u < L) wbuf[vv[u] * wl + jj[u]] = w;
                    }
                }
            } else {
                for (int t = lane; t < L; t += 32) {
                    const int j = t / K, vs = t - j * K;
                    Real w = row[(int64_t)lo * K + t];
                    if (vplus >= 0 || cur < i_final) {
                        const Real L2 = ex2.log2((Real)(wnb - Wg[lo + j]));
                        Real ds = 0, hto = 0;
                        for (int s = 0; s < K; ++s) {
                            const Real kE = ex2(skm1[vs * K + s] * L2 - sc2p[vs * K + s]);
                            ds += kE;
                            hto = (s == vplus) ? kE : hto;
                        }
                        if (vplus < 0) w *= ds;
                        else w *= ds > (Real)0 ? hto * M::rcp(ds) : (Real)0;
                    }
                    wbuf[vs * wl + j] = w;
                }
            }
            __syncwarp();
            {   // pull the live parts of rows cur-1 .. cur-3 into L2, one draw ahead
                const int r = cur - 1 - lane % 3, line = lane / 3;
                const int rlo = __shfl_sync(FULL, jl_c, lane % 3);
                if (r >= 0) {
                    const Real* rb0 = msg + (int64_t)K * r * (r + 1) / 2;
                    for (int ln = line; ln < 64; ln += 10) {
                        const char* p = (const char*)(rb0 + (int64_t)K * rlo) + (size_t)ln * 128;
                        if (p < (const char*)(rb0 + (int64_t)K * (r + 1))) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                    }
                }
            }
            const int q = (L + 31) / 32;
            const int beg = min(lane * q, L), end = min(beg + q, L);
            Acc loc = 0;
            for (int t = beg; t < end; ++t) loc += (Acc)wbuf[t];
            const Acc incl = warp_inclusive_scan(loc, lane);
            Acc excl = __shfl_up_sync(FULL, incl, 1);
            if (lane == 0) excl = 0;
            const Acc total = __shfl_sync(FULL, incl, 31);
            if (!(total > (Acc)0) || !(total < (Acc)INFINITY)) {
                degenerate = true;
                break;
            }
            const double u = a.uniforms ? a.uniforms[w0 + cur] : ubuf[u_hi - cur];
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
                sel = sel >= 0 ? sel : lastpos;
            }
            unsigned own = __ballot_sync(FULL, sel >= 0);
            if (!own) {
                // no owner (u * total rounded up to the total, or a scan that is not monotone to the last bit): the
                // last lane with mass takes its last positive entry, as the reference's inverse CDF does at u -> 1
                const unsigned hm = __ballot_sync(FULL, has);
                if (!hm) {
                    degenerate = true;
                    break;
                }
                const int top = 31 - __clz(hm);
                if (lane == top) {
                    int lastpos = beg;
                    for (int t = beg; t < end; ++t) if ((Acc)wbuf[t] > (Acc)0) lastpos = t;
                    sel = lastpos;
                }
                own = 1u << top;
            }
            const int idx = __shfl_sync(FULL, sel, __ffs(own) - 1);
            const int vs = idx / wl, js = lo + idx - vs * wl;
            __syncwarp();
            if (a.aug_idx)
                for (int t = js + lane; t <= cur; t += 32) a.aug_idx[w0 + t] = vs * n + js;
            if (lane == 0) {
                soj[cnt] = vs;
                soj[(ncap + 1) + cnt] = js;
            }
            ++cnt;
            if (js == 0) break;
            vplus = vs;
            cur = js - 1;
        }
        __syncwarp();
        if (degenerate || overflow) {
            if (lane == 0) {
                a.status[c] = overflow ? (SMJP_ST_OVERFLOW | SMJP_ST_DEGENERATE) : SMJP_ST_DEGENERATE;
                a.path_len[c] = 0;
            }
            continue;
        }
        // ---- path compaction (smjp_utils.py:319-330), end point (:338-348), metrics (mcmc_utils.py:40-56) ----
        for (int p = lane; p < cnt; p += 32) {
            a.path_v[w0 + p] = soj[cnt - 1 - p];
            a.path_t[w0 + p] = Wg[soj[(ncap + 1) + cnt - 1 - p]];
        }
        const int v_last = soj[0];
        const double t_last = Wg[soj[ncap + 1]];
        const bool add_end = !np_isclose(t_last, a.t_end);
        const int plen = cnt + (add_end ? 1 : 0);
        if (lane == 0) {
            if (add_end) {
                a.path_v[w0 + cnt] = v_last;
                a.path_t[w0 + cnt] = a.t_end;
            }
            a.path_len[c] = plen;
            a.status[c] = 0;
        }
        if (lane < K && (a.time_in_state || a.jumps)) {
            double tsum = 0;
            int jc = 0;
            for (int p = 0; p < plen; ++p) {
                const int vs = p < cnt ? soj[cnt - 1 - p] : v_last;
                if (vs == lane) {
                    ++jc;
                    if (p + 1 < plen) {
                        const double t0 = Wg[soj[(ncap + 1) + cnt - 1 - p]];
                        const double t1 = (p + 1 < cnt) ? Wg[soj[(ncap + 1) + cnt - 2 - p]] : a.t_end;
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

}  // namespace smjp
