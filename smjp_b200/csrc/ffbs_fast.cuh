// FFBS, one CTA per chain, K <= 3, SCRATCH MODE ONLY: the kernel the Rao-Teh sweep runs (and bench.py times).
//
// Same reference lines, arithmetic, row layout and outputs as ffbs.cuh (the export / safe kernel).  What differs is
// how a grid step is organised -- the forward pass of ffbs.cuh ends every step in a block-wide reduction, a
// __syncthreads() and a division that every thread waits for; ncu (profiles/r01j_*): 1.29 warps stalled at the
// barrier per issue, 62 % issue-active.  Here:
//
//  * Normalisation is deferred by RS = 8 steps.  Between two renormalisations the state carries the unnormalised
//    message on a drifting scale; the scale factor 1/U of a renormalisation is folded into the observation factor of
//    the next step (thread-uniform), so no pass over the state is needed.  Only every RS-th step ends in
//    __syncthreads(); the live-window judgement rides on the same barrier.
//  * On the other steps the ONLY cross-warp dependency is the inflow J_{i+1}(v) of the new state (v, W[i+1]), and
//    only its owner thread needs it, for the last of its grid pairs of step i+1.  Warps publish their partial sums
//    (a ring of 16 slots), non-owner warps ARRIVE on a named barrier (bar.arrive, ids 1..7 by step parity) and go on
//    to step i+1; the owner warp alone SYNCs on it, sums the partials and seeds the new state in its own column of
//    the state arrays.  Warps may run up to RS-1 steps apart.
//  * The state arrays are a RING of NS slots of B grid points (the exact live window is ~250-300 points on cfg2):
//    shared memory no longer grows with the grid, a window that outgrows the ring flags the chain for the rescue
//    pass (ffbs.cuh), exactly like ffbs_warp.cuh -- results never depend on NS.
//  * The new state enters through the state arrays (message = J / (Omega-1), H = 0): no select per pair.  Omega
//    cancels out of the recursion: the state holds g * A_v (not g * Omega * A_v) and the observation factor carries
//    (Omega - 1) = thin * Omega.
//  * fp64 2^x without clamps and log2 without special cases: the prologue checks the range of the chain's hazard
//    exponents over [log2(smallest gap), log2(span)]; a chain outside (never on the BASELINE configurations) goes to
//    the rescue pass.
//
// Scale bookkeeping.  U_i = sum of the state after step i.  With rn_i the factor folded into step
// i (1 / U_{i-1} after a renormalisation, else 1) the reference's normaliser is
//     Z_0 = Omega U_0,   Z_i = U_i / (rn_i U_{i-1})           (rows with the B_v factor in the emission)
//     Z_0 = U_0,         Z_i = U_i / (Omega rn_i U_{i-1})     (final rows, smjp_utils.py:544-547)
// Rows g_i(v,j) go to the per-CTA scratch on whatever scale the step had: the backward pass only needs ratios
// within a row.
#pragma once
#include "ffbs.cuh"

namespace smjp {

#define PC(w, t) (sizeof(Real) == 8 ? (Real)a.pc_d[w][t] : (Real)a.pc_f[w][t])  // as in ffbs.cuh (K <= 3)

// elements of Real that hold the window starts [ncap + 2] (ints) inside a scratch region (api.cu sizes it the same way)
template <typename Real>
__host__ __device__ inline int64_t ffbs_fast_jlo_elems(int64_t ncap) {
    return ((ncap + 2) * 4 + (int64_t)sizeof(Real) - 1) / (int64_t)sizeof(Real);
}

constexpr int FAST_RS = 8;   // steps between block-wide synchronisations (renormalisation, window judgement)
constexpr int FAST_RR = 16;  // ring of partial-sum slots (> RS + the steps a warp can be ahead)

// dynamic shared memory of ffbs_fast_kernel for (Real, K, ring slots, ncap, threads)
template <typename Real>
__host__ __device__ inline FfbsLayout ffbs_fast_layout(int K, int ring, int ncap, int threads) {
    size_t state = (size_t)ring * K * threads * sizeof(Real);
    const size_t soj = (size_t)2 * (ncap + 1) * sizeof(int);  // backward: sojourn list lives in the H ring
    FfbsLayout L;
    size_t b = 0;
    L.wd = (int)b;   b += align_up((size_t)ring * threads * sizeof(double), 16);  // grid times of the window (ring)
    L.ubuf = (int)b; b += align_up((size_t)threads * sizeof(double), 16);
    L.scan = (int)b; b += align_up((size_t)(threads / 32 + 1) * sizeof(double), 16);
    L.sta = (int)b;  b += align_up(state, 16);
    L.sth = (int)b;  b += align_up(state > soj ? state : soj, 16);
    L.red = (int)b;  b += align_up((size_t)FAST_RR * (threads / 32) * (K + 1) * sizeof(Real), 16);
    L.par = (int)b;  b += align_up((size_t)(4 * K * K) * sizeof(Real), 16);
    L.jlo = (int)b;  // (the window start of every row lives in global scratch, behind the observation factors)
    L.total = (int)b;
    return L;
}

// the backward-only kernel (PHASE 2): uniforms, the weights of one draw (K x window), the per-warp totals, the
// hazard tables -- no state rings, no grid ring; the sojourn list lives in the chain's global scratch region
template <typename Real>
__host__ __device__ inline FfbsLayout ffbs_fast_bwd_layout(int K, int ring, int threads) {
    FfbsLayout L;
    size_t b = 0;
    L.wd = (int)b;
    L.ubuf = (int)b; b += align_up((size_t)threads * sizeof(double), 16);
    L.scan = (int)b; b += align_up((size_t)(threads / 32 + 1) * sizeof(double), 16);
    L.sta = (int)b;  b += align_up((size_t)ring * K * threads * sizeof(Real), 16);
    L.sth = (int)b;
    L.red = (int)b;  b += align_up((size_t)FAST_RR * (threads / 32) * (K + 1) * sizeof(Real), 16);
    L.par = (int)b;  b += align_up((size_t)(4 * K * K) * sizeof(Real), 16);
    L.jlo = (int)b;
    L.total = (int)b;
    return L;
}
// ints behind the observation factors of a chain's scratch region in split mode: window starts [ncap + 2], sojourn
// list [2][ncap + 1]
__host__ __device__ inline int64_t ffbs_fast_split_ints(int64_t ncap) { return 3 * ncap + 4; }

// Named barriers with IMMEDIATE ids: with a run-time id ptxas reserves all 16 hardware barriers for the CTA, and an
// SM has 64 of them -- 4 resident CTAs whatever the registers allow (measured: grid 592 instead of 740).  Ids 1..7.
template <int ID>
__device__ __forceinline__ void named_bar_arrive_c(int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}
template <int ID>
__device__ __forceinline__ void named_bar_sync_c(int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar(int id, bool sync, int nthreads) {  // id in 1..7, block-uniform per warp
#define SMJP_BAR_CASE(I)                                  \
    case I:                                               \
        if (sync) named_bar_sync_c<I>(nthreads);          \
        else named_bar_arrive_c<I>(nthreads);             \
        break;
    switch (id) {
        SMJP_BAR_CASE(1)
        SMJP_BAR_CASE(2)
        SMJP_BAR_CASE(3)
        SMJP_BAR_CASE(4)
        SMJP_BAR_CASE(5)
        SMJP_BAR_CASE(6)
        default:
            if (sync) named_bar_sync_c<7>(nthreads);
            else named_bar_arrive_c<7>(nthreads);
    }
#undef SMJP_BAR_CASE
}

// ---- TMA 1-D bulk copy (cp.async.bulk) of a chain's observation slice into shared memory, completion on an mbarrier ----
// The slice is staged in the (still unused) state ring while the warp validates the grid, so that the per-interval
// binary searches and symbol reads of the prologue hit shared memory instead of L2.  Source, destination and size must
// be multiples of 16 bytes: the copy starts at the 16-byte boundary below the slice and ends at the one above it.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// plan of the staged slice: element offsets of the 16-byte aligned windows around [o0, o0 + D)
struct ObsStage {
    int64_t t0, y0;      // first element copied (obs_t: even, obs_y: multiple of 4)
    uint32_t bt, by;     // bytes copied
    int dt, dy;          // position of element o0 inside the copies
};
__device__ __forceinline__ ObsStage obs_stage_plan(int64_t o0, int D) {
    ObsStage p;
    p.dt = (int)(o0 & 1);
    p.dy = (int)(o0 & 3);
    p.t0 = o0 - p.dt;
    p.y0 = o0 - p.dy;
    p.bt = (uint32_t)(((D + p.dt + 1) & ~1) * 8);
    p.by = (uint32_t)(((D + p.dy + 3) & ~3) * 4);
    return p;
}

// log2 of a positive NORMAL double (the prologue's range check guarantees it): Exp2<double>::log2 without the
// special-case branch and without the int -> double conversion (exponent through the 2^52 trick)
template <typename Real>
struct FastLog2;
template <>
struct FastLog2<double> {
    static __device__ __forceinline__ double run(const Exp2<double>& e, double x) {
        const int hi = __double2hiint(x);
        const double2 tv = e.ltab[(hi >> 13) & 127];
        const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
        const double r = fma(m, tv.x, -1.0);
        double p = fma(r, c_log2[5], c_log2[4]);
        p = fma(r, p, c_log2[3]);
        p = fma(r, p, c_log2[2]);
        p = fma(r, p, c_log2[1]);
        p = fma(r, p, c_log2[0]);
        // (hi >> 20) - 1023 as a double: 2^52 + 2^31 + E has E ^ 0x80000000 in its low word
        const double ed = __hiloint2double(0x43300000, (hi >> 20) ^ 0x80000000) - (4503599627370496.0 + 2147483648.0 + 1023.0);
        return fma(r, p, ed + tv.y);
    }
    static __device__ __forceinline__ double hazard(const Exp2<double>& e, double xs) { return e.core<false>(xs); }
};
template <>
struct FastLog2<float> {
    static __device__ __forceinline__ float run(const Exp2<float>&, float x) { return Math<float>::log2(x); }
    static __device__ __forceinline__ float hazard(const Exp2<float>&, float xs) { return Math<float>::exp2(xs); }
};

// B threads per chain, CPB chains per CTA (CPB > 1: one warp per chain, the warps of a CTA are independent and share
// only the 2^x / log2 tables), WAYS grid pairs in flight per thread
//
// PHASE 0: forward and backward of a chain in one go (scratch regions per resident chain).  PHASE 1 / 2: the same code as
// TWO kernels -- forward pass of every chain (rows, window starts and observation factors in a scratch region PER
// CHAIN), then the backward pass of every chain.  The forward pass is FP64-pipe bound and needs ~230 registers; the
// backward pass is a chain of dependent draws (latency bound) and needs half of them: on its own it runs with twice
// the resident warps, and the forward kernel no longer shares its SMs with warps that wait on L2.
template <typename Real, int K, int B, int MINB, int WAYS, int CPB, int PHASE = 0>
__global__ void __launch_bounds__(B * CPB, MINB) ffbs_fast_kernel(const FfbsArgs a) {
    using M = Math<Real>;
    using Acc = typename ScanAcc<Real>::type;
    using FL = FastLog2<Real>;
    constexpr int NW = B / 32;
    constexpr int RS = FAST_RS, RR = FAST_RR;
    static_assert(RS == 8 && RR >= 2 * RS, "named barrier ids 1..7 and the partial-sum ring assume RS = 8");
    static_assert(CPB == 1 || B == 32, "several chains per CTA: one warp each");
    static_assert(PHASE == 0 || (B == 32 && CPB == 1), "split kernels: one warp per chain");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.guard && (a.guard[0] > a.guard_nmax || a.guard[1] > a.guard_total)) return;  // mis-speculated geometry
    const int cw = CPB > 1 ? (int)threadIdx.x / B : 0;                  // chain slot of this thread within the CTA
    const int tid = CPB > 1 ? (int)threadIdx.x % B : (int)threadIdx.x;  // thread index within the chain's group
    const int lane = tid & 31, wid = tid >> 5;
    const int ncap = a.ncap, Kobs = a.Kobs, NS = a.nslot;  // nslot = ring slots of B grid points
    const int region0 = blockIdx.x * CPB + cw;             // scratch region (messages, observation factors): per resident chain
    auto sync_chain = [&]() {  // every thread working on this chain
        if (CPB > 1) __syncwarp();
        else __syncthreads();
    };

    unsigned char* sm = smem_raw + (size_t)cw * a.lay.total;  // this chain's region of the dynamic shared memory
    double* Wr = (double*)(sm + a.lay.wd);  // W[j] of the live window: ring [slot][B], written when (v, W[j]) is seeded
    double* ubuf = (double*)(sm + a.lay.ubuf);
    Real* st_a = (Real*)(sm + a.lay.sta);  // forward: state ring [slot][v][B]; backward: this thread's weights
    Real* st_h = (Real*)(sm + a.lay.sth);  // forward: H_v(tau) ring; backward: sojourn list
    Real* red = (Real*)(sm + a.lay.red);   // [RR][NW][K+1] partial sums of (U, J[0..K))
    Real* srk = (Real*)(sm + a.lay.par);
    Real* skm1 = srk + K * K;
    Real* sc2p = skm1 + K * K;
    Real* sk = sc2p + K * K;
    int* soj = (int*)st_h;  // (PHASE 2: re-pointed to the chain's global scratch region)

    __shared__ int s_chain_a[CPB];
    __shared__ int s_sel_a[CPB], s_selv_a[CPB];
    __shared__ int s_found_a[CPB][2];
    __shared__ int s_jmin_a[CPB][2 * NW];
    __shared__ int s_flag_a[CPB];
    __shared__ unsigned long long s_dmin_a[CPB];
    __shared__ __align__(8) unsigned long long s_obar_a[CPB];  // completion of the observation bulk copy
    __shared__ double s_tab[Exp2<Real>::TAB ? Exp2<Real>::TAB : 1];
    __shared__ double2 s_ltab[Exp2<Real>::LTAB ? Exp2<Real>::LTAB : 1];
    int& s_chain = s_chain_a[cw];
    int& s_sel = s_sel_a[cw];
    int& s_selv = s_selv_a[cw];
    int* s_found = s_found_a[cw];
    int* s_jmin = s_jmin_a[cw];
    int& s_flag = s_flag_a[cw];
    unsigned long long& s_dmin = s_dmin_a[cw];
    unsigned long long* s_obar = &s_obar_a[cw];
    uint32_t obs_parity = 0;
    if (tid == 0) mbar_init(s_obar, 1);

    for (int t = tid; t < K * K; t += B) {
        const double k = a.model[t];
        sk[t] = (Real)k;
        srk[t] = (Real)(1.0 / k);
        skm1[t] = (Real)((k - 1.0) * (double)Exp2<Real>::SCALE);
        sc2p[t] = (Real)((a.model[K * K + t] - ::log2(k)) * (double)Exp2<Real>::SCALE);
    }
    for (int t = threadIdx.x; t < Exp2<Real>::TAB; t += B * CPB) s_tab[t] = exp2_table_entry(t);
    for (int t = threadIdx.x; t < Exp2<Real>::LTAB; t += B * CPB) {
        const double rc = 1.0 / (1.0 + ((double)t + 0.5) / 128.0);
        s_ltab[t] = make_double2(rc, -::log2(rc));
    }
    Exp2<Real> ex2;
    ex2.tab = s_tab;
    ex2.ltab = s_ltab;
    const double* emis = a.model + 2 * K * K;

    const Real om1 = (Real)(a.omega - 1.0);                 // thin * Omega
    const Real c_new = (Real)(1.0 / (a.omega - 1.0));       // scale of a new state's message (header)
    const Real om_l2e = (Real)(a.omega * SMJP_LOG2E * (double)Exp2<Real>::SCALE);
    const Real power = (Real)a.power;
    const Real weps = (Real)a.window_eps;
    // per-region global scratch: observation factors [ncap][K], then the window start of every row [ncap + 1] (read
    // back once per backward draw: L2)
    const int64_t obs_total = a.obs_offs[a.C];  // entries of obs_t / obs_y: the aligned bulk copies stay inside them
    if (CPB > 1) __syncthreads();  // the tables; from here on the warps of the CTA go their own ways

    while (true) {
        sync_chain();
        if (tid == 0) {
            unsigned int q = atomicAdd(PHASE == 2 ? a.queue2 : a.queue, 1u);
            const unsigned limit = a.C_dev ? (unsigned)*a.C_dev : (unsigned)a.C;
            s_chain = (q < limit) ? (a.order ? a.order[q] : (int)q) : -1;
            s_flag = 0;
            s_dmin = 0x7ff0000000000000ull;  // +inf
        }
        sync_chain();
        const int c = s_chain;
        if (c < 0) break;
        if (PHASE == 2 && !a.fwd_ok[c]) continue;  // the forward kernel has set this chain's status
        const int64_t w0 = a.w_offs[c];
        const int n = (int)(a.w_offs[c + 1] - w0) - 1;
        if (n < 1 || n > ncap) {
            if (tid == 0) {
                a.status[c] = (n < 1) ? SMJP_ST_GRID_NOT_INCREASING : SMJP_ST_OVERFLOW;
                a.path_len[c] = 0;
            }
            continue;
        }
        // scratch region: of this resident chain (fused kernel) or of chain c itself (split kernels)
        const int64_t region = PHASE == 0 ? region0 : c;
        Real* obsf = (Real*)a.obsf_scratch + region * a.obsf_stride;
        int* jlo_arr = (int*)(obsf + (int64_t)ncap * K);
        if (PHASE == 2) soj = jlo_arr + (ncap + 2);  // the sojourn list: behind the window starts, not in shared memory
        const double* Wg = a.W + w0;  // the grid stays in global memory (L2); only the window's times go on chip
        Real* msg = (Real*)a.msg_scratch + region * a.msg_scratch_stride;
        int i_final = n;  // rows from i_final on have no B_v factor in the emission (smjp_utils.py:544)
        while (i_final > 0 && np_isclose(a.t_end, Wg[i_final])) --i_final;
        bool degenerate = false;
        if constexpr (PHASE != 2) {
        // ---- observations: TMA bulk copy of the chain's slice into the (idle) state ring, under the grid validation ----
        const int64_t o0 = a.obs_offs[c];
        const int D = (int)(a.obs_offs[c + 1] - o0);
        const ObsStage os = obs_stage_plan(o0, D);
        const bool staged = D > 0 && (size_t)os.bt + os.by <= 2 * (size_t)NS * K * B * sizeof(Real) &&
                            os.t0 + (int64_t)(os.bt / 8) <= obs_total && os.y0 + (int64_t)(os.by / 4) <= obs_total &&
                            ((reinterpret_cast<uintptr_t>(a.obs_t) | reinterpret_cast<uintptr_t>(a.obs_y)) & 15) == 0;
        double* so_t = (double*)st_a;                         // st_a and st_h are contiguous
        int32_t* so_y = (int32_t*)((char*)st_a + os.bt);
        if (staged && tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the ring was last written by ordinary stores
            mbar_expect_tx(s_obar, os.bt + os.by);
            bulk_g2s(so_t, a.obs_t + os.t0, os.bt, s_obar);
            bulk_g2s(so_y, a.obs_y + os.y0, os.by, s_obar);
        }
        // ---- validate the grid (smjp_utils.py:398-399), smallest gap ----
        {
            int bad = 0;
            double dmin = INFINITY;
            for (int t = tid; t < n; t += B) {
                const double gap = Wg[t + 1] - Wg[t];
                bad |= !(gap > 0.0);
                dmin = fmin(dmin, gap);
            }
            if (bad) atomicOr(&s_flag, 1);
            else if (dmin < INFINITY) atomicMin(&s_dmin, (unsigned long long)__double_as_longlong(dmin));  // positive doubles order as integers
        }
        // ---- observation factor per interval: (sum_{x in [W[i],W[i+1]]} P(x|v))^power (as in ffbs.cuh) ----
        {
            if (staged) {
                mbar_wait(s_obar, obs_parity);
                obs_parity ^= 1u;
            }
            const double* ot = staged ? so_t + os.dt : a.obs_t + o0;
            const int32_t* oy = staged ? so_y + os.dy : a.obs_y + o0;
            for (int i = tid; i < n; i += B) {
                const double wc = Wg[i], wn = Wg[i + 1];
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
        sync_chain();
        if (s_flag) {
            if (tid == 0) {
                a.status[c] = ((s_flag & 1) ? SMJP_ST_GRID_NOT_INCREASING : 0) | ((s_flag & 2) ? SMJP_ST_BAD_INPUT : 0);
                a.path_len[c] = 0;
            }
            continue;
        }
        if (sizeof(Real) == 8) {
            // range of the unclamped fp64 arithmetic: every hazard exponent (k-1) log2 tau - c over
            // tau in [smallest gap, span] within +-900, gaps normal, span below 2^100
            const double dmin = __longlong_as_double((long long)s_dmin), span = Wg[n] - Wg[0];
            bool ok = dmin >= 1.0e-270 && span <= 1.0e30;
            if (ok) {
                const double l0 = ::log2(dmin), l1 = ::log2(span);
                for (int t = 0; t < K * K; ++t) {
                    const double km1 = a.pc_d[1][t] * (1.0 / 256.0), cc = a.pc_d[2][t] * (1.0 / 256.0);
                    ok = ok && fabs(km1 * l0 - cc) <= 900.0 && fabs(km1 * l1 - cc) <= 900.0;
                }
            }
            if (!ok) {  // block-uniform: this chain is redone by the rescue pass (ffbs.cuh, clamped arithmetic)
                if (tid == 0) {
                    a.status[c] = SMJP_ST_OVERFLOW | SMJP_ST_DEGENERATE;
                    a.path_len[c] = 0;
                }
                continue;
            }
        }


        // =========================== forward ===========================
        bool overflow = false;
        int j_lo = 0, ms_lo = 0;  // first live grid point; ring slot of its slot m_lo = j_lo / B
        unsigned long long live = 0;
        Real sc = om1;            // (Omega - 1) * renormalisation factor of the coming step
        Real ofn[K];              // observation factors of the coming step (prefetched)
#pragma unroll
        for (int v = 0; v < K; ++v) ofn[v] = obsf[v];
        const Real* obsf_next = obsf + K;
        Real* row_i = msg;        // start of row i (K * (i + 1) values, entry (v, j) at v * (i + 1) + j)
        double wn_next = Wg[1];   // right end of the coming interval (prefetched one step ahead)
        if (tid == 0) {           // state (v, W[0]): uniform prior (smjp_utils.py:286-289), H = 0
            Wr[0] = Wg[0];
#pragma unroll
            for (int v = 0; v < K; ++v) {
                st_a[v * B] = (Real)(1.0 / K) * c_new;
                st_h[v * B] = (Real)0;
            }
        }

        auto step = [&](auto final_tag, const int i) -> bool {
            constexpr bool FINAL = decltype(final_tag)::value;
            const int mi = i / B, ti = i - mi * B;              // slot and owner thread of the newest state (v, W[i])
            const int m_lo = j_lo / B, t_lo = j_lo - m_lo * B;  // slot / thread of the first live grid point
            if (mi - m_lo >= NS) {                              // the window outgrew the ring (block-uniform)
                overflow = true;
                return false;
            }
            const double wn = wn_next;
            wn_next = i + 2 <= n ? Wg[i + 2] : 0.0;
            if (tid == 0) jlo_arr[i] = j_lo;
            live += (unsigned long long)(i - j_lo + 1);
            Real oft[K];
#pragma unroll
            for (int v = 0; v < K; ++v) oft[v] = ofn[v] * sc;
#pragma unroll
            for (int v = 0; v < K; ++v) ofn[v] = (i + 1 < n) ? obsf_next[v] : (Real)1;
            obsf_next += K;
            Real Zp = 0;
            Real Jp[K];
#pragma unroll
            for (int v = 0; v < K; ++v) Jp[v] = 0;
            // WAYS grid pairs per thread in flight: the pair code is a chain of dependent fp64 operations (2^x after 2^x),
            // one pair at a time leaves a warp waiting on its own results (ncu: "wait" is half of the stall samples of
            // this loop); WAYS independent pairs interleave.  Slot m + p of the iteration lives in ring slot
            // (ms + p) mod NS; a pair outside the window computes on tau = 1 with a zero message and stores nothing.
            int ms = ms_lo;
            Real* rowg = row_i + m_lo * B + tid;
            // slots m .. m + NWAY - 1 of this thread: all K*K hazards, emission, state update, jump contributions
            auto pairs = [&](auto nway_tag, const int m) {
                constexpr int NWAY = decltype(nway_tag)::value;
                bool act[NWAY];
                int so[NWAY];  // element offset of this thread's column in ring slot (ms + p) mod NS
                Real tau[NWAY], L2[NWAY];
#pragma unroll
                for (int p = 0; p < NWAY; ++p) {
                    const int mm = m + p;
                    act[p] = (mm > m_lo || tid >= t_lo) && (mm < mi || tid <= ti);
                    int sp = ms + p;
                    sp = sp >= NS ? sp - NS : sp;
                    so[p] = sp * (K * B) + tid;
                    tau[p] = act[p] ? (Real)(wn - Wr[sp * B + tid]) : (Real)1;
                }
                if (B > 32) {  // warps whose lanes are all outside the window have nothing to do in these slots
                    bool any = false;
#pragma unroll
                    for (int p = 0; p < NWAY; ++p) any = any || act[p];
                    if (!__any_sync(0xffffffffu, any)) {
                        rowg += NWAY * B;
                        ms += NWAY;
                        ms = ms >= NS ? ms - NS : ms;
                        return;
                    }
                }
#pragma unroll
                for (int p = 0; p < NWAY; ++p) L2[p] = FL::run(ex2, tau[p]);
#pragma unroll
                for (int v = 0; v < K; ++v) {
                    Real kE[NWAY][K], esum[NWAY], dsum[NWAY], ap[NWAY], hp[NWAY], hsum[NWAY], g[NWAY];
#pragma unroll
                    for (int p = 0; p < NWAY; ++p) {
                        ap[p] = act[p] ? st_a[so[p] + v * B] : (Real)0;
                        hp[p] = act[p] ? st_h[so[p] + v * B] : (Real)0;
                        esum[p] = 0;
                        dsum[p] = 0;
                    }
#pragma unroll
                    for (int sx = 0; sx < K; ++sx) {
#pragma unroll
                        for (int p = 0; p < NWAY; ++p) {
                            kE[p][sx] = FL::hazard(ex2, PC(1, v * K + sx) * L2[p] - PC(2, v * K + sx));  // A_vs(tau)
                            esum[p] += kE[p][sx] * PC(0, v * K + sx);
                            dsum[p] += kE[p][sx];
                        }
                    }
#pragma unroll
                    for (int p = 0; p < NWAY; ++p) {
                        hsum[p] = tau[p] * esum[p];  // H_v(tau)
                        g[p] = (oft[v] * ex2.neg(-om_l2e * (hsum[p] - hp[p]))) * ap[p];
                    }
#pragma unroll
                    for (int p = 0; p < NWAY; ++p) {
                        Real gj = g[p], an = g[p] * dsum[p];
                        if (FINAL) {  // no B_v in the emission: the state is g itself, the jump term keeps its 1 / A_v
                            an = g[p];
                            gj = dsum[p] > (Real)0 ? g[p] * M::rcp(dsum[p]) : (Real)0;
                        }
                        if (act[p]) {
                            rowg[p * B + v * (i + 1)] = g[p];  // scratch row i: the backward jump weight is g A_{v,v+}
                            st_a[so[p] + v * B] = an;
                            st_h[so[p] + v * B] = hsum[p];
                        }
#pragma unroll
                        for (int sx = 0; sx < K; ++sx) Jp[sx] += gj * kE[p][sx];
                        Zp += an;
                    }
                }
                rowg += NWAY * B;
                ms += NWAY;
                ms = ms >= NS ? ms - NS : ms;
            };
            // WAYS grid pairs per thread in flight: the pair code is a chain of dependent fp64 operations (2^x after
            // 2^x); one pair at a time leaves a warp waiting on its own results (ncu: "wait" is half of the stall
            // samples of this loop).  A pair outside the window (lanes of the first and last slot) computes on tau = 1
            // with a zero message and stores nothing; the slots left over by the WAYS-wide loop go one at a time.
            {
                int m = m_lo;
                if (WAYS > 1) {
#pragma unroll 1
                    for (; m + WAYS - 1 <= mi; m += WAYS) pairs(std::integral_constant<int, WAYS>(), m);
                }
#pragma unroll 1
                for (; m <= mi; ++m) pairs(std::integral_constant<int, 1>(), m);
            }
            const int t_new = (i + 1) % B;  // owner thread of the new state (v, W[i+1])
            const bool sync_step = ((i + 1) % RS) == 0 || i == n - 1;
            Real U = 0, Js[K];
            bool mine = false;  // this thread seeds the new state
            // live-window judgement on the state after step i: smallest grid point of the oldest slot that still
            // carries mass (exact mode: non-zero; pruned mode: above window_eps of the last renormalised total)
            auto my_oldest_live = [&]() {
                int mylive = 0x7fffffff;
                const int j0 = m_lo * B + tid;
                if (j0 >= j_lo && j0 <= i) {
                    const Real* q = st_a + (size_t)ms_lo * K * B + tid;
                    bool alive = false;
#pragma unroll
                    for (int v = 0; v < K; ++v) alive = alive || q[v * B] > weps;
                    if (alive) mylive = j0;
                }
                return __reduce_min_sync(0xffffffffu, mylive);
            };
            auto advance_window = [&](int jl) {
                const int j_new = max(j_lo, min(jl, min((m_lo + 1) * B, i + 1)));
                if (j_new / B != m_lo) ms_lo = (ms_lo + 1 == NS) ? 0 : ms_lo + 1;
                j_lo = j_new;
            };
            constexpr int NVP = pow2_ceil(K + 1), SH = 5 - log2_of(NVP);
            Real vals[NVP];
            vals[0] = Zp;
#pragma unroll
            for (int v = 0; v < K; ++v) vals[1 + v] = Jp[v];
#pragma unroll
            for (int v = K + 1; v < NVP; ++v) vals[v] = 0;
            WarpHalve<Real, NVP, 16>::run(vals, lane);  // lane (idx << SH) holds the warp total of value idx
            if (B == 32) {
                // one warp per chain: the totals reach every lane by shuffle; no shared memory, no barrier
                U = __shfl_sync(0xffffffffu, vals[0], 0);
#pragma unroll
                for (int v = 0; v < K; ++v) Js[v] = __shfl_sync(0xffffffffu, vals[0], (v + 1) << SH);
                if (sync_step) {
                    if (!(U > (Real)0) || !(U < (Real)INFINITY)) {
                        degenerate = true;
                        return false;
                    }
                    sc = om1 / U;  // renormalisation, folded into the next step's observation factors
                    advance_window(my_oldest_live());
                } else {
                    sc = om1;
                }
                mine = tid == t_new;
            } else {
                // partial sums of (U, J[0..K)) of this warp -> ring slot of this step
                Real* rb = red + (size_t)(i % RR) * NW * (K + 1);
                {
                    const int idx = lane >> SH;
                    if ((lane & ((1 << SH) - 1)) == 0 && idx < K + 1) rb[wid * (K + 1) + idx] = vals[0];
                }
                auto total = [&]() {
#pragma unroll
                    for (int v = 0; v < K; ++v) Js[v] = 0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        Real part[K + 1];
                        lds_vec<Real, K + 1>(rb + w * (K + 1), part);
                        U += part[0];
#pragma unroll
                        for (int v = 0; v < K; ++v) Js[v] += part[1 + v];
                    }
                };
                if (sync_step) {
                    const int par = ((i + 1) / RS) & 1;
                    const int mylive = my_oldest_live();
                    if (lane == 0) s_jmin[par * NW + wid] = mylive;
                    sync_chain();
                    total();
                    if (!(U > (Real)0) || !(U < (Real)INFINITY)) {
                        degenerate = true;
                        return false;
                    }
                    sc = om1 / U;  // renormalisation, folded into the next step's observation factors
                    int jl = 0x7fffffff;
#pragma unroll
                    for (int w = 0; w < NW; ++w) jl = min(jl, s_jmin[par * NW + w]);
                    advance_window(jl);
                    mine = tid == t_new;
                } else {
                    sc = om1;
                    const int id = 1 + (i % RS);
                    const bool owner_warp = wid == (t_new >> 5);
                    named_bar(id, owner_warp, B);
                    mine = tid == t_new;
                    if (mine) total();
                }
            }
            if (mine) {
                if (a.logZ) a.logZ[w0 + i] = (double)U;  // raw totals; turned into log Z_i after the pass
                if (i + 1 < n) {
                    const int sl = ((i + 1) / B) % NS;
                    Real* qa = st_a + (size_t)sl * K * B + t_new;
                    Real* qh = st_h + (size_t)sl * K * B + t_new;
                    Wr[sl * B + t_new] = wn;  // W[i + 1]
#pragma unroll
                    for (int v = 0; v < K; ++v) {
                        qa[v * B] = Js[v] * c_new;
                        qh[v * B] = (Real)0;
                    }
                }
            }
            row_i += (int64_t)K * (i + 1);
            return true;
        };
        {
            int i = 0;
            bool go = true;
            for (; go && i < i_final; ++i) go = step(std::false_type(), i);
            for (; go && i < n; ++i) go = step(std::true_type(), i);
        }
        sync_chain();
        if (a.live_pairs && tid == 0) atomicAdd(a.live_pairs, live);
        if (degenerate || overflow) {
            if (tid == 0) {
                a.status[c] = overflow ? (SMJP_ST_OVERFLOW | SMJP_ST_DEGENERATE) : SMJP_ST_DEGENERATE;
                a.path_len[c] = 0;
            }
            continue;
        }
        if (a.logZ) {
            // raw totals U_i -> log Z_i (header), in place, B entries at a time from the end (entry i needs U_{i-1})
            __threadfence_block();
            const double l_om = log(a.omega);
            for (int hi = n - 1; hi >= 0; hi -= B) {
                const int i = hi - tid;
                double z = 0;
                if (i >= 0) {
                    const double u = a.logZ[w0 + i];
                    const bool fin = i >= i_final, renorm = i > 0 && (i % RS) == 0;
                    z = log(u);
                    if (i == 0) z += fin ? 0.0 : l_om;
                    else z -= (renorm ? 0.0 : log(a.logZ[w0 + i - 1])) + (fin ? l_om : 0.0);
                }
                sync_chain();
                if (i >= 0) a.logZ[w0 + i] = z;
                sync_chain();
            }
        }
        if (PHASE == 1) {  // forward kernel: the backward kernel takes the chain from here
            __threadfence();
            if (tid == 0) a.fwd_ok[c] = 1;
            continue;
        }
        }  // PHASE != 2
        if (tid == 0) s_found[0] = s_found[1] = 0;
        sync_chain();

        // =========================== backward (as in ffbs.cuh, scratch rows) ===========================
        // weights of the current draw, each entry written and read by its own thread only: entry t of thread's run of state
        // v at [(v q + t) B + tid] -- a thread's run is contiguous in j, so [v][j - lo] put the lanes of a warp 8 banks
        // apart (q ~ 12 doubles) and every store and walk load was an 8-way bank conflict (ncu r03g: short scoreboard
        // 2.9 stalled warps per issue in the backward-only kernel)
        Real* wbuf = st_a;
        int cur = n - 1, vplus = -1, cnt = 0, u_hi = -1;
        // One warp per chain: a draw is a chain of dependent L2 round trips (window start of the row -> row entries and
        // grid times -> weights).  The window start and the right grid time of the NEXT draw's row are fetched one draw
        // ahead for the 32 rows below the current one (lane l: row cache_top - 1 - l) and handed over by shuffle;
        // sojourns are a few grid steps long, so the next row is nearly always among them.
        int cache_top = -1, jl_c = 0;
        double w_c = 0;
        while (true) {
            const int len = cur + 1;
            const Real* row = msg + (int64_t)K * cur * (cur + 1) / 2;
            int lo;      // live window of this row: entries below lo were never written
            double wnb;  // W[cur + 1]
            {
                const int d = cache_top - 1 - cur;
                if (B == 32 && cache_top >= 0 && d < 32) {  // (d >= 0 always: cur only decreases)
                    lo = __shfl_sync(0xffffffffu, jl_c, d);
                    wnb = __shfl_sync(0xffffffffu, w_c, d);
                } else {
                    lo = jlo_arr[cur];
                    wnb = Wg[cur + 1];
                }
                if (B == 32) {
                    const int r = cur - 1 - lane;
                    jl_c = r >= 0 ? jlo_arr[r] : 0;
                    w_c = r >= 0 ? Wg[r + 1] : 0.0;
                    cache_top = cur;
                }
            }
            const int wl = len - lo;
            if (!a.uniforms && (u_hi < 0 || cur <= u_hi - B)) {  // refill B uniforms: steps cur, cur-1, ...
                if (cur - tid >= 0)
                    ubuf[tid] = philox_u01(a.seed, TAG_BACKWARD, (uint64_t)(cur - tid), 0u, a.sweep,
                                           a.chain_base + (uint32_t)c);
                u_hi = cur;
            }
            // mode 0: a_{n-1} ~ alpha_{n-1} (fast_hmm.py:124-125); mode 1: row holds g, w = g A_{v,vplus}(tau);
            // mode 2 (final rows, row holds alpha): w = alpha A_{v,vplus} / A_v (fast_hmm.py:153-171)
            const int mode = vplus < 0 ? 0 : (cur < i_final ? 1 : 2);
            const bool last_nonfinal = vplus < 0 && cur < i_final;  // a grid that does not end at t_end: alpha = g Omega A_v
            auto weights = [&](int j, Real (&w)[K]) {  // the rare modes: first draw, final rows
                Real rv[K];
#pragma unroll
                for (int v = 0; v < K; ++v) rv[v] = row[(int64_t)v * len + j];
                if (mode == 0 && !last_nonfinal) {
#pragma unroll
                    for (int v = 0; v < K; ++v) w[v] = rv[v];
                    return;
                }
                const Real L2 = ex2.log2((Real)(wnb - Wg[j]));
#pragma unroll
                for (int v = 0; v < K; ++v) {
                    Real dsum = 0, hto = 0;
#pragma unroll
                    for (int s2 = 0; s2 < K; ++s2) {
                        const Real kE = ex2(skm1[v * K + s2] * L2 - sc2p[v * K + s2]);
                        dsum += kE;
                        hto = (s2 == vplus) ? kE : hto;
                    }
                    if (mode == 0) {
                        w[v] = rv[v] * dsum;
                    } else {
                        const Real r = dsum > (Real)0 ? hto * M::rcp(dsum) : (Real)0;
                        w[v] = rv[v] * r;
                    }
                }
            };
            const int q = (wl + B - 1) / B;
            const int jb = min(lo + tid * q, len), je = min(jb + q, len);  // this thread's grid points
            Acc loc[K];
#pragma unroll
            for (int v = 0; v < K; ++v) loc[v] = 0;
            if (mode == 1) {
                // the common draw: w = g_cur(v, j) A_{v,vplus}(W[cur+1] - W[j]).  UB grid points at a time: their row
                // entries and grid times are requested together (one L2 latency per UB points instead of one per
                // point), their 2^x chains interleave
                constexpr int UB = 4;
                Real km1[K], c2v[K];
#pragma unroll
                for (int v = 0; v < K; ++v) {
                    km1[v] = skm1[v * K + vplus];
                    c2v[v] = sc2p[v * K + vplus];
                }
                for (int j0 = jb; j0 < je; j0 += UB) {
                    Real rv[UB][K];
                    double wj[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        const int j = min(j0 + u, je - 1);  // past the end: a harmless repeat of the last point
                        wj[u] = Wg[j];
#pragma unroll
                        for (int v = 0; v < K; ++v) rv[u][v] = row[(int64_t)v * len + j];
                    }
                    Real L2[UB], w[UB][K];
#pragma unroll
                    for (int u = 0; u < UB; ++u) L2[u] = FL::run(ex2, (Real)(wnb - wj[u]));
#pragma unroll
                    for (int v = 0; v < K; ++v) {
#pragma unroll
                        for (int u = 0; u < UB; ++u) w[u][v] = rv[u][v] * FL::hazard(ex2, km1[v] * L2[u] - c2v[v]);
                    }
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        if (j0 + u < je) {
#pragma unroll
                            for (int v = 0; v < K; ++v) {
                                loc[v] += (Acc)w[u][v];
                                wbuf[(v * q + (j0 + u - jb)) * B + tid] = w[u][v];
                            }
                        }
                    }
                }
            } else {
                for (int j = jb; j < je; ++j) {
                    Real w[K];
                    weights(j, w);
#pragma unroll
                    for (int v = 0; v < K; ++v) {
                        loc[v] += (Acc)w[v];
                        wbuf[(v * q + (j - jb)) * B + tid] = w[v];
                    }
                }
            }
            {   // pull the live parts of rows cur-1 .. cur-3 into L2, one draw ahead (ffbs.cuh); by now their window
                // starts have arrived (B == 32: lanes 0..2 of the cache)
                const int seg = tid % (3 * K), line = tid / (3 * K);
                const int r = cur - 1 - seg / K, v = seg - (seg / K) * K;
                const int rlo_c = B == 32 ? __shfl_sync(0xffffffffu, jl_c, seg / K) : 0;
                if (r >= 0) {
                    const int rlo = B == 32 ? rlo_c : jlo_arr[r];
                    const char* p = (const char*)(msg + (int64_t)K * r * (r + 1) / 2 + (int64_t)v * (r + 1) + rlo) + line * 128;
                    if (p < (const char*)(msg + (int64_t)K * r * (r + 1) / 2 + (int64_t)(v + 1) * (r + 1)))
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                }
            }
            Acc incl[K];
#pragma unroll
            for (int v = 0; v < K; ++v) incl[v] = warp_inclusive_scan(loc[v], lane);
            Acc* wsum = (Acc*)red;  // [NW][K] warp totals per state
            if (lane == 31) {
#pragma unroll
                for (int v = 0; v < K; ++v) wsum[wid * K + v] = incl[v];
            }
            sync_chain();
            if (tid == 0) s_found[(cnt + 1) & 1] = 0;
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
            int vs = -1;
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
                const bool has = je > jb && my_incl > excl && my_loc > (Acc)0;
                if (has && target >= excl && target < my_incl) {
                    Acc acc = excl;
                    int sel = -1, lastpos = jb;
                    for (int j = jb; j < je; ++j) {
                        const Acc wv = (Acc)wbuf[(vs * q + (j - jb)) * B + tid];
                        if (wv > (Acc)0) lastpos = j;
                        acc += wv;
                        if (acc > target) { sel = j; break; }
                    }
                    s_sel = sel >= 0 ? sel : lastpos;
                    s_selv = vs;
                    s_found[cnt & 1] = 1;
                }
            }
            sync_chain();
            if (!s_found[cnt & 1]) {  // block-uniform: no owner (ffbs.cuh) -> the last entry with mass takes the draw
                int key = 0;
#pragma unroll
                for (int v = 0; v < K; ++v) key = (je > jb && loc[v] > (Acc)0) ? v * B + tid + 1 : key;
                if (key) atomicMax(&s_flag, key);
                sync_chain();
                if (key && key == s_flag) {
                    const int vf = (key - 1) / B;
                    int lastpos = jb;
                    for (int j = jb; j < je; ++j)
                        if ((Acc)wbuf[(vf * q + (j - jb)) * B + tid] > (Acc)0) lastpos = j;
                    s_sel = lastpos;
                    s_selv = vf;
                }
                sync_chain();
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
        sync_chain();
        if (degenerate) {
            if (tid == 0) {
                a.status[c] = SMJP_ST_DEGENERATE;
                a.path_len[c] = 0;
            }
            continue;
        }
        // ---- path compaction (smjp_utils.py:319-330), end point (:338-348), metrics (mcmc_utils.py:40-56) ----
        for (int p = tid; p < cnt; p += B) {
            a.path_v[w0 + p] = soj[cnt - 1 - p];
            a.path_t[w0 + p] = Wg[soj[(ncap + 1) + cnt - 1 - p]];
        }
        const int v_last = soj[0];
        const double t_last = Wg[soj[ncap + 1]];
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
                        const double t0 = Wg[soj[(ncap + 1) + cnt - 1 - p]];
                        const double t1 = (p + 1 < cnt) ? Wg[soj[(ncap + 1) + cnt - 2 - p]] : a.t_end;
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
