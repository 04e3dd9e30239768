// Tensor-core forward pass of the dense K-state filter (shape == 1 collapse; cfg4): chains as M.
//
//     r_i = e_i o (r_{i-1} B) / sum(r_{i-1})          (dense_smjp.cuh; hidden_markov_model.py:213-313 in plain-HMM form)
//
// B is the same K x K matrix at every grid point and for every chain, so 128 chains at the same step index form one
// [128 x 64] x [64 x 64] product: a CTA owns a tile of 128 chains (thread t = chain t = TMEM lane t) and runs the
// recursion for all of them in lock step:
//   * B^T (split into TF32 hi + lo parts) sits in shared memory in the canonical K-major no-swizzle UMMA layout for the
//     whole kernel; the message rows r_{i-1} (hi + lo) are written by the threads straight into TENSOR MEMORY
//     (tcgen05.st) and consumed from there as the A operand (tcgen05.mma ... [d], [a], b_desc: "TS" form), so the
//     operand never takes the shared-memory round trip and needs no swizzled staging;
//   * one elected thread issues 8 k-steps x 3 products (hi hi, hi lo, lo hi: 3xTF32 keeps the 1e-4 tolerance of the
//     fp32 path with fp32 accumulation) into a 64-column accumulator in TMEM and commits to an mbarrier;
//   * while the tensor pipe works, every thread prepares its chain's NEXT emission row (64 exp2 of the rate table
//     times the interval length, the observations of the interval); then tcgen05.ld brings the thread's accumulator
//     row to registers, the emission / previous normaliser are applied, the row is summed (the reference's
//     normaliser z_i, thread-local: a thread holds a whole row), stored to HBM for the backward pass and written
//     back to TMEM as the next A operand.
// The backward pass (dense_smjp.cuh, warp per chain) then runs on the stored rows (DenseArgs::fwd_done).
// 192 of 512 TMEM columns per CTA (allocated as 256): two CTAs per SM, one multiplies while the other is in its
// epilogue.  fp32 only, K <= 64 (padded to 64).  Pays when there are many tiles: at cfg4's 1024 chains there are 8
// (one per SM, the other 140 SMs idle) and the warp-per-chain kernel is as fast; the launcher switches on C.
#pragma once
#include "dense_smjp.cuh"

namespace smjp {

constexpr int TC_M = 128, TC_N = 64, TC_K = 64;
constexpr int TC_TMEM_COLS = 256;       // D [0,64) | A_hi [64,128) | A_lo [128,192)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_N >> 3) << 17) | ((TC_M >> 4) << 24);  // f32 += tf32 x tf32, K-major A and B

struct DenseTcSmem {
    float Bhi[TC_N * TC_K];   // canonical K-major, no swizzle: [n / 8][k / 4][n % 8][k % 4]
    float Blo[TC_N * TC_K];
    float rate[TC_N];         // Omega A_v (0 for padded states)
    float rd[TC_N];           // -log2(e) (Omega A_v - min rate)
    unsigned long long mbar;
    uint32_t tmem_base;
    int next_tile;
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: start address, leading (k-chunk) byte offset 128, stride (8-row group) byte offset
// 2048, descriptor version 1 (Blackwell), no swizzle
__device__ __forceinline__ uint64_t tc_b_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(2048 >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}

#define TC_R16(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]), \
                     "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])
#define TC_W16(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7]), \
                     "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]), "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), "r"(v[o + 15])

// 16 consecutive 32-bit columns of this thread's TMEM lane (warp-collective)
template <int NV>
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[NV], const int o) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]),
                   "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]),
                   "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])
                 : "r"(taddr)
                 : "memory");
}
template <int NV>
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[NV], const int o) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7]),
                 "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]), "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]),
                 "r"(v[o + 15])
                 : "memory");
}
#undef TC_R16
#undef TC_W16

__device__ __forceinline__ void tc_mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
    } while (!done);
}

// a.msg rows, a.logZ, a.status (0 / failure bits) and a.path_len (0 for failed chains) are written; tiles of 128 chains
// in the order given (longest grids first keeps the chains of a tile alike).  TWO threads per chain: thread
// (row, half) owns the states [32 half, 32 half + 32) of chain `row` -- 64-wide rows in one thread cost 254 registers
// and left 8 warps per SM (ncu: 15 % issue-active, every stall a latency); halves make it 16.
constexpr int TC_THREADS = 2 * TC_M;
constexpr int TC_H = TC_N / 2;
__global__ void __launch_bounds__(TC_THREADS, 2) dense_tc_forward_kernel(const DenseArgs a, const int32_t* order, int* tile_counter) {
    extern __shared__ __align__(1024) unsigned char tc_raw[];
    DenseTcSmem& S = *reinterpret_cast<DenseTcSmem*>(tc_raw);
    // emission table transposed, rows padded to 65 floats: threads of a warp read DIFFERENT rows (their chains' symbols)
    // at the same column, and (y * 65 + v) mod 32 spreads those over the banks (a stride of 64 put all 32 in one bank)
    constexpr int EMS = TC_N + 1;
    float* emT = reinterpret_cast<float*>(tc_raw + sizeof(DenseTcSmem));  // [Kobs][65]
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (TC_M - 1), half = tid >> 7, vb = half * TC_H;  // chain of the tile, first owned state
    const int K = a.K, Kobs = a.Kobs;
    const double* scale = a.model + 2 * K * K + K * Kobs;
    const double* emis = a.model + 2 * K * K;

    // ---- B = (1 - 1/Omega) I + diag(1 / (Omega A_v)) A, split into TF32 hi / lo, as the B operand: [n = v_to][k = v_from]
    __shared__ double Av_d[TC_N];
    __shared__ float zs[2][TC_M];
    __shared__ int s_nmax;
    for (int v = tid; v < TC_N; v += TC_THREADS) {
        double s = 0;
        if (v < K)
            for (int t = 0; t < K; ++t) s += 1.0 / scale[v * K + t];
        Av_d[v] = s;
    }
    __syncthreads();
    double rmin_d = 1e300;
    for (int v = 0; v < K; ++v) rmin_d = fmin(rmin_d, a.omega * Av_d[v]);
    for (int v = tid; v < TC_N; v += TC_THREADS) {
        S.rate[v] = v < K ? (float)(a.omega * Av_d[v]) : 0.f;
        S.rd[v] = v < K ? (float)(-(double)SMJP_LOG2E * (a.omega * Av_d[v] - rmin_d)) : 0.f;
    }
    for (int t = tid; t < TC_N * TC_K; t += TC_THREADS) {
        const int n = t / TC_K, k = t - n * TC_K;  // n = target state, k = source state
        double b = 0;
        if (n < K && k < K) b = (n == k ? 1.0 - 1.0 / a.omega : 0.0) + (1.0 / scale[k * K + n]) / (a.omega * Av_d[k]);
        const float bf = (float)b;
        const float hi = __uint_as_float(__float_as_uint(bf) & 0xffffe000u);  // what the tensor core reads of a tf32 operand
        const int idx = (n >> 3) * 512 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
        S.Bhi[idx] = hi;
        S.Blo[idx] = bf - hi;
    }
    for (int t = tid; t < Kobs * TC_N; t += TC_THREADS) {
        const int y = t / TC_N, v = t - y * TC_N;
        emT[y * EMS + v] = v < K ? (float)emis[v * Kobs + y] : 0.f;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&S.mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&S.tmem_base)), "n"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // B tiles written by the generic proxy, read by the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    // a warp reaches the TMEM lanes of its quarter (warp % 4); the two halves of a row sit in warps w and w + 4
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)vb;
    const uint32_t mbar = tc_smem_u32(&S.mbar);
    const uint64_t bhi0 = tc_b_desc(tc_smem_u32(S.Bhi)), blo0 = tc_b_desc(tc_smem_u32(S.Blo));
    uint32_t parity = 0;
    const float power = (float)a.power;
    const int n_tiles = (a.C + TC_M - 1) / TC_M;

    while (true) {
        __syncthreads();
        if (tid == 0) {
            S.next_tile = atomicAdd(tile_counter, 1);
            s_nmax = 0;
        }
        __syncthreads();
        const int tile = S.next_tile;
        if (tile >= n_tiles) break;
        const int slot = tile * TC_M + row;
        const int c = slot < a.C ? (order ? order[slot] : slot) : -1;
        int n = 0, D = 0, i_final = 0, flag = 0, dlo = 0;
        int64_t w0 = 0;
        const double* W = nullptr;
        const double* ot = nullptr;
        const int32_t* oy = nullptr;
        if (c >= 0) {
            w0 = a.w_offs[c];
            n = (int)(a.w_offs[c + 1] - w0) - 1;
            W = a.W + w0;
            const int64_t o0 = a.obs_offs[c];
            D = (int)(a.obs_offs[c + 1] - o0);
            ot = a.obs_t + o0;
            oy = a.obs_y + o0;
            if (n < 1) {
                flag = SMJP_ST_GRID_NOT_INCREASING;
                n = 0;
            } else {
                i_final = n;
                while (i_final > 0 && np_isclose(a.t_end, W[i_final])) --i_final;
            }
        }
        float* mh = c >= 0 ? (float*)a.msg + w0 * K : nullptr;
        atomicMax(&s_nmax, n);  // longest grid of the tile
        __syncthreads();
        const int nmax = s_nmax;
        float iz = 1.f;
        double wc = c >= 0 && n > 0 ? W[0] : 0.0;
        // one step ahead of the recursion: the right end of the next interval and the next unconsumed observation
        // (the loads of step i + 1 are in flight while step i computes)
        double wn_next = c >= 0 && n > 0 ? W[1] : 0.0;
        double t_obs = D > 0 ? ot[0] : (double)INFINITY;  // ot[dlo]
        for (int i = 0; i < nmax; ++i) {
            const bool act = i < n && !flag;
            // ---- this thread's half of the emission row of step i (independent of the product in flight) ----
            uint32_t e[TC_H];
            double dwd = 0;
            if (act) {
                const double wn = wn_next;
                wn_next = i + 2 <= n ? W[i + 2] : 0.0;
                dwd = wn - wc;
                if (!(dwd > 0.0)) flag = SMJP_ST_GRID_NOT_INCREASING;
                const float dw = (float)dwd;
                // observations of the closed interval [wc, wn] (smjp_utils.py:40-48): dlo = first one at or after wc
                while (t_obs < wc) {
                    ++dlo;
                    t_obs = dlo < D ? ot[dlo] : (double)INFINITY;
                }
                int cnt = 0;
                if (t_obs <= wn) {
                    cnt = 1;
                    while (dlo + cnt < D && ot[dlo + cnt] <= wn) ++cnt;
                }
                const bool nonfinal = i < i_final;
                const float head = (i == 0 ? (float)(1.0 / K) : iz);
#pragma unroll
                for (int j = 0; j < TC_H; ++j) {
                    const float rt = nonfinal ? S.rate[vb + j] : (vb + j < K ? 1.f : 0.f);
                    e[j] = __float_as_uint(rt * head * Math<float>::exp2(S.rd[vb + j] * dw));
                }
                if (cnt > 0 && power != 0.f) {
                    float f[TC_H];
#pragma unroll
                    for (int j = 0; j < TC_H; ++j) f[j] = a.obs_product ? 1.f : 0.f;
                    for (int d = 0; d < cnt; ++d) {
                        const float* pe = emT + oy[dlo + d] * EMS + vb;
#pragma unroll
                        for (int j = 0; j < TC_H; ++j) f[j] = a.obs_product ? f[j] * pe[j] : f[j] + pe[j];
                    }
                    if (power == 1.f) {
#pragma unroll
                        for (int j = 0; j < TC_H; ++j) e[j] = __float_as_uint(__uint_as_float(e[j]) * f[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < TC_H; ++j) e[j] = __float_as_uint(__uint_as_float(e[j]) * Math<float>::pow(f[j], power));
                    }
                }
                wc = wn;
            } else {
#pragma unroll
                for (int j = 0; j < TC_H; ++j) e[j] = 0u;
            }
            // ---- r_i = e_i o (r_{i-1} B): accumulator half row from tensor memory ----
            if (i > 0) {
                tc_mbar_wait(mbar, parity);
                parity ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t dacc[TC_H];
                tc_ld16(lane_base + 0, dacc, 0);
                tc_ld16(lane_base + 16, dacc, 16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < TC_H; ++j) e[j] = __float_as_uint(__uint_as_float(e[j]) * __uint_as_float(dacc[j]));
            }
            float zp = 0.f;
#pragma unroll
            for (int j = 0; j < TC_H; ++j) zp += __uint_as_float(e[j]);
            zs[half][row] = zp;
            const bool live = act && !flag;
            if (live) {  // the row goes to HBM for the backward pass (a bad row is caught below and never read)
                if (K == TC_N) {
                    float4* dst = reinterpret_cast<float4*>(mh + (int64_t)i * K + vb);
#pragma unroll
                    for (int q = 0; q < TC_H / 4; ++q)
                        dst[q] = make_float4(__uint_as_float(e[4 * q]), __uint_as_float(e[4 * q + 1]), __uint_as_float(e[4 * q + 2]),
                                             __uint_as_float(e[4 * q + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < TC_H; ++j)
                        if (vb + j < K) mh[(int64_t)i * K + vb + j] = __uint_as_float(e[j]);
                }
            }
            if (i + 1 < nmax) {
                // ---- next A operand: the half row, split into tf32 hi + lo, into tensor memory ----
                uint32_t lo[TC_H];
#pragma unroll
                for (int j = 0; j < TC_H; ++j) {
                    const float x = live ? __uint_as_float(e[j]) : 0.f;
                    const uint32_t hi = __float_as_uint(x) & 0xffffe000u;
                    lo[j] = __float_as_uint(x - __uint_as_float(hi));
                    e[j] = hi;
                }
                tc_st16(lane_base + 64, e, 0);
                tc_st16(lane_base + 80, e, 16);
                tc_st16(lane_base + 128, lo, 0);
                tc_st16(lane_base + 144, lo, 16);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (i + 1 < nmax && tid == 0) {  // ---- the products of step i + 1 ----
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int kk = 0; kk < TC_K / 8; ++kk) {
                    const uint64_t bh = bhi0 + (uint64_t)((kk * 256) >> 4), bl = blo0 + (uint64_t)((kk * 256) >> 4);
                    tc_mma_ts(tmem, tmem + 64 + kk * 8, bh, kk > 0);   // hi x hi
                    tc_mma_ts(tmem, tmem + 64 + kk * 8, bl, 1u);       // hi x lo
                    tc_mma_ts(tmem, tmem + 128 + kk * 8, bh, 1u);      // lo x hi
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
            }
            // ---- the row sum z_i (the reference's normaliser): both halves ----
            if (live) {
                const float z = zs[0][row] + zs[1][row];
                if (!(z > 0.f) || !(z < INFINITY)) {
                    flag = SMJP_ST_DEGENERATE;
                } else {
                    iz = 1.f / z;
                    if (a.logZ && half == 0) a.logZ[w0 + i] = log((double)z) - rmin_d * dwd;
                }
            }
            __syncthreads();  // zs is rewritten in the next step
        }
        if (c >= 0 && half == 0) {
            a.status[c] = flag;  // the backward pass skips failed chains
            if (flag) a.path_len[c] = 0;
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS));
}

}  // namespace smjp
