// Dense S-state HMM forward filter + backward sampler, one warp per chain.
//
// Replaces (reference):
//   likelihood_and_decoding_hmm   pkg/hidden_markov_model.py:213-313
//       alpha_t[s] = sum_s' alpha_{t-1}[s'] pi[s',s] e_t[s]   (:256-299), alpha_0 = pi0 * e_0 (:232-238)
//   backward_sampling_hmm         pkg/hidden_markov_model.py:88-202
//       s_T ~ alpha_T (:106-107);  s_t ~ alpha_hat_t * pi[:, s_{t+1}] (:140-151, :186)
// and is the K-state form the sMJP filter collapses to when every Weibull shape is 1
// (B = (1 - 1/Omega) I + diag(1/(Omega A_v)) A).
//
// The reference keeps unnormalised log messages; here rows are normalised (alpha_hat) and the log
// scale factors logc[t] are returned, log alpha_t = log alpha_hat_t + sum_{u<=t} logc[u], so long
// chains do not underflow.  Lane s owns states s, s+32, ...; the matvec broadcasts alpha_{t-1}[s']
// by shuffle and reads row s' of P from shared memory (consecutive lanes, consecutive columns).
#pragma once
#include "../../include/smjp_b200.h"
#include "common.cuh"

namespace smjp {

struct HmmArgs {
    int C, S;
    const int64_t* t_offs;   // [C+1] step windows
    const double* P;         // [S*S] (p_steps == 0) or [p_steps][S*S] (P[t] maps step t-1 -> t)
    int p_steps;
    const double* E;         // dense likelihoods [total_steps][S], or null
    const double* emis;      // table [S*Kobs] with symbols y, when E is null
    const int32_t* y;
    int Kobs;
    const double* pi0;       // [S]
    const double* uniforms;  // per step, or null -> Philox
    uint64_t seed;
    uint32_t sweep, chain_base;
    void* alpha_hat;         // [total_steps][S] Real (required: the backward pass reads it)
    double* logc;            // [total_steps]
    double* loglik;          // [C]
    int32_t* states;         // [total_steps] sampled path, or null (forward only)
    int32_t* status;         // [C]
};

template <typename Real, int SPL>  // SPL = states per lane = ceil(S / 32)
__global__ void hmm_dense_kernel(const HmmArgs a) {
    extern __shared__ __align__(16) unsigned char hmm_smem[];
    Real* Ps = (Real*)hmm_smem;  // [S][S] current transition matrix (shared by the CTA's warps)
    const int S = a.S, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const bool shared_P = a.p_steps == 0;
    if (shared_P) {
        for (int t = threadIdx.x; t < S * S; t += blockDim.x) Ps[t] = (Real)a.P[t];
        __syncthreads();
    }
    for (int c = blockIdx.x * NW + wid; c < a.C; c += gridDim.x * NW) {
        const int64_t t0 = a.t_offs[c];
        const int n = (int)(a.t_offs[c + 1] - t0);
        if (n < 1) {
            if (lane == 0) a.status[c] = SMJP_ST_GRID_NOT_INCREASING;
            continue;
        }
        Real* ah = (Real*)a.alpha_hat + t0 * S;
        Real al[SPL];
        double ll = 0;
        bool bad = false;
        for (int t = 0; t < n; ++t) {
            Real e[SPL];
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                const int s = lane + 32 * k;
                e[k] = 0;
                if (s < S) e[k] = a.E ? (Real)a.E[(t0 + t) * S + s] : (Real)a.emis[s * a.Kobs + a.y[t0 + t]];
            }
            Real nw[SPL];
            if (t == 0) {
#pragma unroll
                for (int k = 0; k < SPL; ++k) {
                    const int s = lane + 32 * k;
                    nw[k] = s < S ? (Real)a.pi0[s] * e[k] : (Real)0;
                }
            } else {
#pragma unroll
                for (int k = 0; k < SPL; ++k) nw[k] = 0;
                const double* Pg = shared_P ? nullptr : a.P + (size_t)min(t, a.p_steps - 1) * S * S;
                for (int sp = 0; sp < S; ++sp) {
                    const Real av = __shfl_sync(0xffffffffu, al[sp >> 5], sp & 31);
#pragma unroll
                    for (int k = 0; k < SPL; ++k) {
                        const int s = lane + 32 * k;
                        if (s < S) nw[k] += av * (shared_P ? Ps[sp * S + s] : (Real)Pg[sp * S + s]);
                    }
                }
#pragma unroll
                for (int k = 0; k < SPL; ++k) nw[k] *= e[k];
            }
            Real z = 0;
#pragma unroll
            for (int k = 0; k < SPL; ++k) z += nw[k];
            z = warp_sum(z);
            if (!(z > (Real)0) || !(z < (Real)INFINITY)) {
                bad = true;
                break;
            }
            const Real iz = (Real)1 / z;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                al[k] = nw[k] * iz;
                const int s = lane + 32 * k;
                if (s < S) ah[(int64_t)t * S + s] = al[k];
            }
            const double lz = log((double)z);
            ll += lz;
            if (lane == 0 && a.logc) a.logc[t0 + t] = lz;
        }
        if (bad) {  // hidden_markov_model.py:170-185 exits the process here
            if (lane == 0) {
                a.status[c] = SMJP_ST_DEGENERATE;
                if (a.loglik) a.loglik[c] = -INFINITY;
            }
            continue;
        }
        if (lane == 0) {
            if (a.loglik) a.loglik[c] = ll;
            a.status[c] = 0;
        }
        if (!a.states) continue;
        __syncwarp();
        // ---- backward sampling: inverse CDF in state order ----
        int s_next = -1;
        for (int t = n - 1; t >= 0; --t) {
            double w[SPL];
            const double* Pg = shared_P ? nullptr : a.P + (size_t)min(t + 1, a.p_steps - 1) * S * S;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                const int s = lane + 32 * k;
                w[k] = 0;
                if (s < S) {
                    const double av = (double)ah[(int64_t)t * S + s];
                    w[k] = s_next < 0 ? av : av * (shared_P ? (double)Ps[s * S + s_next] : Pg[s * S + s_next]);
                }
            }
            // state order is s = lane + 32k: scan block k over lanes, carry block totals
            double incl[SPL], carry = 0;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                incl[k] = warp_inclusive_scan(w[k], lane) + carry;
                carry = __shfl_sync(0xffffffffu, incl[k], 31);
            }
            const double total = carry;
            const double u = a.uniforms ? a.uniforms[t0 + t]
                                        : philox_u01(a.seed, TAG_BACKWARD, (uint64_t)t, 0u, a.sweep,
                                                     a.chain_base + (uint32_t)c);
            const double target = u * total;
            int pick = 0x7fffffff, lastpos = -1;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                const int s = lane + 32 * k;
                if (s < S && w[k] > 0) {
                    lastpos = s;
                    if (incl[k] > target && s < pick) pick = s;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                pick = min(pick, __shfl_xor_sync(0xffffffffu, pick, o));
                lastpos = max(lastpos, __shfl_xor_sync(0xffffffffu, lastpos, o));
            }
            if (pick == 0x7fffffff) pick = lastpos;
            if (pick < 0) {  // all-zero sampling vector
                if (lane == 0) a.status[c] = SMJP_ST_DEGENERATE;
                break;
            }
            if (lane == 0) a.states[t0 + t] = pick;
            s_next = pick;
        }
    }
}

}  // namespace smjp
