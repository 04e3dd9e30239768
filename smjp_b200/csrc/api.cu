// C ABI of the engine (include/smjp_b200.h): context, workspace, model upload, launchers.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/smjp_b200.h"
#include "candidates.cuh"
#include "ffbs_launch.cuh"
#include "dense_smjp.cuh"
#include "dense_tc.cuh"
#include "ffbs_big.cuh"
#include "ffbs_warp.cuh"
#include "hmm_dense.cuh"
#include "params.cuh"
#include "pf.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(x)                                                                        \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess)                                                             \
            return fail(SMJP_E_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                               \
    } while (0)

// growable device / pinned-host buffers owned by the context
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return SMJP_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;  // headroom: grid sizes drift from sweep to sweep
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(SMJP_E_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        }
        cap = want;
        return SMJP_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return SMJP_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(SMJP_E_NOMEM, "cudaMallocHost of %zu bytes failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return SMJP_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct smjp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    size_t smem_optin = 0;
    int64_t launches = 0;
    // model
    int K = 0, Kobs = 0;
    double omega = 0, power = 1, t_end = 0;
    std::vector<double> h_model;
    DevBuf d_model;
    // workspaces
    DevBuf d_queue, d_msg_scratch, d_obsf, d_order, d_dense_msg, d_big, d_failed, d_live;
    int f32_rescue = 1;  // chains that fail in fp32 are redone by the fp64 kernel
    int window_log2 = 0;  // 0: exact live window; -L: states below 2^-L of their row's sum leave the window
    int ffbs_warp = -1;   // one warp per chain, live window in a shared-memory ring (K <= 3, scratch mode): -1 auto, 0 off, 1 on
    int ring_boost[2] = {0, 0};      // extra ring slots learnt from rescue counts, per precision (f64, f32)
    int warp_off[2] = {0, 0};        // the ring kernel kept overflowing for this precision: block kernel from now on
    int ffbs_ring = 0;    // ring slots of 32 grid points (0 = auto)
    int ffbs_fast = -1;   // ffbs_fast.cuh (K <= 3, scratch mode): -1 auto (exact fp64), 0 off, 1 on (both precisions)
    int fast_ring = 0;    // its ring slots of `threads` grid points (0 = auto: ~384 grid points)
    int fast_ways = 0;    // grid pairs in flight per thread (0 = auto; 1, 2, 4)
    int dense_tc = -1;    // tensor-core forward of the dense K-state path (dense_tc.cuh): -1 auto (fp32, 32 <= K <= 64, >= 32768 chains), 0, 1
    DevBuf d_tile, d_pfws;
    int last_dense_tc = 0;
    int64_t dense_nmax_hint = 0;
    int obs_zero_copy = -1;  // smjp_sweep_host: FFBS reads pinned observations in place (-1 auto, 0 off, 1 on)
    int fast_split = -1;  // K <= 3 sweep kernel as forward-only + backward-only launches (-1 auto, 0, 1)
    int last_split = 0;
    DevBuf d_fwdok;
    int cand_R = 1;       // candidate kernels: item slots per sojourn (most target classes of a row, smjp_set_model)
    int fast_cpb = 0;     // chains per CTA, one warp each (0 = auto; 1..11) -- 32-thread groups only
    int fast_boost[2] = {0, 0};  // extra ring slots learnt from rescue counts, per precision
    int fast_off[2] = {0, 0};    // kept overflowing: ffbs.cuh from now on
    bool lpt_order = true;
    int obs_product = 0;
    int pf_extend = 0;
    int hazard_family = SMJP_HAZARD_WEIBULL;
    int pf_cluster = 0, last_pf_cluster = 0, pf_threads = 0;
    int item_k3 = 0;
    int last_grid = 0, last_threads = 0, last_kernel = 0;
    int64_t ffbs_ncap_hint = 0;
    int64_t last_smem = 0;
    DevBuf d_in, d_out;  // staging for the host API
    // host API overlap: a second stream carries the observation upload and the grid download of a sweep
    // while the candidate kernels / the FFBS kernel run on the main stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_obs = nullptr, ev_fill = nullptr;
    bool wait_obs = false;        // the next FFBS launch must wait for ev_obs
    bool host_pending = false;    // smjp_sweep_host_begin returned with work in flight
    bool obs_in_place = false;    // last smjp_sweep_host: the kernels read the caller's pinned observations (stat)
    bool paths_gathered = false;  // last stage_paths: a kernel pulled the paths' used entries out of pinned memory (stat)
    double* w_copy_dst = nullptr;  // pinned host destination of the grid, copied right after cand_fill
    PinBuf h_in, h_out;
    int ffbs_threads = 0;  // 0 = auto
    uint32_t chain_base = 0;
    DevBuf d_stats, d_pi0, d_wlen, d_soff;
    PinBuf h_stats;
    // speculative sweeps (smjp_sweep_dev): the FFBS kernel of a sweep is enqueued for a predicted grid size before
    // this sweep's counts have reached the host, so that the host is never between the GPU and its next kernel
    int speculate = 1;
    int64_t pred_nmax = 0;          // geometry the next sweep is launched for (0: no prediction yet)
    int pred_C = 0, pred_prec = -1;
    cudaEvent_t ev_stats = nullptr;  // {n_max, total} of the current sweep are in h_stats
    const int64_t* guard = nullptr;  // while a speculative sweep is being enqueued: device {n_max, total} + limits
    int64_t guard_nmax = 0, guard_total = 0;
    int64_t spec_launches = 0, spec_misses = 0;
    // rescue counts of ring-kernel launches come home asynchronously, two launches in flight
    PinBuf h_resc;
    cudaEvent_t ev_resc[2] = {nullptr, nullptr};
    int resc_prec[2] = {-1, -1}, resc_C[2] = {0, 0}, resc_fast[2] = {0, 0};
    int resc_idx = 0;
    // optional per-kernel timing (CUDA events on the launching stream)
    bool time_kernels = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ffbs_events, hmm_events, pf_events;
};

using namespace smjp;

namespace {
int weibull_only(smjp_ctx* c, const char* what) {
    if (c->hazard_family != SMJP_HAZARD_WEIBULL)
        return fail(SMJP_E_UNSUPPORTED, "%s is implemented for Weibull hazards only (hazard_family = gamma is a "
                    "particle-filter option)", what);
    return SMJP_OK;
}
}  // namespace

extern "C" {

int smjp_abi_version(void) { return SMJP_ABI_VERSION; }
const char* smjp_last_error(void) { return g_err.c_str(); }

smjp_ctx_t* smjp_ctx_create(int device, void* stream) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fail(SMJP_E_CUDA, "no CUDA device: the smjp_b200 engine has no CPU path");
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        fail(SMJP_E_INVALID, "device %d out of range [0,%d)", device, ndev);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        fail(SMJP_E_CUDA, "cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    smjp_ctx* c = new smjp_ctx();
    c->device = device;
    c->stream = (cudaStream_t)stream;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->num_sms = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    if (prop.major < 10) {
        fail(SMJP_E_UNSUPPORTED, "device sm_%d%d: this library is built for sm_100a only", prop.major, prop.minor);
        delete c;
        return nullptr;
    }
    return c;
}

void smjp_ctx_destroy(smjp_ctx_t* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->d_model.release();
    c->d_queue.release();
    c->d_msg_scratch.release();
    c->d_obsf.release();
    c->d_order.release();
    c->d_dense_msg.release();
    c->d_big.release();
    c->d_failed.release();
    c->d_fwdok.release();
    c->d_live.release();
    c->d_tile.release();
    c->d_pfws.release();
    c->d_in.release();
    c->d_out.release();
    c->h_in.release();
    c->h_out.release();
    c->d_stats.release();
    c->d_pi0.release();
    c->d_wlen.release();
    c->d_soff.release();
    c->h_stats.release();
    if (c->ev_obs) cudaEventDestroy(c->ev_obs);
    if (c->ev_fill) cudaEventDestroy(c->ev_fill);
    if (c->ev_stats) cudaEventDestroy(c->ev_stats);
    for (int i = 0; i < 2; ++i)
        if (c->ev_resc[i]) cudaEventDestroy(c->ev_resc[i]);
    c->h_resc.release();
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (auto* evs : {&c->ffbs_events, &c->hmm_events, &c->pf_events})
        for (auto& ev : *evs) {
            cudaEventDestroy(ev.first);
            cudaEventDestroy(ev.second);
        }
    delete c;
}

int smjp_ctx_sync(smjp_ctx_t* c) {
    if (!c) return fail(SMJP_E_INVALID, "null context");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SMJP_OK;
}

int64_t smjp_ctx_launch_count(smjp_ctx_t* c) { return c ? c->launches : 0; }

int64_t smjp_ctx_get_stat(smjp_ctx_t* c, const char* key) {
    if (!c || !key) return -1;
    if (!strcmp(key, "ffbs_grid")) return c->last_grid;
    if (!strcmp(key, "ffbs_threads")) return c->last_threads;
    if (!strcmp(key, "ffbs_smem")) return c->last_smem;
    if (!strcmp(key, "num_sms")) return c->num_sms;
    if (!strcmp(key, "pf_cluster")) return c->last_pf_cluster;
    if (!strcmp(key, "ffbs_ring_boost_f32")) return c->ring_boost[1] + 100 * c->warp_off[1];
    if (!strcmp(key, "ffbs_ring_boost_f64")) return c->ring_boost[0] + 100 * c->warp_off[0];
    if (!strcmp(key, "ffbs_fast_boost_f32")) return c->fast_boost[1] + 100 * c->fast_off[1];
    if (!strcmp(key, "ffbs_fast_boost_f64")) return c->fast_boost[0] + 100 * c->fast_off[0];
    if (!strcmp(key, "dense_tc")) return c->last_dense_tc;
    if (!strcmp(key, "fast_split")) return c->last_split;
    if (!strcmp(key, "obs_in_place")) return c->obs_in_place ? 1 : 0;
    if (!strcmp(key, "paths_gathered")) return c->paths_gathered ? 1 : 0;
    if (!strcmp(key, "ffbs_kernel")) return c->last_kernel;  // 0 ffbs.cuh, 1 ffbs_item, 2 ffbs_big, 3 ffbs_warp, 4 ffbs_fast, 5 ffbs_fastk
    if (!strcmp(key, "spec_launches")) return c->spec_launches;
    if (!strcmp(key, "spec_misses")) return c->spec_misses;
    if (!strcmp(key, "ffbs_rescued")) {  // chains the last launch handed to the rescue pass (synchronises)
        if (!c->d_failed.p) return 0;
        int32_t v = 0;
        cudaSetDevice(c->device);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
        if (cudaMemcpy(&v, c->d_failed.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return v;
    }
    if (!strcmp(key, "ffbs_live_pairs")) {  // grid pairs processed by the FFBS kernels since the last query
        if (!c->d_live.p) return 0;
        unsigned long long v = 0;
        cudaSetDevice(c->device);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
        if (cudaMemcpy(&v, c->d_live.p, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        cudaMemsetAsync(c->d_live.p, 0, 8, c->stream);
        return (int64_t)v;
    }
    return -1;
}

int smjp_ctx_kernel_time(smjp_ctx_t* c, const char* kernel, double* total_ms, int64_t* count) {
    if (!c || !kernel || !total_ms || !count) return fail(SMJP_E_INVALID, "null argument");
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* evs = nullptr;
    if (!strcmp(kernel, "ffbs")) evs = &c->ffbs_events;
    else if (!strcmp(kernel, "hmm")) evs = &c->hmm_events;
    else if (!strcmp(kernel, "pf")) evs = &c->pf_events;
    else return fail(SMJP_E_INVALID, "unknown kernel '%s'", kernel);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    double ms = 0;
    for (auto& ev : *evs) {
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, ev.first, ev.second));
        ms += t;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    *total_ms = ms;
    *count = (int64_t)evs->size();
    evs->clear();
    return SMJP_OK;
}

int smjp_ctx_set_option(smjp_ctx_t* c, const char* key, int64_t value) {
    if (!c || !key) return fail(SMJP_E_INVALID, "null argument");
    if (!strcmp(key, "ffbs_threads")) {
        if (value != 0 && (value < 32 || value > 512 || value % 32))
            return fail(SMJP_E_INVALID, "ffbs_threads must be 0 (auto) or a multiple of 32 in [32,512]");
        c->ffbs_threads = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "item_k3")) {
        c->item_k3 = value != 0;
        return SMJP_OK;
    }
    if (!strcmp(key, "ffbs_warp")) {
        if (value < -1 || value > 1) return fail(SMJP_E_INVALID, "ffbs_warp must be -1 (auto), 0 or 1");
        c->ffbs_warp = (int)value;
        c->warp_off[0] = c->warp_off[1] = 0;
        return SMJP_OK;
    }
    if (!strcmp(key, "ffbs_ring")) {
        if (value < 0 || value > 64) return fail(SMJP_E_INVALID, "ffbs_ring must be in [0,64]");
        c->ffbs_ring = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "ffbs_fast")) {
        if (value < -1 || value > 1) return fail(SMJP_E_INVALID, "ffbs_fast must be -1 (auto), 0 or 1");
        c->ffbs_fast = (int)value;
        c->fast_off[0] = c->fast_off[1] = 0;
        return SMJP_OK;
    }
    if (!strcmp(key, "fast_ways")) {
        if (value < 0 || value > 4) return fail(SMJP_E_INVALID, "fast_ways must be 0 (auto), 1, 2, 3 or 4");
        c->fast_ways = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "dense_tc")) {
        if (value < -1 || value > 1) return fail(SMJP_E_INVALID, "dense_tc must be -1 (auto), 0 or 1");
        c->dense_tc = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "fast_split")) {
        if (value < -1 || value > 1) return fail(SMJP_E_INVALID, "fast_split must be -1 (auto), 0 or 1");
        c->fast_split = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "obs_zero_copy")) {
        if (value < -1 || value > 1) return fail(SMJP_E_INVALID, "obs_zero_copy must be -1 (auto), 0 or 1");
        c->obs_zero_copy = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "fast_cpb")) {
        if (value < 0 || value > 11) return fail(SMJP_E_INVALID, "fast_cpb must be in [0,11]");
        c->fast_cpb = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "fast_ring")) {
        if (value < 0 || value > 64) return fail(SMJP_E_INVALID, "fast_ring must be in [0,64]");
        c->fast_ring = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "window_log2")) {
        if (value > 0 || value < -1000) return fail(SMJP_E_INVALID, "window_log2 must be 0 (exact) or in [-1000,-1]");
        c->window_log2 = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "speculate")) {  // 2 (tests): predict 3/4 of the previous size, so that most launches are repeated
        c->speculate = value == 2 ? 2 : (value ? 1 : 0);
        c->pred_nmax = 0;
        return SMJP_OK;
    }
    if (!strcmp(key, "f32_rescue")) {
        c->f32_rescue = value ? 1 : 0;
        return SMJP_OK;
    }
    if (!strcmp(key, "pf_threads")) {
        if (value != 0 && value != 256 && value != 512 && value != 1024)
            return fail(SMJP_E_INVALID, "pf_threads must be 0 (auto), 256, 512 or 1024");
        c->pf_threads = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "pf_cluster")) {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8 && value != 16)
            return fail(SMJP_E_INVALID, "pf_cluster must be 0 (auto), 1, 2, 4, 8 or 16");
        c->pf_cluster = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "hazard_family")) {
        if (value != SMJP_HAZARD_WEIBULL && value != SMJP_HAZARD_GAMMA) return fail(SMJP_E_INVALID, "bad hazard family");
        c->hazard_family = (int)value;
        return SMJP_OK;
    }
    if (!strcmp(key, "pf_extend")) {
        c->pf_extend = value != 0;
        return SMJP_OK;
    }
    if (!strcmp(key, "obs_product")) {
        c->obs_product = value != 0;
        return SMJP_OK;
    }
    if (!strcmp(key, "lpt_order")) {
        c->lpt_order = value != 0;
        return SMJP_OK;
    }
    if (!strcmp(key, "time_kernels")) {
        c->time_kernels = value != 0;
        return SMJP_OK;
    }
    if (!strcmp(key, "chain_base")) {
        if (value < 0 || value > 0xFFFFFFFFll) return fail(SMJP_E_INVALID, "chain_base out of range");
        c->chain_base = (uint32_t)value;
        return SMJP_OK;
    }
    return fail(SMJP_E_INVALID, "unknown option '%s'", key);
}

int smjp_set_model(smjp_ctx_t* c, int K, int Kobs, const double* shape, const double* scale,
                   double omega, const double* emis, double likelihood_power, double t_end) {
    if (!c) return fail(SMJP_E_INVALID, "null context");
    if (K < 2 || K > 64) return fail(SMJP_E_INVALID, "K=%d outside [2,64]", K);
    if (Kobs < 1 || Kobs > 4096) return fail(SMJP_E_INVALID, "Kobs=%d outside [1,4096]", Kobs);
    if (!(omega > 1.0)) return fail(SMJP_E_INVALID, "omega must be > 1 (thinning needs B = omega*A > A)");
    if (!(t_end > 0.0)) return fail(SMJP_E_INVALID, "t_end must be positive");
    if (!shape || !scale || !emis) return fail(SMJP_E_INVALID, "null parameter array");
    for (int t = 0; t < K * K; ++t)
        if (!(shape[t] > 0.0) || !(scale[t] > 0.0) || !std::isfinite(shape[t]) || !std::isfinite(scale[t]))
            return fail(SMJP_E_INVALID, "shape/scale entry %d is not positive finite", t);
    // the fp64 kernels evaluate 2^x with a 32-bit integer part of 256 x (Exp2<double>), so |x| must stay below
    // 2^23 for every positive double tau (|log2 tau| <= 1075): x = (k-1) log2 tau - (k log2 lambda - log2 k)
    // for a hazard, k log2(tau/lambda) and log2 lambda + log2(e)/k in the particle filter's holding times
    for (int t = 0; t < K * K; ++t) {
        const double k = shape[t], l2 = std::fabs(std::log2(scale[t]));
        if ((k + 1.0) * 1075.0 + k * l2 + std::fabs(std::log2(k)) >= 8388608.0 || 1075.0 / k + l2 >= 8388608.0)
            return fail(SMJP_E_INVALID, "shape/scale entry %d: exponents leave the +-2^23 range of the device arithmetic", t);
    }
    for (int t = 0; t < K * Kobs; ++t)
        if (!(emis[t] >= 0.0) || !std::isfinite(emis[t]))
            return fail(SMJP_E_INVALID, "emission entry %d is negative or not finite", t);
    CUDA_TRY(cudaSetDevice(c->device));
    if (K != c->K) c->ffbs_ncap_hint = 0;
    c->pred_nmax = 0;  // grid sizes follow the model: no speculative launch for the next sweep
    for (int p = 0; p < 2; ++p) {  // ring sizes learnt from rescue counts belong to the previous model's windows
        c->ring_boost[p] = c->warp_off[p] = 0;
        c->fast_boost[p] = c->fast_off[p] = 0;
    }
    c->K = K;
    c->Kobs = Kobs;
    c->omega = omega;
    c->power = likelihood_power;
    c->t_end = t_end;
    // device layout: shape[K*K], c2[K*K] = k*log2(lambda), emis[K*Kobs], scale[K*K], then for the candidate
    // sampler crep[K*K] and csum[K*K]: targets of one row that share a shape form a class (thinned events of
    // equal-shape Weibull hazards superpose into one Poisson process with the same position density);
    // crep = first target of the class, csum = sum over the class of lambda^-k (at the representative), replist[K*K] =
    // the class representatives of every row in order (-1 padded)
    c->h_model.assign((size_t)6 * K * K + (size_t)K * Kobs, 0.0);
    for (int t = 0; t < K * K; ++t) {
        c->h_model[t] = shape[t];
        c->h_model[K * K + t] = shape[t] * std::log2(scale[t]);
        c->h_model[2 * K * K + K * Kobs + t] = scale[t];
    }
    {
        double* crep = c->h_model.data() + 3 * K * K + (size_t)K * Kobs;
        double* csum = crep + K * K;
        for (int v = 0; v < K; ++v)
            for (int s = 0; s < K; ++s) {
                int r = s;
                for (int u = 0; u < s; ++u)
                    if (shape[v * K + u] == shape[v * K + s]) { r = u; break; }
                crep[v * K + s] = (double)r;
                csum[v * K + r] += std::pow(scale[v * K + s], -shape[v * K + s]);
            }
        // replist[v][r] = the r-th class representative of row v (-1 beyond the row's classes); cand_R = most classes
        // of any row: the candidate kernels give every sojourn cand_R item slots
        double* replist = csum + K * K;
        c->cand_R = 1;
        for (int v = 0; v < K; ++v) {
            int nr = 0;
            for (int s = 0; s < K; ++s)
                if ((int)crep[v * K + s] == s) replist[v * K + nr++] = (double)s;
            for (int r = nr; r < K; ++r) replist[v * K + r] = -1.0;
            if (nr > c->cand_R) c->cand_R = nr;
        }
    }
    for (int t = 0; t < K * Kobs; ++t) c->h_model[2 * K * K + t] = emis[t];
    int rc = c->d_model.reserve(c->h_model.size() * sizeof(double));
    if (rc) return rc;
    // stream-ordered with respect to kernels already enqueued that read the previous model
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(c->d_model.p, c->h_model.data(), c->h_model.size() * sizeof(double),
                        cudaMemcpyHostToDevice));
    return SMJP_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// FFBS launcher
// ------------------------------------------------------------------------------------------------
namespace {

template <typename Real>
cudaError_t ffbs_call(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    if (sizeof(Real) == 8) return L.K <= 5 ? ffbs_dispatch_f64_lo(a, L, query) : ffbs_dispatch_f64_hi(a, L, query);
    return L.K <= 5 ? ffbs_dispatch_f32_lo(a, L, query) : ffbs_dispatch_f32_hi(a, L, query);
}

// run-time-K kernel (ffbs_big.cuh): 128 threads when the chain state fits next to the per-thread columns, else 64
template <typename Real>
cudaError_t ffbs_big_call(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    auto k128 = ffbs_big_kernel<Real, 128>;
    auto k64 = ffbs_big_kernel<Real, 64>;
    auto kern = L.threads == 128 ? k128 : k64;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return e;
    if (query) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L.grid, kern, L.threads, L.smem);
    if (L.threads == 128) k128<<<L.grid, 128, L.smem, L.stream>>>(a);
    else k64<<<L.grid, 64, L.smem, L.stream>>>(a);
    return cudaGetLastError();
}

// warp-per-chain ring kernel (ffbs_warp.cuh), K = 2, 3
template <typename Real, int K>
cudaError_t ffbs_warp_do(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    auto kern = ffbs_warp_kernel<Real, K>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return e;
    if (query) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L.grid, kern, FFBS_WARPS * 32, L.smem);
    kern<<<L.grid, FFBS_WARPS * 32, L.smem, L.stream>>>(a);
    return cudaGetLastError();
}
// K >= 4 was tried in fp32 (K = 10: 228 registers, 128 with spills; 28-32 ms per cfg5 sweep against 15.5 ms for
// the item-parallel kernel): the ring kernel stays with K <= 3
constexpr int ffbs_warp_max_k(size_t) { return 3; }
template <typename Real>
cudaError_t ffbs_warp_call(const FfbsArgs& a, FfbsLaunch& L, bool query) {
    switch (L.K) {
        case 2: return ffbs_warp_do<Real, 2>(a, L, query);
        case 3: return ffbs_warp_do<Real, 3>(a, L, query);
    }
    return cudaErrorInvalidValue;
}

template <typename Real>
int ffbs_dev_impl(smjp_ctx* c, FfbsArgs& a, int64_t n_max) {
    const int K = c->K;
    const bool big = K > 10;  // general shapes, run-time K
    const bool item_kernel = !big && (K >= 4 || (K == 3 && c->item_k3));
    a.K = K;
    if (K <= 3) {  // hazard parameters of the thread-per-grid-point kernel travel as kernel arguments
        for (int t = 0; t < K * K; ++t) {
            const double k = c->h_model[t], c2p = c->h_model[K * K + t] - std::log2(k);
            a.pc_d[0][t] = 1.0 / k;
            a.pc_d[1][t] = (k - 1.0) * 256.0;  // Exp2<double>::SCALE
            a.pc_d[2][t] = c2p * 256.0;
            a.pc_f[0][t] = (float)(1.0 / k);
            a.pc_f[1][t] = (float)(k - 1.0);
            a.pc_f[2][t] = (float)c2p;
        }
    }
    a.use_item = item_kernel;  // thread per (v, j) for larger state spaces, thread per j for K <= 3
    int threads;
    size_t smem;
    int big_slots = 0;
    if (big) {
        // the chain state (2 K reals per live grid point) sits in a ring of slots of `threads` grid points: as many
        // slots as shared memory holds (never more than the whole grid needs).  Exported messages need them all.
        const size_t budget = c->smem_optin - 6 * 1024;  // the kernel's static shared memory (2^x / log2 tables)
        auto slots_that_fit = [&](int th) {
            const size_t fixed = ffbs_big_smem_bytes<Real>(K, 0, th), per = 2 * smjp::align_up((size_t)K * th * sizeof(Real), 16);
            const int full = ffbs_nslot((int)n_max, th);
            const int fit = fixed >= budget ? 0 : (int)((budget - fixed) / per);
            return fit < full ? fit : full;
        };
        threads = 128;
        big_slots = slots_that_fit(128);
        // 64-thread slots waste less of the ring (the window is covered in finer steps) and leave room for more of them
        if (big_slots < ffbs_nslot((int)n_max, 128) && slots_that_fit(64) * 64 > big_slots * 128) {
            threads = 64;
            big_slots = slots_that_fit(64);
        }
        if (big_slots < 1 || (a.messages && big_slots < ffbs_nslot((int)n_max, threads)))
            return fail(SMJP_E_UNSUPPORTED,
                        "K=%d, grid of %lld intervals: %s does not fit in shared memory", K, (long long)n_max,
                        a.messages ? "the whole chain state (exported messages need every grid point)" : "one ring slot of chain state");
        smem = ffbs_big_smem_bytes<Real>(K, big_slots, threads);
        // hazard parameters in the kernel's precision
        const double S = sizeof(Real) == 8 ? 256.0 : 1.0;  // Exp2<Real>::SCALE
        std::vector<Real> hp((size_t)3 * K * K);
        for (int t = 0; t < K * K; ++t) {
            const double k = c->h_model[t];
            hp[t] = (Real)(1.0 / k);
            hp[K * K + t] = (Real)((k - 1.0) * S);
            hp[2 * K * K + t] = (Real)((c->h_model[K * K + t] - std::log2(k)) * S);
        }
        int rcb = c->d_big.reserve(hp.size() * sizeof(Real));
        if (rcb) return rcb;
        CUDA_TRY(cudaMemcpyAsync(c->d_big.p, hp.data(), hp.size() * sizeof(Real), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));  // hp is a stack-lifetime host buffer
        a.big_params = c->d_big.p;
    } else if (item_kernel) {
        threads = ffbs_item_threads(K, n_max);
        if (c->ffbs_threads == ffbs_item_base_threads(K) || c->ffbs_threads == 2 * ffbs_item_base_threads(K))
            threads = c->ffbs_threads;
        smem = ffbs_item_smem_bytes<Real>(K, (int)n_max, threads);
        if (smem > c->smem_optin && threads == ffbs_item_base_threads(K)) {
            threads *= 2;  // more threads never need more shared memory per slot; retry for the message
            smem = ffbs_item_smem_bytes<Real>(K, (int)n_max, threads);
        }
    } else {
        threads = c->ffbs_threads;
        if (threads == 0) {
            // a full row should give every thread a few grid points; small grids use small CTAs
            // fp32: the live window (~100 grid points on cfg2) keeps two warps busy, and two warps reduce and
            // synchronise faster than four (4.6 vs 5.2 ms on cfg2); fp64's window is ~250 points: four warps
            threads = sizeof(Real) == 4 ? 64 : 128;
            if (n_max > 1024) threads = 256;
            if (n_max > 3072) threads = 512;
            if (n_max <= 96) threads = 64;
            if (n_max <= 32) threads = 32;
        }
        threads = ffbs_round_threads(threads);
        smem = ffbs_smem_bytes<Real>(K, (int)n_max, threads);
    }
    // warp-per-chain ring kernel (option ffbs_warp): K <= 3, scratch mode
    const int pidx = sizeof(Real) == 4 ? 1 : 0;
    {
        // an earlier ring-kernel launch: how many of its chains outgrew the ring and went to the rescue pass?
        // (a few are harmless; many mean this workload's windows are wider than the ring: widen it, or give up).
        // The count of launch i travels home behind it and is read when launch i+2 is prepared.
        const int slot = c->resc_idx & 1;
        if (c->resc_prec[slot] >= 0) {
            CUDA_TRY(cudaEventSynchronize(c->ev_resc[slot]));
            const int32_t resc = ((const int32_t*)c->h_resc.p)[slot];
            const int pp = c->resc_prec[slot];
            if ((int64_t)resc * 50 > c->resc_C[slot]) {
                if (c->resc_fast[slot]) {
                    c->fast_boost[pp] += 1;
                    if (c->fast_boost[pp] > 4) c->fast_off[pp] = 1;
                } else {
                    c->ring_boost[pp] += 2;
                    if (c->ring_boost[pp] > 8) c->warp_off[pp] = 1;
                }
            }
            c->resc_prec[slot] = -1;
        }
    }
    // auto: fp32 (window ~100-150 grid points on cfg2) and pruned windows; the exact fp64 window (~250-300 points)
    // needs a ring that leaves 12 warps per SM and gains 5 % only: block kernel
    const bool want_warp = c->ffbs_warp == 1 || (c->ffbs_warp == -1 && (sizeof(Real) == 4 || c->window_log2 < 0));
    bool use_warp = want_warp && !c->warp_off[pidx] && K <= ffbs_warp_max_k(sizeof(Real)) && !(K == 3 && c->item_k3) &&
                    !a.messages && n_max <= 65535 && c->f32_rescue;
    int ring = 0;
    size_t warp_smem = 0;
    if (use_warp) {
        ring = c->ffbs_ring > 0 ? c->ffbs_ring : (c->window_log2 < 0 || K > 3 ? 4 : (sizeof(Real) == 4 ? 6 : 10)) + c->ring_boost[pidx];
        const int full = (int)(n_max / 32) + 1;  // slots that hold a whole grid: the ring never needs more
        if (ring > full) ring = full;
        warp_smem = ffbs_warp_smem_bytes<Real>(K, ring, (int)n_max);
        if (warp_smem > c->smem_optin) use_warp = false;
    }
    // ffbs_fast.cuh (K <= 3, scratch mode): deferred normalisation, split barriers, state ring.  Auto: the exact fp64
    // window (fp32 and pruned windows are narrower and run best one warp per chain, above)
    const bool want_fast = c->ffbs_fast == 1 || (c->ffbs_fast == -1 && sizeof(Real) == 8 && c->window_log2 == 0);
    // ffbs_fastk.cuh is the same design for 4 <= K <= 10 (lane = (grid point, source state)), both precisions
    const bool want_fastk = c->ffbs_fast == 1 || (c->ffbs_fast == -1 && c->window_log2 == 0);
    const bool fastk = item_kernel && K >= 4 && K <= 10;
    bool use_fast = (fastk ? want_fastk : want_fast && !item_kernel && K <= 3) && !use_warp && !c->fast_off[pidx] && !big &&
                    !a.messages && c->f32_rescue;
    int fast_threads = 0, fast_ns = 0, fast_cpb = 1, fast_ways = 1;
    size_t fast_smem = 0;
    if (use_fast) {
        // default geometry: ONE WARP PER CHAIN, two grid pairs in flight per lane, several chains per CTA (they share
        // the 2^x / log2 tables) -- measured on cfg2: 128 threads per chain 14.1 ms, 64: 11.8, 32: 11.3 (7 chains per
        // SM, shared memory bound), 32 with the grid and the row starts moved to L2 and 5 chains per CTA: see DESIGN
        fast_threads = fastk ? 32 : c->ffbs_threads ? ffbs_fast_round_threads(c->ffbs_threads) : 32;
        fast_ways = c->fast_ways ? c->fast_ways : (fast_threads <= 64 ? 2 : 1);
        fast_cpb = fast_threads == 32 ? (c->fast_cpb ? c->fast_cpb : 1) : 1;
        if (!ffbs_fast_has_variant(fast_threads, fast_ways, fast_cpb)) {
            fast_ways = fast_threads <= 64 ? 2 : 1;
            fast_cpb = 1;
        }
        if (fastk) {
            // ring of slots of JPS = 32 / K grid points: the exact window of a 10-state chain at T = 20 is <= ~90 grid
            // points in fp64 (~40 in fp32); + RS steps between judgements + slack; learnt upwards from rescue counts
            const int jps = 32 / K;
            const int pts = (sizeof(Real) == 8 ? 104 : 96) + 16 * c->fast_boost[pidx];
            const int full = (int)(n_max / jps) + 1;
            fast_cpb = 1;
            fast_ways = c->fast_ways == 1 ? 1 : 2;
            fast_ns = c->fast_ring > 0 ? c->fast_ring : (pts + jps - 1) / jps;
            if (fast_ns > full) fast_ns = full;
            fast_smem = (size_t)ffbs_fastk_layout<Real>(K, fast_ns, (int)n_max).total;
        } else {
            const int full = ffbs_nslot((int)n_max, fast_threads);
            // ring: the exact fp64 window of cfg2 is <= ~250 grid points; + RS steps between judgements + one slot of slack
            fast_ns = c->fast_ring > 0 ? c->fast_ring : (288 + fast_threads - 1) / fast_threads + 1 + c->fast_boost[pidx];
            if (fast_ns > full) fast_ns = full;
            fast_smem = (size_t)ffbs_fast_layout<Real>(K, fast_ns, (int)n_max, fast_threads).total * fast_cpb;
        }
        if (fast_smem > c->smem_optin) use_fast = false;
    }
    // rescue pass: chains that fail in fp32, or whose live window outgrows the ring, are redone by the full-size
    // fp64 kernel (scratch mode, K <= 10): geometry of that launch
    constexpr int RESCUE_GRID = 32;
    bool rescue = (sizeof(Real) == 4 || use_warp || use_fast) && c->f32_rescue && !a.messages && !big;
    int rescue_threads = 0;
    size_t rescue_smem = 0;
    if (rescue) {
        if (item_kernel) {
            rescue_threads = threads;
            rescue_smem = ffbs_item_smem_bytes<double>(K, (int)n_max, rescue_threads);
            if (rescue_smem > c->smem_optin && rescue_threads == ffbs_item_base_threads(K)) {
                rescue_threads *= 2;
                rescue_smem = ffbs_item_smem_bytes<double>(K, (int)n_max, rescue_threads);
            }
        } else {
            // the rescued chains run alone at the end of the launch: give each the CTA size that is fastest per chain
            rescue_threads = (threads < 128 && n_max > 96) ? 128 : threads;
            rescue_smem = ffbs_smem_bytes<double>(K, (int)n_max, rescue_threads);
        }
        if (rescue_smem > c->smem_optin) rescue = false;
    }
    // the ring kernel hands chains whose window outgrows the ring to the rescue pass: without one (the full-size fp64
    // kernel does not fit) it must not run -- the block kernel of this precision handles every chain itself
    if (use_warp && !rescue) use_warp = false;
    if (use_fast && !rescue) use_fast = false;
    if (use_fast) {
        threads = fast_threads;
        smem = fast_smem;
    }
    a.use_fast = use_fast;
    a.fast_ways = fast_ways;
    a.fast_cpb = fast_cpb;
    if (smem > c->smem_optin)
        return fail(SMJP_E_UNSUPPORTED,
                    "grid with %lld intervals x K=%d needs %zu B of shared memory (> %zu): "
                    "chain state does not fit on chip", (long long)n_max, K, smem, c->smem_optin);
    FfbsLaunch L;
    L.K = K;
    L.threads = use_warp ? FFBS_WARPS * 32 : threads;  // (fast kernel: threads per CHAIN; the CTA has threads * fast_cpb)
    L.smem = use_warp ? warp_smem : smem;
    L.grid = 0;
    L.stream = c->stream;
    CUDA_TRY(use_warp ? ffbs_warp_call<Real>(a, L, true) : big ? ffbs_big_call<Real>(a, L, true) : ffbs_call<Real>(a, L, true));
    const int per_sm = L.grid;
    if (per_sm < 1) return fail(SMJP_E_UNSUPPORTED, "kernel cannot be resident (threads=%d smem=%zu)", L.threads, L.smem);
    int grid = c->num_sms * per_sm;
    const int regions_per_cta = use_warp ? FFBS_WARPS : use_fast ? fast_cpb : 1;  // scratch regions: one per chain in flight
    if (grid > (a.C + regions_per_cta - 1) / regions_per_cta) grid = (a.C + regions_per_cta - 1) / regions_per_cta;
    // K <= 3 sweep kernel as two launches (ffbs_fast.cuh, PHASE 1 / 2): needs the rows of EVERY chain in scratch
    bool split = use_fast && !fastk && threads == 32 && fast_cpb == 1 && fast_ways == 2 && c->fast_split != 0 && !a.messages;
    a.ncap = (int)n_max;
    a.nslot = big ? big_slots : use_warp ? ring : use_fast ? fast_ns : item_kernel ? ffbs_item_nslot(K, (int)n_max, threads) : ffbs_nslot((int)n_max, threads);
    if (use_fast && fastk) a.lay = ffbs_fastk_layout<Real>(K, fast_ns, (int)n_max);
    else if (!item_kernel && !big) a.lay = use_fast ? ffbs_fast_layout<Real>(K, fast_ns, (int)n_max, threads) : ffbs_layout<Real>(K, (int)n_max, threads);
    a.window_eps = c->window_log2 < 0 ? std::ldexp(1.0, c->window_log2) : 0.0;
    int rc;
    if (!a.messages) {
        // grid lengths drift from sweep to sweep: size the per-CTA scratch for a rounded-up, never
        // shrinking length so that steady-state sweeps do not reallocate
        int64_t hint = (n_max + n_max / 8 + 63) / 64 * 64;
        if (hint < c->ffbs_ncap_hint) hint = c->ffbs_ncap_hint;
        c->ffbs_ncap_hint = hint;
        const int64_t stride = (int64_t)K * hint * (hint + 1) / 2;
        size_t bytes = (size_t)stride * sizeof(Real) * (size_t)grid * regions_per_cta;
        if (split) {  // a region per chain: only while that stays a modest share of the device (cfg2: 14 GB)
            const size_t all = (size_t)stride * sizeof(Real) * (size_t)a.C;
            if (all > c->d_msg_scratch.cap) {
                size_t free_b = 0, total_b = 0;
                CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
                if (all > (free_b + c->d_msg_scratch.cap) / 3) split = false;
            }
            if (split) bytes = all;
        }
        if (rescue && bytes < (size_t)stride * 8 * RESCUE_GRID) bytes = (size_t)stride * 8 * RESCUE_GRID;
        if (bytes > c->d_msg_scratch.cap) {  // (steady state: no growth, and no trip into the driver for the query)
            size_t free_b = 0, total_b = 0;
            CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
            size_t budget = free_b + c->d_msg_scratch.cap;
            budget -= budget / 10;
            if (bytes > budget) {  // fewer resident CTAs rather than failing
                int g2 = (int)(budget / ((size_t)stride * sizeof(Real) * regions_per_cta));
                if (g2 < 1)
                    return fail(SMJP_E_NOMEM, "message scratch for one chain (%zu B) does not fit", (size_t)stride * sizeof(Real));
                grid = g2;
                bytes = (size_t)stride * sizeof(Real) * (size_t)grid * regions_per_cta;
            }
        }
        rc = c->d_msg_scratch.reserve(bytes);
        if (rc) return rc;
        a.msg_scratch = c->d_msg_scratch.p;
        a.msg_scratch_stride = stride;
    }
    {
        // per chain in flight: K observation factors per interval; the warp kernel reuses its region for the
        // sojourn records (12 bytes per grid point)
        int64_t per = (int64_t)n_max * K;
        if (use_warp) {
            per = (int64_t)(n_max + 2) * K;
            const int64_t rec = (12 * (n_max + 2) + (int64_t)sizeof(Real) - 1) / (int64_t)sizeof(Real);
            if (per < rec) per = rec;
            per = (per + 3) / 4 * 4;  // keeps every region 16-byte aligned
            a.obsf_stride = per;
        } else if (use_fast) {  // + the window start of every row (ints), see ffbs_fast.cuh
            per += ffbs_fast_jlo_elems<Real>(n_max);
            if (fastk) per += 3 * K * K;  // + the hazard tables the backward pass reads (ffbs_fastk.cuh)
            if (split) per += (ffbs_fast_split_ints(n_max) * 4 + (int64_t)sizeof(Real) - 1) / (int64_t)sizeof(Real);  // + sojourn list
            per = (per + 3) / 4 * 4;
            a.obsf_stride = per;
        } else if (big) {  // + window starts [n + 1] and the sojourn list [2][n + 1] (ints), see ffbs_big.cuh
            per += ((int64_t)3 * (n_max + 2) * 4 + (int64_t)sizeof(Real) - 1) / (int64_t)sizeof(Real);
            per = (per + 3) / 4 * 4;
            a.obsf_stride = per;
        }
        size_t ob = (size_t)(split ? a.C : grid * regions_per_cta) * per * sizeof(Real);
        if (rescue && ob < (size_t)RESCUE_GRID * n_max * K * 8) ob = (size_t)RESCUE_GRID * n_max * K * 8;
        rc = c->d_obsf.reserve(ob);
    }
    if (rc) return rc;
    a.obsf_scratch = c->d_obsf.p;
    rc = c->d_queue.reserve(256);
    if (rc) return rc;
    if (!c->d_live.p) {
        rc = c->d_live.reserve(8);
        if (rc) return rc;
        CUDA_TRY(cudaMemsetAsync(c->d_live.p, 0, 8, c->stream));
    }
    a.live_pairs = (unsigned long long*)c->d_live.p;
    CUDA_TRY(cudaMemsetAsync(c->d_queue.p, 0, 3 * sizeof(unsigned int), c->stream));
    a.queue = (unsigned int*)c->d_queue.p;
    a.queue2 = a.queue + 2;  // (+ 1: the rescue pass)
    if (a.C > grid && c->lpt_order) {  // more chains than resident CTAs: start the long grids first
        rc = c->d_order.reserve((size_t)a.C * sizeof(int32_t));
        if (rc) return rc;
        chain_order_kernel<<<1, 1024, 0, c->stream>>>(a.C, a.w_offs, (int)n_max, (int32_t*)c->d_order.p);
        CUDA_TRY(cudaGetLastError());
        c->launches += 1;
        a.order = (const int32_t*)c->d_order.p;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->time_kernels) {
        CUDA_TRY(cudaEventCreate(&e0));
        CUDA_TRY(cudaEventCreate(&e1));
        CUDA_TRY(cudaEventRecord(e0, c->stream));
    }
    L.grid = grid;
    if (split) {
        rc = c->d_fwdok.reserve((size_t)a.C * sizeof(int32_t));
        if (rc) return rc;
        CUDA_TRY(cudaMemsetAsync(c->d_fwdok.p, 0, (size_t)a.C * sizeof(int32_t), c->stream));
        a.fwd_ok = (int32_t*)c->d_fwdok.p;
        a.fast_phase = 1;
    }
    CUDA_TRY(use_warp ? ffbs_warp_call<Real>(a, L, false) : big ? ffbs_big_call<Real>(a, L, false) : ffbs_call<Real>(a, L, false));
    if (split) {  // the backward pass of every chain: small shared memory, half the registers, twice the warps
        FfbsArgs b = a;
        b.fast_phase = 2;
        b.lay = ffbs_fast_bwd_layout<Real>(K, fast_ns, threads);
        FfbsLaunch B2 = L;
        B2.smem = (size_t)b.lay.total;
        B2.grid = 0;
        CUDA_TRY(ffbs_call<Real>(b, B2, true));
        if (B2.grid < 1) return fail(SMJP_E_UNSUPPORTED, "backward kernel cannot be resident (smem=%zu)", B2.smem);
        B2.grid *= c->num_sms;
        if (B2.grid > a.C) B2.grid = a.C;
        CUDA_TRY(ffbs_call<Real>(b, B2, false));
        c->launches += 1;
        a.fast_phase = 0;
    }
    c->last_split = split ? 1 : 0;
    if (rescue) {
        // fp32 has ~38 orders of magnitude per grid step: a step over which every live state loses more than
        // that (or gains it) leaves a chain without a row sum.  Those chains -- rare, none on most sweeps --
        // are listed on the device and redone by the fp64 kernel on a few CTAs, same scratch, same stream.
        rc = c->d_failed.reserve((size_t)a.C * sizeof(int32_t) + 16);
        if (rc) return rc;
        int32_t* cnt = (int32_t*)c->d_failed.p;
        int32_t* list = cnt + 4;
        collect_failed_kernel<<<1, 1024, 0, c->stream>>>(a.C, a.status, list, cnt);
        CUDA_TRY(cudaGetLastError());
        FfbsArgs r = a;
        r.use_fast = 0;
        r.order = list;
        r.C_dev = cnt;
        r.queue = (unsigned int*)c->d_queue.p + 1;
        FfbsLaunch R;
        R.K = K;
        R.threads = rescue_threads;
        R.smem = rescue_smem;
        R.stream = c->stream;
        R.grid = RESCUE_GRID < a.C ? RESCUE_GRID : a.C;
        r.nslot = item_kernel ? ffbs_item_nslot(K, (int)n_max, rescue_threads) : ffbs_nslot((int)n_max, rescue_threads);
        if (!item_kernel) r.lay = ffbs_layout<double>(K, (int)n_max, rescue_threads);
        CUDA_TRY(ffbs_call<double>(r, R, false));
        c->launches += 2;
        if (use_warp || use_fast) {
            const int slot = c->resc_idx & 1;
            c->resc_fast[slot] = use_fast ? 1 : 0;
            rc = c->h_resc.reserve(2 * sizeof(int32_t));
            if (rc) return rc;
            if (!c->ev_resc[slot]) CUDA_TRY(cudaEventCreateWithFlags(&c->ev_resc[slot], cudaEventDisableTiming));
            CUDA_TRY(cudaMemcpyAsync((int32_t*)c->h_resc.p + slot, cnt, 4, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaEventRecord(c->ev_resc[slot], c->stream));
            c->resc_prec[slot] = pidx;
            c->resc_C[slot] = a.C;
        }
    }
    if (c->time_kernels) {
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        c->ffbs_events.emplace_back(e0, e1);
    }
    c->launches += 1;
    ++c->resc_idx;
    c->last_grid = grid;
    c->last_threads = L.threads;
    c->last_kernel = use_warp ? 3 : use_fast ? (fastk ? 5 : 4) : big ? 2 : item_kernel ? 1 : 0;
    c->last_smem = (int64_t)L.smem;
    return SMJP_OK;
}

}  // namespace

extern "C" {

int smjp_ffbs_dev(smjp_ctx_t* c, int C, const int64_t* w_offs, const double* W,
                  const int64_t* obs_offs, const double* obs_t, const int32_t* obs_y,
                  const double* uniforms, uint64_t seed, uint32_t sweep, int precision,
                  void* messages, const int64_t* msg_offs, double* logZ, int32_t* aug_idx,
                  int32_t* path_v, double* path_t, int32_t* path_len, double* time_in_state,
                  int32_t* jumps, int32_t* status, int64_t n_max, int64_t total_w) {
    if (!c) return fail(SMJP_E_INVALID, "null context");
    if (c->K == 0) return fail(SMJP_E_INVALID, "smjp_set_model has not been called");
    if (C < 0) return fail(SMJP_E_INVALID, "C < 0");
    if (C == 0) return SMJP_OK;
    if (!w_offs || !W || !obs_offs || !path_v || !path_t || !path_len || !status)
        return fail(SMJP_E_INVALID, "null required pointer");
    if (messages && !msg_offs) return fail(SMJP_E_INVALID, "messages given without msg_offs");
    if (n_max < 1) return fail(SMJP_E_INVALID, "n_max < 1: every grid needs at least [0, t_end]");
    if (total_w < (int64_t)C * 2) return fail(SMJP_E_INVALID, "total_w inconsistent with C");
    if (precision != SMJP_F64 && precision != SMJP_F32) return fail(SMJP_E_INVALID, "bad precision");
    {
        const int rcw = weibull_only(c, "the FFBS kernel");
        if (rcw) return rcw;
    }
    CUDA_TRY(cudaSetDevice(c->device));
    FfbsArgs a;
    memset(&a, 0, sizeof a);
    a.C = C;
    a.w_offs = w_offs;
    a.W = W;
    a.obs_offs = obs_offs;
    a.obs_t = obs_t;
    a.obs_y = obs_y;
    a.uniforms = uniforms;
    a.seed = seed;
    a.sweep = sweep;
    a.chain_base = c->chain_base;
    a.messages = messages;
    a.msg_offs = msg_offs;
    a.logZ = logZ;
    a.aug_idx = aug_idx;
    a.path_v = path_v;
    a.path_t = path_t;
    a.path_len = path_len;
    a.time_in_state = time_in_state;
    a.jumps = jumps;
    a.status = status;
    a.order = nullptr;
    a.model = (const double*)c->d_model.p;
    a.Kobs = c->Kobs;
    a.obs_product = c->obs_product;
    a.omega = c->omega;
    a.power = c->power;
    a.t_end = c->t_end;
    a.guard = c->guard;
    a.guard_nmax = c->guard_nmax;
    a.guard_total = c->guard_total;
    return precision == SMJP_F64 ? ffbs_dev_impl<double>(c, a, n_max) : ffbs_dev_impl<float>(c, a, n_max);
}

int smjp_ffbs_host(smjp_ctx_t* c, int C, const int64_t* w_offs, const double* W,
                   const int64_t* obs_offs, const double* obs_t, const int32_t* obs_y,
                   const double* uniforms, uint64_t seed, uint32_t sweep, int precision,
                   void* messages, const int64_t* msg_offs, double* logZ, int32_t* aug_idx,
                   int32_t* path_v, double* path_t, int32_t* path_len, double* time_in_state,
                   int32_t* jumps, int32_t* status) {
    if (!c) return fail(SMJP_E_INVALID, "null context");
    if (c->K == 0) return fail(SMJP_E_INVALID, "smjp_set_model has not been called");
    if (C < 0) return fail(SMJP_E_INVALID, "C < 0");
    if (C == 0) return SMJP_OK;
    if (!w_offs || !W || !obs_offs || !path_v || !path_t || !path_len || !status)
        return fail(SMJP_E_INVALID, "null required pointer");
    if (messages && !msg_offs) return fail(SMJP_E_INVALID, "messages given without msg_offs");
    CUDA_TRY(cudaSetDevice(c->device));
    const int K = c->K;
    const int64_t total_w = w_offs[C], total_o = obs_offs[C];
    if (w_offs[0] != 0 || obs_offs[0] != 0) return fail(SMJP_E_INVALID, "offsets must start at 0");
    int64_t n_max = 0;
    for (int i = 0; i < C; ++i) {
        const int64_t len = w_offs[i + 1] - w_offs[i];
        if (len < 2) return fail(SMJP_E_INVALID, "chain %d: a grid needs at least the two end points", i);
        if (obs_offs[i + 1] < obs_offs[i]) return fail(SMJP_E_INVALID, "obs_offs not monotone at chain %d", i);
        if (len - 1 > n_max) n_max = len - 1;
    }
    if (total_o > 0 && (!obs_t || !obs_y)) return fail(SMJP_E_INVALID, "observations missing");
    for (int64_t d = 0; d < total_o; ++d)  // the kernels index emis[v * Kobs + y] unchecked
        if (obs_y[d] < 0 || obs_y[d] >= c->Kobs) return fail(SMJP_E_INVALID, "observation symbol outside [0,Kobs)");
    const size_t esz = precision == SMJP_F64 ? 8 : 4;
    const int64_t total_msg = messages ? msg_offs[C] : 0;

    // ---- input staging layout (all 16-byte aligned) ----
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += smjp::align_up(bytes, 16); return o; };
    const size_t o_woffs = take((C + 1) * 8), o_W = take(total_w * 8), o_ooffs = take((C + 1) * 8);
    const size_t o_ot = take(total_o * 8), o_oy = take(total_o * 4);
    const size_t o_u = uniforms ? take(total_w * 8) : 0;
    const size_t o_moffs = messages ? take((C + 1) * 8) : 0;
    const size_t in_bytes = off;
    int rc = c->h_in.reserve(in_bytes);
    if (rc) return rc;
    rc = c->d_in.reserve(in_bytes);
    if (rc) return rc;
    char* hin = (char*)c->h_in.p;
    memcpy(hin + o_woffs, w_offs, (C + 1) * 8);
    memcpy(hin + o_W, W, total_w * 8);
    memcpy(hin + o_ooffs, obs_offs, (C + 1) * 8);
    if (total_o) {
        memcpy(hin + o_ot, obs_t, total_o * 8);
        memcpy(hin + o_oy, obs_y, total_o * 4);
    }
    if (uniforms) memcpy(hin + o_u, uniforms, total_w * 8);
    if (messages) memcpy(hin + o_moffs, msg_offs, (C + 1) * 8);
    CUDA_TRY(cudaMemcpyAsync(c->d_in.p, hin, in_bytes, cudaMemcpyHostToDevice, c->stream));

    // ---- output staging layout ----
    off = 0;
    const size_t o_pv = take(total_w * 4), o_pt = take(total_w * 8), o_pl = take(C * 4), o_st = take(C * 4);
    const size_t o_lz = logZ ? take(total_w * 8) : 0;
    const size_t o_ai = aug_idx ? take(total_w * 4) : 0;
    const size_t o_tis = time_in_state ? take((size_t)C * K * 8) : 0;
    const size_t o_jm = jumps ? take((size_t)C * K * 4) : 0;
    const size_t small_bytes = off;
    const size_t o_msg = messages ? take(total_msg * esz) : 0;
    const size_t out_bytes = off;
    rc = c->d_out.reserve(out_bytes);
    if (rc) return rc;
    rc = c->h_out.reserve(small_bytes);
    if (rc) return rc;
    char* din = (char*)c->d_in.p;
    char* dout = (char*)c->d_out.p;
    if (aug_idx) CUDA_TRY(cudaMemsetAsync(dout + o_ai, 0, total_w * 4, c->stream));
    if (logZ) CUDA_TRY(cudaMemsetAsync(dout + o_lz, 0, total_w * 8, c->stream));
    rc = smjp_ffbs_dev(c, C, (const int64_t*)(din + o_woffs), (const double*)(din + o_W),
                       (const int64_t*)(din + o_ooffs), (const double*)(din + o_ot),
                       (const int32_t*)(din + o_oy), uniforms ? (const double*)(din + o_u) : nullptr,
                       seed, sweep, precision, messages ? (void*)(dout + o_msg) : nullptr,
                       messages ? (const int64_t*)(din + o_moffs) : nullptr,
                       logZ ? (double*)(dout + o_lz) : nullptr, aug_idx ? (int32_t*)(dout + o_ai) : nullptr,
                       (int32_t*)(dout + o_pv), (double*)(dout + o_pt), (int32_t*)(dout + o_pl),
                       time_in_state ? (double*)(dout + o_tis) : nullptr,
                       jumps ? (int32_t*)(dout + o_jm) : nullptr, (int32_t*)(dout + o_st), n_max, total_w);
    if (rc) return rc;
    char* hout = (char*)c->h_out.p;
    CUDA_TRY(cudaMemcpyAsync(hout, dout, small_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (messages)  // large: straight into the caller's (pageable) buffer
        CUDA_TRY(cudaMemcpyAsync(messages, dout + o_msg, total_msg * esz, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(path_v, hout + o_pv, total_w * 4);
    memcpy(path_t, hout + o_pt, total_w * 8);
    memcpy(path_len, hout + o_pl, C * 4);
    memcpy(status, hout + o_st, C * 4);
    if (logZ) memcpy(logZ, hout + o_lz, total_w * 8);
    if (aug_idx) memcpy(aug_idx, hout + o_ai, total_w * 4);
    if (time_in_state) memcpy(time_in_state, hout + o_tis, (size_t)C * K * 8);
    if (jumps) memcpy(jumps, hout + o_jm, (size_t)C * K * 4);
    return SMJP_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// prior / observations / candidates / sweep
// ------------------------------------------------------------------------------------------------
namespace {

int model_ready(smjp_ctx* c) {
    if (!c) return fail(SMJP_E_INVALID, "null context");
    if (c->K == 0) return fail(SMJP_E_INVALID, "smjp_set_model has not been called");
    return SMJP_OK;
}


CandArgs make_cand_args(smjp_ctx* c, int C, const int64_t* p_offs, const int32_t* path_v,
                        const double* path_t, const int32_t* path_len, uint64_t seed, uint32_t sweep) {
    CandArgs a;
    memset(&a, 0, sizeof a);
    a.C = C;
    a.p_offs = p_offs;
    a.path_v = path_v;
    a.path_t = path_t;
    a.path_len = path_len;
    a.model = (const double*)c->d_model.p;
    a.K = c->K;
    a.Kobs = c->Kobs;
    a.omega = c->omega;
    a.seed = seed;
    a.sweep = sweep;
    a.chain_base = c->chain_base;
    a.guard = c->guard;
    a.guard_total = c->guard_total;
    a.R = c->cand_R;
    return a;
}

}  // namespace

extern "C" {

int smjp_prior_dev(smjp_ctx_t* c, int C, const int64_t* p_offs, const double* pi0, int honour_scale,
                   uint64_t seed, int32_t* path_v, double* path_t, int32_t* path_len, int32_t* status) {
    int rc = model_ready(c);
    if (rc) return rc;
    rc = weibull_only(c, "the prior sampler");
    if (rc) return rc;
    if (C <= 0) return C == 0 ? SMJP_OK : fail(SMJP_E_INVALID, "C < 0");
    if (!p_offs || !pi0 || !path_v || !path_t || !path_len || !status)
        return fail(SMJP_E_INVALID, "null required pointer");
    double tot = 0;
    for (int s = 0; s < c->K; ++s) {
        if (!(pi0[s] >= 0)) return fail(SMJP_E_INVALID, "pi0[%d] negative", s);
        tot += pi0[s];
    }
    if (!(tot > 0)) return fail(SMJP_E_INVALID, "pi0 has no mass");
    CUDA_TRY(cudaSetDevice(c->device));
    rc = c->d_pi0.reserve(64 * sizeof(double));
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(c->d_pi0.p, pi0, c->K * sizeof(double), cudaMemcpyHostToDevice));
    PriorArgs a;
    memset(&a, 0, sizeof a);
    a.C = C;
    a.p_offs = p_offs;
    a.model = (const double*)c->d_model.p;
    a.pi0 = (const double*)c->d_pi0.p;
    a.K = c->K;
    a.Kobs = c->Kobs;
    a.t_end = c->t_end;
    a.honour_scale = honour_scale;
    a.seed = seed;
    a.chain_base = c->chain_base;
    a.path_v = path_v;
    a.path_t = path_t;
    a.path_len = path_len;
    a.status = status;
    prior_kernel<<<(C + 63) / 64, 64, 0, c->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    c->launches += 1;
    return SMJP_OK;
}

int smjp_simulate_obs_dev(smjp_ctx_t* c, int C, const int64_t* p_offs, const int32_t* path_v,
                          const double* path_t, const int32_t* path_len, const int64_t* obs_offs,
                          const double* obs_t, uint64_t seed, int32_t* obs_y) {
    int rc = model_ready(c);
    if (rc) return rc;
    if (C <= 0) return C == 0 ? SMJP_OK : fail(SMJP_E_INVALID, "C < 0");
    if (!p_offs || !path_v || !path_t || !path_len || !obs_offs || !obs_t || !obs_y)
        return fail(SMJP_E_INVALID, "null required pointer");
    CUDA_TRY(cudaSetDevice(c->device));
    ObsSimArgs a;
    memset(&a, 0, sizeof a);
    a.C = C;
    a.p_offs = p_offs;
    a.path_v = path_v;
    a.path_t = path_t;
    a.path_len = path_len;
    a.obs_offs = obs_offs;
    a.obs_t = obs_t;
    a.obs_y = obs_y;
    a.model = (const double*)c->d_model.p;
    a.K = c->K;
    a.Kobs = c->Kobs;
    a.seed = seed;
    a.chain_base = c->chain_base;
    int grid = C < c->num_sms * 8 ? C : c->num_sms * 8;
    obs_sim_kernel<<<grid, 128, 0, c->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    c->launches += 1;
    return SMJP_OK;
}

}  // extern "C"

namespace {
// count pass of the candidate sampler: per-chain grid lengths, their prefix sums (w_offs) and {n_max, total} on
// their way to c->h_stats; c->ev_stats is recorded behind the copy
int cand_count_enqueue(smjp_ctx* c, int C, const int64_t* p_offs, const int32_t* path_v, const double* path_t,
                       const int32_t* path_len, uint64_t seed, uint32_t sweep, int64_t* w_offs, int32_t* soff) {
    int rc = model_ready(c);
    if (rc) return rc;
    rc = weibull_only(c, "the candidate sampler");
    if (rc) return rc;
    if (C <= 0) return fail(SMJP_E_INVALID, "C <= 0");
    if (!p_offs || !path_v || !path_t || !path_len || !w_offs || !soff)
        return fail(SMJP_E_INVALID, "null required pointer");
    CUDA_TRY(cudaSetDevice(c->device));
    rc = c->d_wlen.reserve((size_t)C * sizeof(int64_t));
    if (rc) return rc;
    rc = c->d_stats.reserve(2 * sizeof(int64_t));
    if (rc) return rc;
    rc = c->h_stats.reserve(2 * sizeof(int64_t));
    if (rc) return rc;
    if (!c->ev_stats) CUDA_TRY(cudaEventCreateWithFlags(&c->ev_stats, cudaEventDisableTiming));
    CandArgs a = make_cand_args(c, C, p_offs, path_v, path_t, path_len, seed, sweep);
    a.soff = soff;
    a.w_len = (int64_t*)c->d_wlen.p;
    int grid = (C + 3) / 4 < c->num_sms * 16 ? (C + 3) / 4 : c->num_sms * 16;  // a warp per chain, 4 per CTA
    cand_count_kernel<<<grid, 128, 0, c->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    chain_scan_kernel<<<1, 1024, 0, c->stream>>>(C, (const int64_t*)c->d_wlen.p, w_offs, (int64_t*)c->d_stats.p);
    CUDA_TRY(cudaGetLastError());
    c->launches += 2;
    CUDA_TRY(cudaMemcpyAsync(c->h_stats.p, c->d_stats.p, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev_stats, c->stream));
    return SMJP_OK;
}

// geometry the NEXT sweep is launched for, from this sweep's longest grid: a few per cent of head room (the
// maximum over thousands of chains moves by ~2 % between sweeps), but not across a 128-point slot boundary of
// the block kernel when the boundary itself leaves room -- a slot more costs a resident CTA per SM
int64_t predict_nmax(int64_t n_max) {
    int64_t pred = n_max + (n_max / 24 > 16 ? n_max / 24 : 16);
    const int64_t edge = (n_max / 128 + 1) * 128 - 1;
    if (pred > edge && edge >= n_max + 8) pred = edge;
    return pred;
}
}  // namespace

extern "C" {

int smjp_candidates_count_dev(smjp_ctx_t* c, int C, const int64_t* p_offs, const int32_t* path_v,
                              const double* path_t, const int32_t* path_len, uint64_t seed,
                              uint32_t sweep, int64_t* w_offs, int32_t* soff, int64_t* n_max_host,
                              int64_t* total_w_host) {
    if (!n_max_host || !total_w_host) return fail(SMJP_E_INVALID, "null required pointer");
    int rc = cand_count_enqueue(c, C, p_offs, path_v, path_t, path_len, seed, sweep, w_offs, soff);
    if (rc) return rc;
    CUDA_TRY(cudaEventSynchronize(c->ev_stats));
    *n_max_host = ((int64_t*)c->h_stats.p)[0];
    *total_w_host = ((int64_t*)c->h_stats.p)[1];
    return SMJP_OK;
}

int smjp_candidates_fill_dev(smjp_ctx_t* c, int C, const int64_t* p_offs, const int32_t* path_v,
                             const double* path_t, const int32_t* path_len, uint64_t seed,
                             uint32_t sweep, int mode, const int64_t* w_offs, const int32_t* soff,
                             double* W) {
    int rc = model_ready(c);
    if (rc) return rc;
    if (C <= 0) return fail(SMJP_E_INVALID, "C <= 0");
    if (mode != SMJP_CAND_REFERENCE && mode != SMJP_CAND_EXACT) return fail(SMJP_E_INVALID, "bad candidate mode");
    if (!p_offs || !path_v || !path_t || !path_len || !w_offs || !soff || !W)
        return fail(SMJP_E_INVALID, "null required pointer");
    CUDA_TRY(cudaSetDevice(c->device));
    CandArgs a = make_cand_args(c, C, p_offs, path_v, path_t, path_len, seed, sweep);
    a.mode = mode;
    a.soff = const_cast<int32_t*>(soff);
    a.w_offs = w_offs;
    a.W = W;
    int grid = (C + 3) / 4 < c->num_sms * 16 ? (C + 3) / 4 : c->num_sms * 16;
    cand_fill_kernel<<<grid, 128, 0, c->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    c->launches += 1;
    return SMJP_OK;
}

int smjp_sweep_dev(smjp_ctx_t* c, int C, const int64_t* p_offs, const int32_t* path_v,
                   const double* path_t, const int32_t* path_len, const int64_t* obs_offs,
                   const double* obs_t, const int32_t* obs_y, uint64_t seed, uint32_t sweep, int mode,
                   int precision, int64_t capacity, int64_t* w_offs, int32_t* soff, double* W,
                   int32_t* path_v_out, double* path_t_out, int32_t* path_len_out,
                   double* time_in_state, int32_t* jumps, int32_t* status, int64_t* n_max_host,
                   int64_t* total_w_host) {
    if (!c) return fail(SMJP_E_INVALID, "null context");
    if (!n_max_host || !total_w_host) return fail(SMJP_E_INVALID, "null required pointer");
    int rc = cand_count_enqueue(c, C, p_offs, path_v, path_t, path_len, seed, sweep, w_offs, soff);
    if (rc) return rc;
    // Speculative launch.  The FFBS geometry (shared memory, scratch, grid) needs this sweep's longest grid, which
    // is known on the device only.  Waiting for it leaves the GPU idle while the host prepares the launch, and
    // exposes every host hiccup.  Instead the fill and FFBS kernels are enqueued NOW for the size predicted from
    // the previous sweep, guarded on the device by the real counts; the host then collects the counts while the
    // kernels run, and only a sweep that outgrew the prediction (its kernels returned at once) is launched again.
    const bool spec = c->speculate && c->pred_nmax > 0 && c->pred_C == C && c->pred_prec == precision;
    const int64_t pred = c->pred_nmax;
    int spec_rc = SMJP_OK;
    size_t events_before = c->ffbs_events.size();
    if (spec) {
        ++c->spec_launches;
        c->guard = (const int64_t*)c->d_stats.p;
        c->guard_nmax = pred;
        c->guard_total = capacity;
        spec_rc = smjp_candidates_fill_dev(c, C, p_offs, path_v, path_t, path_len, seed, sweep, mode, w_offs, soff, W);
        if (!spec_rc) {
            if (c->w_copy_dst) cudaEventRecord(c->ev_fill, c->stream);  // host call: the grid is copied once its size is known
            if (c->wait_obs) {
                cudaStreamWaitEvent(c->stream, c->ev_obs, 0);
                c->wait_obs = false;
            }
            spec_rc = smjp_ffbs_dev(c, C, w_offs, W, obs_offs, obs_t, obs_y, nullptr, seed, sweep, precision, nullptr,
                                    nullptr, nullptr, nullptr, path_v_out, path_t_out, path_len_out, time_in_state,
                                    jumps, status, pred, capacity);
        }
        c->guard = nullptr;
        c->guard_nmax = c->guard_total = 0;
    }
    CUDA_TRY(cudaEventSynchronize(c->ev_stats));
    *n_max_host = ((int64_t*)c->h_stats.p)[0];
    *total_w_host = ((int64_t*)c->h_stats.p)[1];
    if (*total_w_host > capacity) {
        c->pred_nmax = 0;
        return fail(SMJP_E_NOMEM, "sweep needs %lld grid points, buffers hold %lld", (long long)*total_w_host,
                    (long long)capacity);
    }
    c->pred_nmax = c->speculate == 2 ? (*n_max_host * 3 / 4 > 1 ? *n_max_host * 3 / 4 : 1) : predict_nmax(*n_max_host);
    c->pred_C = C;
    c->pred_prec = precision;
    if (spec) {
        if (!spec_rc && *n_max_host <= pred) {
            if (c->w_copy_dst) {  // host call: the grid goes home while the FFBS kernel runs
                CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_fill, 0));
                CUDA_TRY(cudaMemcpyAsync(c->w_copy_dst, W, (size_t)*total_w_host * 8, cudaMemcpyDeviceToHost, c->copy_stream));
            }
            return SMJP_OK;
        }
        // the guarded kernels returned without touching anything (or were never launched): launch for the real size
        ++c->spec_misses;
        while (c->ffbs_events.size() > events_before) {  // the empty launch is not a kernel time
            cudaEventDestroy(c->ffbs_events.back().first);
            cudaEventDestroy(c->ffbs_events.back().second);
            c->ffbs_events.pop_back();
        }
    }
    rc = smjp_candidates_fill_dev(c, C, p_offs, path_v, path_t, path_len, seed, sweep, mode, w_offs, soff, W);
    if (rc) return rc;
    if (c->w_copy_dst) {  // host call: the grid goes home while the FFBS kernel runs
        CUDA_TRY(cudaEventRecord(c->ev_fill, c->stream));
        CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_fill, 0));
        CUDA_TRY(cudaMemcpyAsync(c->w_copy_dst, W, (size_t)*total_w_host * 8, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    if (c->wait_obs) {  // host call: the observations were uploaded on the copy stream
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_obs, 0));
        c->wait_obs = false;
    }
    return smjp_ffbs_dev(c, C, w_offs, W, obs_offs, obs_t, obs_y, nullptr, seed, sweep, precision, nullptr,
                         nullptr, nullptr, nullptr, path_v_out, path_t_out, path_len_out, time_in_state, jumps,
                         status, *n_max_host, *total_w_host);
}

}  // extern "C"

namespace {

// stage the current paths (+ optionally the observations) of a host call into ctx->d_in
struct StagedPaths {
    const int64_t* p_offs;
    const int32_t* path_v;
    const double* path_t;
    const int32_t* path_len;
    const int64_t* obs_offs;
    const double* obs_t;
    const int32_t* obs_y;
};

int stage_paths(smjp_ctx* c, int C, const int64_t* p_offs, const int32_t* path_v, const double* path_t,
                const int32_t* path_len, const int64_t* obs_offs, const double* obs_t, const int32_t* obs_y,
                StagedPaths* out, bool obs_on_copy_stream = false, const double* obs_t_mapped = nullptr,
                const int32_t* obs_y_mapped = nullptr, const int32_t* path_v_mapped = nullptr,
                const double* path_t_mapped = nullptr) {
    if (p_offs[0] != 0) return fail(SMJP_E_INVALID, "offsets must start at 0");
    const int64_t total_p = p_offs[C];
    int64_t used_p = 0;
    for (int i = 0; i < C; ++i) {
        if (p_offs[i + 1] < p_offs[i]) return fail(SMJP_E_INVALID, "p_offs not monotone");
        if (path_len[i] < 2 || path_len[i] > p_offs[i + 1] - p_offs[i])
            return fail(SMJP_E_INVALID, "chain %d: path_len %d outside [2, window]", i, path_len[i]);
        used_p += path_len[i];
    }
    const int64_t total_o = obs_offs ? obs_offs[C] : 0;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += smjp::align_up(bytes, 16); return o; };
    const size_t o_po = take((C + 1) * 8), o_pv = take(total_p * 4), o_pt = take(total_p * 8), o_pl = take(C * 4);
    const size_t o_oo = take((C + 1) * 8), o_ot = take(total_o * 8), o_oy = take(total_o * 4);
    // the caller's arrays go straight to the device: full PCIe rate when they are pinned, the driver's
    // own staging when they are pageable -- no extra host copy here
    int rc = c->d_in.reserve(off);
    if (rc) return rc;
    char* dd = (char*)c->d_in.p;
    if (off <= ((size_t)1 << 18) && !obs_on_copy_stream) {
        // small batch: pack everything into the pinned staging buffer, one host-to-device copy
        rc = c->h_in.reserve(off);
        if (rc) return rc;
        char* h = (char*)c->h_in.p;
        CUDA_TRY(cudaStreamSynchronize(c->stream));  // an earlier copy out of h_in has completed
        memcpy(h + o_po, p_offs, (C + 1) * 8);
        memcpy(h + o_pv, path_v, total_p * 4);
        memcpy(h + o_pt, path_t, total_p * 8);
        memcpy(h + o_pl, path_len, C * 4);
        if (obs_offs) {
            memcpy(h + o_oo, obs_offs, (C + 1) * 8);
            if (total_o) {
                memcpy(h + o_ot, obs_t, total_o * 8);
                memcpy(h + o_oy, obs_y, total_o * 4);
            }
        }
        CUDA_TRY(cudaMemcpyAsync(dd, h, off, cudaMemcpyHostToDevice, c->stream));
        out->p_offs = (const int64_t*)(dd + o_po);
        out->path_v = (const int32_t*)(dd + o_pv);
        out->path_t = (const double*)(dd + o_pt);
        out->path_len = (const int32_t*)(dd + o_pl);
        out->obs_offs = (const int64_t*)(dd + o_oo);
        out->obs_t = (const double*)(dd + o_ot);
        out->obs_y = (const int32_t*)(dd + o_oy);
        return SMJP_OK;
    }
    auto put = [&](size_t o, const void* src, size_t bytes) -> cudaError_t {
        return bytes ? cudaMemcpyAsync(dd + o, src, bytes, cudaMemcpyHostToDevice, c->stream) : cudaSuccess;
    };
    CUDA_TRY(put(o_po, p_offs, (C + 1) * 8));
    CUDA_TRY(put(o_pl, path_len, C * 4));
    if (path_v_mapped && path_t_mapped && used_p * 10 < total_p * 7) {
        // sparse windows in pinned memory (a sweep's output windows are grid-sized, its paths a fraction of that): a
        // kernel pulls only the entries in use over PCIe (cfg2: 5 MB instead of 26 MB)
        int grid = (C + 7) / 8 < c->num_sms * 8 ? (C + 7) / 8 : c->num_sms * 8;
        smjp::path_gather_kernel<<<grid, 256, 0, c->stream>>>(C, (const int64_t*)(dd + o_po), (const int32_t*)(dd + o_pl),
                                                               path_v_mapped, path_t_mapped, (int32_t*)(dd + o_pv),
                                                               (double*)(dd + o_pt));
        CUDA_TRY(cudaGetLastError());
        c->launches += 1;
        c->paths_gathered = true;
    } else {
        CUDA_TRY(put(o_pv, path_v, total_p * 4));
        CUDA_TRY(put(o_pt, path_t, total_p * 8));
        c->paths_gathered = false;
    }
    if (obs_offs) {
        cudaStream_t os = obs_on_copy_stream ? c->copy_stream : c->stream;
        auto puto = [&](size_t o, const void* src, size_t bytes) -> cudaError_t {
            return bytes ? cudaMemcpyAsync(dd + o, src, bytes, cudaMemcpyHostToDevice, os) : cudaSuccess;
        };
        CUDA_TRY(puto(o_oo, obs_offs, (C + 1) * 8));
        if (!obs_t_mapped) {
            CUDA_TRY(puto(o_ot, obs_t, total_o * 8));
            CUDA_TRY(puto(o_oy, obs_y, total_o * 4));
        }
        if (obs_on_copy_stream) {
            CUDA_TRY(cudaEventRecord(c->ev_obs, c->copy_stream));
            c->wait_obs = true;
        }
    }
    char* d = (char*)c->d_in.p;
    out->p_offs = (const int64_t*)(d + o_po);
    out->path_v = (const int32_t*)(d + o_pv);
    out->path_t = (const double*)(d + o_pt);
    out->path_len = (const int32_t*)(d + o_pl);
    out->obs_offs = (const int64_t*)(d + o_oo);
    out->obs_t = obs_t_mapped ? obs_t_mapped : (const double*)(d + o_ot);
    out->obs_y = obs_t_mapped ? obs_y_mapped : (const int32_t*)(d + o_oy);
    return SMJP_OK;
}

}  // namespace

extern "C" {

int smjp_candidates_host(smjp_ctx_t* c, int C, const int64_t* p_offs, const int32_t* path_v,
                         const double* path_t, const int32_t* path_len, uint64_t seed, uint32_t sweep,
                         int mode, int64_t* w_offs, double* W, int64_t w_capacity, int64_t* total_w) {
    int rc = model_ready(c);
    if (rc) return rc;
    if (C <= 0) return fail(SMJP_E_INVALID, "C <= 0");
    if (!p_offs || !path_v || !path_t || !path_len || !w_offs || !total_w)
        return fail(SMJP_E_INVALID, "null required pointer");
    CUDA_TRY(cudaSetDevice(c->device));
    StagedPaths sp;
    rc = stage_paths(c, C, p_offs, path_v, path_t, path_len, nullptr, nullptr, nullptr, &sp);
    if (rc) return rc;
    rc = c->d_soff.reserve((size_t)p_offs[C] * 4 + 16);
    if (rc) return rc;
    rc = c->d_out.reserve((size_t)(C + 1) * 8);
    if (rc) return rc;
    int64_t n_max = 0;
    rc = smjp_candidates_count_dev(c, C, sp.p_offs, sp.path_v, sp.path_t, sp.path_len, seed, sweep,
                                   (int64_t*)c->d_out.p, (int32_t*)c->d_soff.p, &n_max, total_w);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(w_offs, c->d_out.p, (size_t)(C + 1) * 8, cudaMemcpyDeviceToHost));
    if (*total_w > w_capacity || !W)
        return fail(SMJP_E_NOMEM, "grids need %lld doubles, W holds %lld", (long long)*total_w, (long long)w_capacity);
    // d_out: [w_offs | W]
    const size_t o_W = smjp::align_up((size_t)(C + 1) * 8, 16);
    DevBuf tmp;  // keep w_offs alive while d_out may be regrown
    rc = tmp.reserve(o_W + (size_t)*total_w * 8);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(tmp.p, c->d_out.p, (size_t)(C + 1) * 8, cudaMemcpyDeviceToDevice, c->stream));
    rc = smjp_candidates_fill_dev(c, C, sp.p_offs, sp.path_v, sp.path_t, sp.path_len, seed, sweep, mode,
                                  (const int64_t*)tmp.p, (const int32_t*)c->d_soff.p, (double*)((char*)tmp.p + o_W));
    if (rc) { tmp.release(); return rc; }
    cudaError_t e = cudaMemcpyAsync(W, (char*)tmp.p + o_W, (size_t)*total_w * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    tmp.release();
    if (e != cudaSuccess) return fail(SMJP_E_CUDA, "candidate copy-back failed: %s", cudaGetErrorString(e));
    return SMJP_OK;
}

int smjp_sweep_host_end(smjp_ctx_t* c) {
    if (!c) return fail(SMJP_E_INVALID, "null context");
    if (!c->host_pending) return SMJP_OK;
    c->host_pending = false;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SMJP_OK;
}

int smjp_sweep_host(smjp_ctx_t* c, int C, const int64_t* p_offs, const int32_t* path_v,
                    const double* path_t, const int32_t* path_len, const int64_t* obs_offs,
                    const double* obs_t, const int32_t* obs_y, uint64_t seed, uint32_t sweep, int mode,
                    int precision, int64_t capacity, int64_t* w_offs, double* W, int32_t* path_v_out,
                    double* path_t_out, int32_t* path_len_out, double* time_in_state, int32_t* jumps,
                    int32_t* status, int64_t* total_w) {
    int rc = smjp_sweep_host_begin(c, C, p_offs, path_v, path_t, path_len, obs_offs, obs_t, obs_y, seed, sweep, mode,
                                   precision, capacity, w_offs, W, path_v_out, path_t_out, path_len_out, time_in_state,
                                   jumps, status, total_w);
    if (rc) return rc;
    return smjp_sweep_host_end(c);
}

int smjp_sweep_host_begin(smjp_ctx_t* c, int C, const int64_t* p_offs, const int32_t* path_v,
                          const double* path_t, const int32_t* path_len, const int64_t* obs_offs,
                          const double* obs_t, const int32_t* obs_y, uint64_t seed, uint32_t sweep, int mode,
                          int precision, int64_t capacity, int64_t* w_offs, double* W, int32_t* path_v_out,
                          double* path_t_out, int32_t* path_len_out, double* time_in_state, int32_t* jumps,
                          int32_t* status, int64_t* total_w) {
    int rc = model_ready(c);
    if (rc) return rc;
    if (c->host_pending) {  // an earlier begin without its end: finish it first
        rc = smjp_sweep_host_end(c);
        if (rc) return rc;
    }
    if (C <= 0) return fail(SMJP_E_INVALID, "C <= 0");
    if (!p_offs || !path_v || !path_t || !path_len || !obs_offs || !w_offs || !W || !path_v_out ||
        !path_t_out || !path_len_out || !status || !total_w)
        return fail(SMJP_E_INVALID, "null required pointer");
    if (obs_offs[0] != 0) return fail(SMJP_E_INVALID, "offsets must start at 0");
    if (obs_offs[C] > 0 && (!obs_t || !obs_y)) return fail(SMJP_E_INVALID, "observations missing");
    CUDA_TRY(cudaSetDevice(c->device));
    const int K = c->K;
    if (!c->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_obs, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fill, cudaEventDisableTiming));
    }
    // Pinned (page-locked) caller buffers are addressable from the device: the kernels then write the new paths,
    // the per-chain metrics and the status words STRAIGHT into the caller's memory (a few MB of scattered
    // stores, overlapped with the sweep) instead of into a staging buffer that is copied back afterwards
    // (capacity-sized: 26 MB for cfg2).  Pageable buffers keep the staged copy.
    auto mapped = [&](void* hp) -> void* {
        if (!hp) return nullptr;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, hp) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
    };
    int32_t* m_pv = (int32_t*)mapped(path_v_out);
    double* m_pt = (double*)mapped(path_t_out);
    int32_t* m_pl = (int32_t*)mapped(path_len_out);
    int32_t* m_st = (int32_t*)mapped(status);
    double* m_tis = (double*)mapped(time_in_state);
    int32_t* m_jm = (int32_t*)mapped(jumps);
    const bool w_pinned = mapped(W) != nullptr;
    const bool in_pinned = mapped((void*)obs_t) != nullptr || !obs_t;
    // Zero-copy observations.  The sweep kernels ffbs_fast.cuh / ffbs_fastk.cuh fetch a chain's observation slice with
    // ONE bulk-async copy per array (TMA) when the chain starts; when the caller's arrays are pinned that copy can read
    // them in place over PCIe -- 12 KB per chain spread over the whole kernel -- instead of waiting for a 49 MB
    // (cfg2) upload before the kernel may start.  Only when those kernels will run (exact window, K <= 10, fp64 or
    // K >= 4, not switched off by rescue counts) and every slice fits their staging area; anything else uploads.
    const double* obs_t_m = nullptr;
    const int32_t* obs_y_m = nullptr;
    if (c->obs_zero_copy != 0 && obs_t && obs_y && obs_offs[C] > 0) {
        const int pidx = precision == SMJP_F64 ? 0 : 1;
        const bool fast_kernels = c->ffbs_fast != 0 && c->window_log2 == 0 && c->f32_rescue && K <= 10 &&
                                  (precision == SMJP_F64 || K >= 4) && !c->fast_off[pidx] && c->ffbs_warp != 1;
        int64_t dmax = 0;
        for (int i = 0; i < C; ++i) dmax = std::max(dmax, obs_offs[i + 1] - obs_offs[i]);
        const size_t esz = precision == SMJP_F64 ? 8 : 4;
        // smallest staging area those kernels have: the state ring (api.cu, ffbs launcher: ring sizes)
        const size_t ring_bytes = K <= 3 ? (size_t)2 * 10 * K * 32 * esz : (size_t)2 * (96 / (32 / K)) * (32 / K) * K * esz;
        const bool fits = (size_t)dmax * 12 + 64 <= ring_bytes;
        if ((c->obs_zero_copy == 1 || fast_kernels) && fits) {
            obs_t_m = (const double*)mapped((void*)obs_t);
            obs_y_m = (const int32_t*)mapped((void*)obs_y);
            if (!obs_t_m || !obs_y_m || (((uintptr_t)obs_t_m | (uintptr_t)obs_y_m) & 15)) obs_t_m = nullptr, obs_y_m = nullptr;
        }
    }
    StagedPaths sp;
    const bool gather_ok = c->obs_zero_copy != 0;  // (one switch for both in-place reads of pinned input)
    rc = stage_paths(c, C, p_offs, path_v, path_t, path_len, obs_offs, obs_t, obs_y, &sp, in_pinned, obs_t_m, obs_y_m,
                     gather_ok ? (const int32_t*)mapped((void*)path_v) : nullptr,
                     gather_ok ? (const double*)mapped((void*)path_t) : nullptr);
    if (rc) return rc;
    c->obs_in_place = obs_t_m != nullptr;
    rc = c->d_soff.reserve((size_t)p_offs[C] * 4 + 16);
    if (rc) return rc;
    // output staging: fixed part sized by `capacity`
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += smjp::align_up(bytes, 16); return o; };
    const size_t o_wo = take((C + 1) * 8), o_pl = take(C * 4), o_st = take(C * 4);
    const size_t o_tis = take((size_t)C * K * 8), o_jm = take((size_t)C * K * 4);
    const size_t o_W = take(capacity * 8);
    const size_t o_pv = m_pv ? 0 : take(capacity * 4), o_pt = m_pt ? 0 : take(capacity * 8);
    rc = c->d_out.reserve(off);
    if (rc) return rc;
    char* d = (char*)c->d_out.p;
    int64_t n_max = 0;
    c->w_copy_dst = w_pinned ? W : nullptr;
    rc = smjp_sweep_dev(c, C, sp.p_offs, sp.path_v, sp.path_t, sp.path_len, sp.obs_offs, sp.obs_t, sp.obs_y,
                        seed, sweep, mode, precision, capacity, (int64_t*)(d + o_wo), (int32_t*)c->d_soff.p,
                        (double*)(d + o_W), m_pv ? m_pv : (int32_t*)(d + o_pv), m_pt ? m_pt : (double*)(d + o_pt),
                        m_pl ? m_pl : (int32_t*)(d + o_pl), time_in_state ? (m_tis ? m_tis : (double*)(d + o_tis)) : nullptr,
                        jumps ? (m_jm ? m_jm : (int32_t*)(d + o_jm)) : nullptr, m_st ? m_st : (int32_t*)(d + o_st),
                        &n_max, total_w);
    c->w_copy_dst = nullptr;
    c->wait_obs = false;
    if (rc) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->stream);
        return rc;
    }
    const int64_t tw = *total_w;
    if (off <= ((size_t)1 << 18) && !w_pinned && !m_pv && !m_pt && !m_pl && !m_st && !m_tis && !m_jm) {
        // small batch (the drop-in's single chain), pageable buffers: ONE device-to-host copy of the whole staging
        // area through the pinned buffer instead of eight driver-staged ones (~10 us each)
        rc = c->h_out.reserve(off);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->h_out.p, d, off, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        const char* h = (const char*)c->h_out.p;
        memcpy(w_offs, h + o_wo, (C + 1) * 8);
        memcpy(path_len_out, h + o_pl, C * 4);
        memcpy(status, h + o_st, C * 4);
        if (time_in_state) memcpy(time_in_state, h + o_tis, (size_t)C * K * 8);
        if (jumps) memcpy(jumps, h + o_jm, (size_t)C * K * 4);
        memcpy(W, h + o_W, tw * 8);
        memcpy(path_v_out, h + o_pv, tw * 4);
        memcpy(path_t_out, h + o_pt, tw * 8);
        return SMJP_OK;
    }
    auto get = [&](void* dst, size_t o, size_t bytes) -> cudaError_t {
        return (dst && bytes) ? cudaMemcpyAsync(dst, d + o, bytes, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess;
    };
    CUDA_TRY(get(w_offs, o_wo, (C + 1) * 8));
    if (!m_pl) CUDA_TRY(get(path_len_out, o_pl, C * 4));
    if (!m_st) CUDA_TRY(get(status, o_st, C * 4));
    if (!m_tis) CUDA_TRY(get(time_in_state, o_tis, (size_t)C * K * 8));
    if (!m_jm) CUDA_TRY(get(jumps, o_jm, (size_t)C * K * 4));
    if (!w_pinned) CUDA_TRY(get(W, o_W, tw * 8));
    if (!m_pv) CUDA_TRY(get(path_v_out, o_pv, tw * 4));
    if (!m_pt) CUDA_TRY(get(path_t_out, o_pt, tw * 8));
    c->host_pending = true;  // smjp_sweep_host_end waits for the kernels and the copies
    return SMJP_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// dense HMM and particle filter (host-buffer entry points)
// ------------------------------------------------------------------------------------------------
namespace {

template <typename Real>
int hmm_launch(smjp_ctx* c, const HmmArgs& a) {
    const int S = a.S;
    const int spl = (S + 31) / 32;
    const int warps = 4;
    const size_t smem = (size_t)S * S * sizeof(Real);
    int grid = (a.C + warps - 1) / warps;
    if (grid > c->num_sms * 8) grid = c->num_sms * 8;
#define HMM_CASE(n)                                                                                      \
    case n:                                                                                              \
        CUDA_TRY(cudaFuncSetAttribute(hmm_dense_kernel<Real, n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        hmm_dense_kernel<Real, n><<<grid, warps * 32, smem, c->stream>>>(a);                            \
        break;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->time_kernels) {
        CUDA_TRY(cudaEventCreate(&e0));
        CUDA_TRY(cudaEventCreate(&e1));
        CUDA_TRY(cudaEventRecord(e0, c->stream));
    }
    switch (spl) {
        HMM_CASE(1) HMM_CASE(2) HMM_CASE(3) HMM_CASE(4)
        default: return fail(SMJP_E_UNSUPPORTED, "dense HMM kernel supports S <= 128 (S=%d)", S);
    }
#undef HMM_CASE
    CUDA_TRY(cudaGetLastError());
    if (c->time_kernels) {
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        c->hmm_events.emplace_back(e0, e1);
    }
    c->launches += 1;
    return SMJP_OK;
}

}  // namespace

extern "C" {

int smjp_hmm_dense_dev(smjp_ctx_t* c, int C, int S, const int64_t* t_offs, const double* P, int p_steps,
                       const double* E, const double* emis, int Kobs, const int32_t* y, const double* pi0,
                       const double* uniforms, uint64_t seed, uint32_t sweep, int precision, void* alpha_hat,
                       double* logc, double* loglik, int32_t* states, int32_t* status) {
    if (!c) return fail(SMJP_E_INVALID, "null context");
    if (C <= 0) return C == 0 ? SMJP_OK : fail(SMJP_E_INVALID, "C < 0");
    if (S < 1 || S > 128) return fail(SMJP_E_UNSUPPORTED, "dense HMM kernel supports 1 <= S <= 128 (S=%d)", S);
    if (!t_offs || !P || !pi0 || !alpha_hat || !logc || !loglik || !status) return fail(SMJP_E_INVALID, "null required pointer");
    if (!E && (!emis || !y || Kobs < 1)) return fail(SMJP_E_INVALID, "need dense E or an emission table with symbols");
    if (precision != SMJP_F64 && precision != SMJP_F32) return fail(SMJP_E_INVALID, "bad precision");
    CUDA_TRY(cudaSetDevice(c->device));
    HmmArgs a;
    memset(&a, 0, sizeof a);
    a.C = C;
    a.S = S;
    a.t_offs = t_offs;
    a.P = P;
    a.p_steps = p_steps > 0 ? p_steps : 0;
    a.E = E;
    a.emis = E ? nullptr : emis;
    a.y = E ? nullptr : y;
    a.Kobs = Kobs;
    a.pi0 = pi0;
    a.uniforms = uniforms;
    a.seed = seed;
    a.sweep = sweep;
    a.chain_base = c->chain_base;
    a.alpha_hat = alpha_hat;
    a.logc = logc;
    a.loglik = loglik;
    a.states = states;
    a.status = status;
    return precision == SMJP_F64 ? hmm_launch<double>(c, a) : hmm_launch<float>(c, a);
}

int smjp_hmm_dense_host(smjp_ctx_t* c, int C, int S, const int64_t* t_offs, const double* P, int p_steps,
                        const double* E, const double* emis, int Kobs, const int32_t* y, const double* pi0,
                        const double* uniforms, uint64_t seed, int precision, void* alpha_hat, double* logc,
                        double* loglik, int32_t* states, int32_t* status) {
    if (!c) return fail(SMJP_E_INVALID, "null context");
    if (C <= 0) return C == 0 ? SMJP_OK : fail(SMJP_E_INVALID, "C < 0");
    if (S < 1 || S > 128) return fail(SMJP_E_UNSUPPORTED, "dense HMM kernel supports 1 <= S <= 128 (S=%d)", S);
    if (!t_offs || !P || !pi0 || !alpha_hat || !status) return fail(SMJP_E_INVALID, "null required pointer");
    if (!E && (!emis || !y || Kobs < 1)) return fail(SMJP_E_INVALID, "need dense E or an emission table with symbols");
    if (precision != SMJP_F64 && precision != SMJP_F32) return fail(SMJP_E_INVALID, "bad precision");
    if (t_offs[0] != 0) return fail(SMJP_E_INVALID, "offsets must start at 0");
    const int64_t total = t_offs[C];
    for (int i = 0; i < C; ++i)
        if (t_offs[i + 1] <= t_offs[i]) return fail(SMJP_E_INVALID, "chain %d has no steps", i);
    if (!E)
        for (int64_t t = 0; t < total; ++t)
            if (y[t] < 0 || y[t] >= Kobs) return fail(SMJP_E_INVALID, "symbol %lld outside [0,Kobs)", (long long)t);
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t esz = precision == SMJP_F64 ? 8 : 4;
    const size_t nP = (size_t)(p_steps > 0 ? p_steps : 1) * S * S;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += smjp::align_up(bytes, 16); return o; };
    const size_t o_to = take((C + 1) * 8), o_P = take(nP * 8), o_pi = take(S * 8);
    const size_t o_E = E ? take((size_t)total * S * 8) : 0;
    const size_t o_em = E ? 0 : take((size_t)S * Kobs * 8), o_y = E ? 0 : take(total * 4);
    const size_t o_u = uniforms ? take(total * 8) : 0;
    const size_t in_bytes = off;
    int rc = c->h_in.reserve(in_bytes);
    if (rc) return rc;
    rc = c->d_in.reserve(in_bytes);
    if (rc) return rc;
    char* h = (char*)c->h_in.p;
    memcpy(h + o_to, t_offs, (C + 1) * 8);
    memcpy(h + o_P, P, nP * 8);
    memcpy(h + o_pi, pi0, S * 8);
    if (E) memcpy(h + o_E, E, (size_t)total * S * 8);
    else {
        memcpy(h + o_em, emis, (size_t)S * Kobs * 8);
        memcpy(h + o_y, y, total * 4);
    }
    if (uniforms) memcpy(h + o_u, uniforms, total * 8);
    CUDA_TRY(cudaMemcpyAsync(c->d_in.p, h, in_bytes, cudaMemcpyHostToDevice, c->stream));
    off = 0;
    const size_t o_st = take(C * 4), o_ll = take(C * 8), o_lc = take(total * 8), o_s = take(total * 4);
    const size_t o_ah = take((size_t)total * S * esz);
    rc = c->d_out.reserve(off);
    if (rc) return rc;
    rc = c->h_out.reserve(o_ah);
    if (rc) return rc;
    char* d = (char*)c->d_in.p;
    char* o = (char*)c->d_out.p;
    HmmArgs a;
    memset(&a, 0, sizeof a);
    a.C = C;
    a.S = S;
    a.t_offs = (const int64_t*)(d + o_to);
    a.P = (const double*)(d + o_P);
    a.p_steps = p_steps > 0 ? p_steps : 0;
    a.E = E ? (const double*)(d + o_E) : nullptr;
    a.emis = E ? nullptr : (const double*)(d + o_em);
    a.y = E ? nullptr : (const int32_t*)(d + o_y);
    a.Kobs = Kobs;
    a.pi0 = (const double*)(d + o_pi);
    a.uniforms = uniforms ? (const double*)(d + o_u) : nullptr;
    a.seed = seed;
    a.sweep = 0;
    a.chain_base = c->chain_base;
    a.alpha_hat = o + o_ah;
    a.logc = (double*)(o + o_lc);
    a.loglik = (double*)(o + o_ll);
    a.states = states ? (int32_t*)(o + o_s) : nullptr;
    a.status = (int32_t*)(o + o_st);
    rc = precision == SMJP_F64 ? hmm_launch<double>(c, a) : hmm_launch<float>(c, a);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->h_out.p, o, o_ah, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(alpha_hat, o + o_ah, (size_t)total * S * esz, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    char* ho = (char*)c->h_out.p;
    memcpy(status, ho + o_st, C * 4);
    if (loglik) memcpy(loglik, ho + o_ll, C * 8);
    if (logc) memcpy(logc, ho + o_lc, total * 8);
    if (states) memcpy(states, ho + o_s, total * 4);
    return SMJP_OK;
}

}  // extern "C"

namespace {

// workspaces + launch of the particle filter; the caller has set the input / output pointers of `a` (device memory)
int pf_launch(smjp_ctx* c, PfArgs a, int C, int N, int64_t Dmax, uint64_t seed, int resampler) {
    const int K = c->K;
    const int64_t Dw = Dmax > 0 ? Dmax : 1;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += smjp::align_up(bytes, 16); return o; };
    const size_t o_wv = take((size_t)C * 2 * N * 4), o_wl = take((size_t)C * 2 * N * 8), o_cdf = take((size_t)C * N * 8);
    const size_t o_anc = take((size_t)C * Dw * N * 4), o_line = take((size_t)C * Dw * 4), o_lz = take((size_t)C * Dw * 8 + 8);
    int rc = c->d_pfws.reserve(off);
    if (rc) return rc;
    char* o = (char*)c->d_pfws.p;
    a.C = C;
    a.N = N;
    a.K = K;
    a.Kobs = c->Kobs;
    a.model = (const double*)c->d_model.p;
    a.t_end = c->t_end;
    a.seed = seed;
    a.chain_base = c->chain_base;
    a.resampler = resampler;
    a.max_iters = 4096;
    a.extend = c->pf_extend;
    a.family = c->hazard_family;
    a.pv = (int32_t*)(o + o_wv);
    a.pl = (double*)(o + o_wl);
    a.cdf = (double*)(o + o_cdf);
    a.anc = (int32_t*)(o + o_anc);
    a.anc_stride = Dw * N;
    a.lineage = (int32_t*)(o + o_line);
    a.lineage_stride = Dw;
    if (!a.logZ) a.logZ = (double*)(o + o_lz);  // (per observation, ragged by obs_offs: never more than C * Dmax entries)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->time_kernels) {
        CUDA_TRY(cudaEventCreate(&e0));
        CUDA_TRY(cudaEventCreate(&e1));
        CUDA_TRY(cudaEventRecord(e0, c->stream));
    }
    {
        // cluster size: spread one filter run over as many SMs as the chain count leaves free
        int cs = 1;
        if (N >= 4096) {
            while (cs < 16 && (int64_t)C * cs * 2 <= c->num_sms && N / (cs * 2) >= 2048) cs *= 2;
        }
        if (c->pf_cluster > 0) cs = c->pf_cluster;
        if ((N + cs - 1) / cs > 32 * PF_MAXW)
            return fail(SMJP_E_UNSUPPORTED, "particle filter: at most %d particles per CTA (N=%d over a cluster of %d)",
                        32 * PF_MAXW, N, cs);
        const int threads = c->pf_threads > 0 ? c->pf_threads : (N / cs >= 4096 ? 1024 : 256);
        int n_clusters = C;
        if ((int64_t)n_clusters * cs > (int64_t)c->num_sms * 2) n_clusters = (c->num_sms * 2) / cs;
        if (n_clusters < 1) n_clusters = 1;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3((unsigned)(n_clusters * cs));
        cfg.blockDim = dim3((unsigned)threads);
        cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
#define PF_LAUNCH(BT, CSV)                                                                                        \
    do {                                                                                                          \
        if (CSV > 8) CUDA_TRY(cudaFuncSetAttribute(pf_kernel<BT, CSV>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); \
        CUDA_TRY(cudaLaunchKernelEx(&cfg, pf_kernel<BT, CSV>, a));                                                \
    } while (0)
#define PF_PICK(CSV)                                   \
    case CSV:                                          \
        if (threads == 1024) PF_LAUNCH(1024, CSV);     \
        else if (threads == 512) PF_LAUNCH(512, CSV);  \
        else PF_LAUNCH(256, CSV);                      \
        break;
        switch (cs) {
            PF_PICK(1) PF_PICK(2) PF_PICK(4) PF_PICK(8) PF_PICK(16)
            default: return fail(SMJP_E_INVALID, "pf_cluster must be 1, 2, 4, 8 or 16");
        }
#undef PF_PICK
#undef PF_LAUNCH
        c->last_pf_cluster = cs;
    }
    CUDA_TRY(cudaGetLastError());
    if (c->time_kernels) {
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        c->pf_events.emplace_back(e0, e1);
    }
    c->launches += 1;
    return SMJP_OK;
}

}  // namespace

extern "C" {

int smjp_pf_dev(smjp_ctx_t* c, int C, const int64_t* obs_offs, const double* obs_t, const int32_t* obs_y,
                const double* pi0, int N, uint64_t seed, int resampler, int64_t max_obs_per_chain, double* logZ,
                double* loglik, const int64_t* p_offs, int32_t* path_v, double* path_t, int32_t* path_len,
                int32_t* status) {
    int rc = model_ready(c);
    if (rc) return rc;
    if (C <= 0) return C == 0 ? SMJP_OK : fail(SMJP_E_INVALID, "C < 0");
    if (N < 1 || N > (1 << 24)) return fail(SMJP_E_INVALID, "N=%d outside [1, 2^24]", N);
    if (resampler != SMJP_RESAMPLE_MULTINOMIAL && resampler != SMJP_RESAMPLE_SYSTEMATIC)
        return fail(SMJP_E_INVALID, "bad resampler");
    if (max_obs_per_chain < 0) return fail(SMJP_E_INVALID, "max_obs_per_chain < 0");
    if (!obs_offs || !pi0 || !loglik || !p_offs || !path_v || !path_t || !path_len || !status)
        return fail(SMJP_E_INVALID, "null required pointer");
    CUDA_TRY(cudaSetDevice(c->device));
    PfArgs a;
    memset(&a, 0, sizeof a);
    a.obs_offs = obs_offs;
    a.obs_t = obs_t;
    a.obs_y = obs_y;
    a.pi0 = pi0;
    a.logZ = logZ;
    a.loglik = loglik;
    a.p_offs = p_offs;
    a.path_v = path_v;
    a.path_t = path_t;
    a.path_len = path_len;
    a.status = status;
    return pf_launch(c, a, C, N, max_obs_per_chain, seed, resampler);
}

int smjp_pf_host(smjp_ctx_t* c, int C, const int64_t* obs_offs, const double* obs_t, const int32_t* obs_y,
                 const double* pi0, int N, uint64_t seed, int resampler, double* logZ, double* loglik,
                 const int64_t* p_offs, int32_t* path_v, double* path_t, int32_t* path_len, int32_t* status) {
    int rc = model_ready(c);
    if (rc) return rc;
    if (C <= 0) return C == 0 ? SMJP_OK : fail(SMJP_E_INVALID, "C < 0");
    if (N < 1 || N > (1 << 24)) return fail(SMJP_E_INVALID, "N=%d outside [1, 2^24]", N);
    if (resampler != SMJP_RESAMPLE_MULTINOMIAL && resampler != SMJP_RESAMPLE_SYSTEMATIC)
        return fail(SMJP_E_INVALID, "bad resampler");
    if (!obs_offs || !pi0 || !loglik || !p_offs || !path_v || !path_t || !path_len || !status)
        return fail(SMJP_E_INVALID, "null required pointer");
    if (obs_offs[0] != 0 || p_offs[0] != 0) return fail(SMJP_E_INVALID, "offsets must start at 0");
    const int K = c->K;
    const int64_t total_o = obs_offs[C], total_p = p_offs[C];
    int64_t Dmax = 0;
    for (int i = 0; i < C; ++i) {
        const int64_t D = obs_offs[i + 1] - obs_offs[i];
        if (D < 0) return fail(SMJP_E_INVALID, "obs_offs not monotone");
        if (D > Dmax) Dmax = D;
        if (p_offs[i + 1] - p_offs[i] < 2) return fail(SMJP_E_INVALID, "path window of chain %d smaller than 2", i);
        for (int64_t d = obs_offs[i]; d < obs_offs[i + 1]; ++d) {
            if (obs_y[d] < 0 || obs_y[d] >= c->Kobs) return fail(SMJP_E_INVALID, "observation symbol outside [0,Kobs)");
            if (d > obs_offs[i] && !(obs_t[d] >= obs_t[d - 1])) return fail(SMJP_E_INVALID, "observation times not sorted");
        }
    }
    double ptot = 0;
    for (int s = 0; s < K; ++s) {
        if (!(pi0[s] >= 0)) return fail(SMJP_E_INVALID, "pi0[%d] negative", s);
        ptot += pi0[s];
    }
    if (!(ptot > 0)) return fail(SMJP_E_INVALID, "pi0 has no mass");
    CUDA_TRY(cudaSetDevice(c->device));
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += smjp::align_up(bytes, 16); return o; };
    const size_t o_oo = take((C + 1) * 8), o_ot = take(total_o * 8), o_oy = take(total_o * 4);
    const size_t o_po = take((C + 1) * 8), o_pi = take(K * 8);
    const size_t in_bytes = off;
    rc = c->h_in.reserve(in_bytes);
    if (rc) return rc;
    rc = c->d_in.reserve(in_bytes);
    if (rc) return rc;
    char* h = (char*)c->h_in.p;
    memcpy(h + o_oo, obs_offs, (C + 1) * 8);
    if (total_o) {
        memcpy(h + o_ot, obs_t, total_o * 8);
        memcpy(h + o_oy, obs_y, total_o * 4);
    }
    memcpy(h + o_po, p_offs, (C + 1) * 8);
    memcpy(h + o_pi, pi0, K * 8);
    CUDA_TRY(cudaMemcpyAsync(c->d_in.p, h, in_bytes, cudaMemcpyHostToDevice, c->stream));
    // outputs + workspaces
    off = 0;
    const size_t o_st = take(C * 4), o_pl = take(C * 4), o_ll = take(C * 8), o_lz = take(total_o * 8 + 8);
    const size_t o_pv = take(total_p * 4), o_pt = take(total_p * 8);
    const size_t out_bytes = off;
    rc = c->d_out.reserve(off);
    if (rc) return rc;
    rc = c->h_out.reserve(out_bytes);
    if (rc) return rc;
    char* d = (char*)c->d_in.p;
    char* o = (char*)c->d_out.p;
    PfArgs a;
    memset(&a, 0, sizeof a);
    a.obs_offs = (const int64_t*)(d + o_oo);
    a.obs_t = (const double*)(d + o_ot);
    a.obs_y = (const int32_t*)(d + o_oy);
    a.pi0 = (const double*)(d + o_pi);
    a.logZ = (double*)(o + o_lz);
    a.loglik = (double*)(o + o_ll);
    a.p_offs = (const int64_t*)(d + o_po);
    a.path_v = (int32_t*)(o + o_pv);
    a.path_t = (double*)(o + o_pt);
    a.path_len = (int32_t*)(o + o_pl);
    a.status = (int32_t*)(o + o_st);
    rc = pf_launch(c, a, C, N, Dmax, seed, resampler);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->h_out.p, o, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    char* ho = (char*)c->h_out.p;
    memcpy(status, ho + o_st, C * 4);
    memcpy(path_len, ho + o_pl, C * 4);
    memcpy(loglik, ho + o_ll, C * 8);
    if (logZ && total_o) memcpy(logZ, ho + o_lz, total_o * 8);
    memcpy(path_v, ho + o_pv, total_p * 4);
    memcpy(path_t, ho + o_pt, total_p * 8);
    return SMJP_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// dense K-state FFBS (shape == 1 collapse)
// ------------------------------------------------------------------------------------------------
namespace {

// forward pass on the tensor cores (dense_tc.cuh): tiles of 128 chains, longest grids first
int dense_tc_forward(smjp_ctx* c, const DenseArgs& a, int64_t n_max) {
    const size_t smem = sizeof(DenseTcSmem) + (size_t)a.Kobs * (TC_N + 1) * sizeof(float);
    if (smem > c->smem_optin) return fail(SMJP_E_UNSUPPORTED, "dense_tc: %zu B of shared memory", smem);
    int rc = c->d_tile.reserve(16);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(c->d_tile.p, 0, 16, c->stream));
    const int32_t* order = nullptr;
    if (a.C > TC_M && c->lpt_order && n_max > 0) {
        rc = c->d_order.reserve((size_t)a.C * sizeof(int32_t));
        if (rc) return rc;
        chain_order_kernel<<<1, 1024, 0, c->stream>>>(a.C, a.w_offs, (int)n_max, (int32_t*)c->d_order.p);
        CUDA_TRY(cudaGetLastError());
        c->launches += 1;
        order = (const int32_t*)c->d_order.p;
    }
    const int tiles = (a.C + TC_M - 1) / TC_M;
    int grid = tiles < 2 * c->num_sms ? tiles : 2 * c->num_sms;
    CUDA_TRY(cudaFuncSetAttribute(dense_tc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_tc_forward_kernel<<<grid, TC_THREADS, smem, c->stream>>>(a, order, (int*)c->d_tile.p);
    CUDA_TRY(cudaGetLastError());
    c->launches += 1;
    return SMJP_OK;
}

template <typename Real>
int dense_launch(smjp_ctx* c, DenseArgs a, int64_t n_max = 0) {
    const int K = a.K, spl = (K + 31) / 32, warps = DENSE_WARPS;
    const DensePlan plan = dense_plan(K, a.Kobs, sizeof(Real));
    const size_t smem = plan.bytes;
    int grid = (a.C + warps - 1) / warps;
    if (grid > c->num_sms * 8) grid = c->num_sms * 8;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->time_kernels) {
        CUDA_TRY(cudaEventCreate(&e0));
        CUDA_TRY(cudaEventCreate(&e1));
        CUDA_TRY(cudaEventRecord(e0, c->stream));
    }
    // many chains: the forward matvecs run as [128 chains x 64] x [64 x 64] products on the tensor cores (3xTF32)
    const bool tc = sizeof(Real) == 4 && K <= TC_N &&
                    (c->dense_tc == 1 || (c->dense_tc == -1 && K >= 32 && a.C >= 32768));
    c->last_dense_tc = tc ? 1 : 0;
    if (tc) {
        const int rct = dense_tc_forward(c, a, n_max);
        if (rct) return rct;
        a.fwd_done = 1;
    }
#define DENSE_CASE(n)                                                                                       \
    case n:                                                                                                 \
        CUDA_TRY(cudaFuncSetAttribute(dense_smjp_kernel<Real, n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dense_smjp_kernel<Real, n><<<grid, warps * 32, smem, c->stream>>>(a, plan.has_bt, plan.has_em);                              \
        break;
    switch (spl) {
        DENSE_CASE(1) DENSE_CASE(2) DENSE_CASE(3) DENSE_CASE(4)
        default: return fail(SMJP_E_UNSUPPORTED, "dense sMJP kernel supports K <= 128 (K=%d)", K);
    }
#undef DENSE_CASE
    CUDA_TRY(cudaGetLastError());
    if (c->time_kernels) {
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        c->hmm_events.emplace_back(e0, e1);
    }
    c->launches += 1;
    return SMJP_OK;
}

}  // namespace

extern "C" {

int smjp_ffbs_dense_dev(smjp_ctx_t* c, int C, const int64_t* w_offs, const double* W,
                        const int64_t* obs_offs, const double* obs_t, const int32_t* obs_y,
                        const double* u_state, const double* u_jump, uint64_t seed, uint32_t sweep,
                        int precision, double* logZ, int32_t* path_v, double* path_t, int32_t* path_len,
                        double* time_in_state, int32_t* jumps, int32_t* status, int64_t total_w) {
    int rc = model_ready(c);
    if (rc) return rc;
    rc = weibull_only(c, "the dense FFBS kernel");
    if (rc) return rc;
    if (C <= 0) return C == 0 ? SMJP_OK : fail(SMJP_E_INVALID, "C < 0");
    if (!w_offs || !W || !obs_offs || !path_v || !path_t || !path_len || !status)
        return fail(SMJP_E_INVALID, "null required pointer");
    if (precision != SMJP_F64 && precision != SMJP_F32) return fail(SMJP_E_INVALID, "bad precision");
    if (total_w < (int64_t)C * 2) return fail(SMJP_E_INVALID, "total_w inconsistent with C");
    if ((u_state == nullptr) != (u_jump == nullptr)) return fail(SMJP_E_INVALID, "give both u_state and u_jump, or neither");
    const int K = c->K;
    for (int t = 0; t < K * K; ++t)
        if (c->h_model[t] != 1.0)
            return fail(SMJP_E_INVALID, "the dense K-state path needs every Weibull shape to be 1 (entry %d is %g)", t,
                        c->h_model[t]);
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t esz = precision == SMJP_F64 ? 8 : 4;
    rc = c->d_dense_msg.reserve((size_t)total_w * K * esz);
    if (rc) return rc;
    DenseArgs a;
    memset(&a, 0, sizeof a);
    a.C = C;
    a.K = K;
    a.Kobs = c->Kobs;
    a.w_offs = w_offs;
    a.W = W;
    a.obs_offs = obs_offs;
    a.obs_t = obs_t;
    a.obs_y = obs_y;
    a.u_state = u_state;
    a.u_jump = u_jump;
    a.seed = seed;
    a.sweep = sweep;
    a.chain_base = c->chain_base;
    a.model = (const double*)c->d_model.p;
    a.obs_product = c->obs_product;
    a.omega = c->omega;
    a.power = c->power;
    a.t_end = c->t_end;
    a.msg = c->d_dense_msg.p;
    a.logZ = logZ;
    a.path_v = path_v;
    a.path_t = path_t;
    a.path_len = path_len;
    a.time_in_state = time_in_state;
    a.jumps = jumps;
    a.status = status;
    const int64_t n_max = c->dense_nmax_hint;  // set by smjp_sweep_dense_dev (tile ordering of the tensor-core forward)
    c->dense_nmax_hint = 0;
    return precision == SMJP_F64 ? dense_launch<double>(c, a, n_max) : dense_launch<float>(c, a, n_max);
}

int smjp_sweep_dense_dev(smjp_ctx_t* c, int C, const int64_t* p_offs, const int32_t* path_v,
                         const double* path_t, const int32_t* path_len, const int64_t* obs_offs,
                         const double* obs_t, const int32_t* obs_y, uint64_t seed, uint32_t sweep, int mode,
                         int precision, int64_t capacity, int64_t* w_offs, int32_t* soff, double* W,
                         int32_t* path_v_out, double* path_t_out, int32_t* path_len_out,
                         double* time_in_state, int32_t* jumps, int32_t* status, int64_t* n_max_host,
                         int64_t* total_w_host) {
    int rc = smjp_candidates_count_dev(c, C, p_offs, path_v, path_t, path_len, seed, sweep, w_offs, soff,
                                       n_max_host, total_w_host);
    if (rc) return rc;
    if (*total_w_host > capacity)
        return fail(SMJP_E_NOMEM, "sweep needs %lld grid points, buffers hold %lld", (long long)*total_w_host,
                    (long long)capacity);
    rc = smjp_candidates_fill_dev(c, C, p_offs, path_v, path_t, path_len, seed, sweep, mode, w_offs, soff, W);
    if (rc) return rc;
    c->dense_nmax_hint = *n_max_host;
    return smjp_ffbs_dense_dev(c, C, w_offs, W, obs_offs, obs_t, obs_y, nullptr, nullptr, seed, sweep, precision,
                               nullptr, path_v_out, path_t_out, path_len_out, time_in_state, jumps, status,
                               *total_w_host);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// parameter step: holding-time statistics + Weibull shape estimate
// ------------------------------------------------------------------------------------------------
extern "C" int smjp_shape_mle_dev(smjp_ctx_t* c, int C, int K, const int64_t* p_offs, const int32_t* path_v,
                                  const double* path_t, const int32_t* path_len, int pooled, double grid,
                                  double grid_max, double* khat, double* gval, int32_t* count, double* sum_log,
                                  double* sum_t) {
    if (!c) return fail(SMJP_E_INVALID, "null context");
    if (K < 1 || K > 1024) return fail(SMJP_E_INVALID, "K=%d outside [1,1024]", K);
    if (C <= 0) return C == 0 ? SMJP_OK : fail(SMJP_E_INVALID, "C < 0");
    if (!p_offs || !path_v || !path_t || !path_len || !khat || !gval || !count || !sum_log || !sum_t)
        return fail(SMJP_E_INVALID, "null required pointer");
    if (!(grid > 0) || !(grid_max > grid)) return fail(SMJP_E_INVALID, "need 0 < grid < grid_max");
    if ((grid_max - grid) / grid > 1e8) return fail(SMJP_E_INVALID, "shape grid has more than 1e8 points");
    CUDA_TRY(cudaSetDevice(c->device));
    ShapeMleArgs a;
    memset(&a, 0, sizeof a);
    a.C = C;
    a.K = K;
    a.pooled = pooled ? 1 : 0;
    a.p_offs = p_offs;
    a.path_v = path_v;
    a.path_t = path_t;
    a.path_len = path_len;
    a.grid = grid;
    a.grid_max = grid_max;
    a.khat = khat;
    a.gval = gval;
    a.count = count;
    a.sum_log = sum_log;
    a.sum_t = sum_t;
    const int64_t blocks = (int64_t)(pooled ? 1 : C) * K * K;
    if (blocks > 0x7fffffff) return fail(SMJP_E_INVALID, "too many (chain, v, v') groups");
    shape_mle_kernel<<<(unsigned)blocks, MLE_THREADS, 0, c->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    c->launches += 1;
    return SMJP_OK;
}
