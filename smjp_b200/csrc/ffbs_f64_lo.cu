#include "ffbs_launch.cuh"
namespace smjp {
cudaError_t ffbs_dispatch_f64_lo(const FfbsArgs& a, FfbsLaunch& L, bool query) { return ffbs_dispatch_lo<double>(a, L, query); }
}
