#include "ffbs_launch.cuh"
namespace smjp {
cudaError_t ffbs_dispatch_f64_hi(const FfbsArgs& a, FfbsLaunch& L, bool query) { return ffbs_dispatch_hi<double>(a, L, query); }
}
