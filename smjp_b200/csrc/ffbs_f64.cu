#include "ffbs_launch.cuh"
namespace smjp {
cudaError_t ffbs_dispatch_f64(const FfbsArgs& a, FfbsLaunch& L, bool query) { return ffbs_dispatch<double>(a, L, query); }
}
