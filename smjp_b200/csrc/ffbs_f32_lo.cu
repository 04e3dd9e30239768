#include "ffbs_launch.cuh"
namespace smjp {
cudaError_t ffbs_dispatch_f32_lo(const FfbsArgs& a, FfbsLaunch& L, bool query) { return ffbs_dispatch_lo<float>(a, L, query); }
}
