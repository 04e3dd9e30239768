#include "ffbs_launch.cuh"
namespace smjp {
cudaError_t ffbs_dispatch_f32_hi(const FfbsArgs& a, FfbsLaunch& L, bool query) { return ffbs_dispatch_hi<float>(a, L, query); }
}
