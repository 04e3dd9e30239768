#include "ffbs_launch.cuh"
namespace smjp {
cudaError_t ffbs_dispatch_f32(const FfbsArgs& a, FfbsLaunch& L, bool query) { return ffbs_dispatch<float>(a, L, query); }
}
