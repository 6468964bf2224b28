// fused kernels, fp32, mixed-radix sizes of BASELINE config 4
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f32_c(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED(float, 1000, 10, 10, 10, 1, 100, 2, 2));
    v.push_back(SSFFT_FUSED(float, 2187, 27, 9, 9, 1, 81, 3, 2));
    v.push_back(SSFFT_FUSED(float, 3125, 25, 25, 5, 1, 125, 2, 2));
    v.push_back(SSFFT_FUSED(float, 6000, 10, 10, 10, 6, 200, 1, 2));
}
}  // namespace ssfft
