// fused kernels, fp64, mixed-radix sizes of BASELINE config 4 (no padding, see fused_f32_c.cu)
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f64_b(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED_X(double, 1000, 10, 10, 10, 1, 100, 2, 2, 31, 0));
    v.push_back(SSFFT_FUSED_X(double, 2187, 9, 9, 9, 3, 243, 1, 2, 31, 0));
    v.push_back(SSFFT_FUSED_X(double, 3125, 25, 25, 5, 1, 125, 2, 1, 31, 0));
    v.push_back(SSFFT_FUSED_X(double, 6000, 10, 10, 10, 6, 200, 1, 1, 31, 0));
}
}  // namespace ssfft
