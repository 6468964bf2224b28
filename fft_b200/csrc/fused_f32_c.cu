// fused kernels, fp32, mixed-radix sizes of BASELINE config 4 (configs chosen from profiles/kbench_r01*.txt)
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f32_c(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED_PF(float, 1000, 10, 10, 10, 1, 100, 2, 4));  // 67 %
    v.push_back(SSFFT_FUSED(float, 2187, 27, 9, 9, 1, 81, 3, 2));        // 61 %
    v.push_back(SSFFT_FUSED(float, 3125, 25, 25, 5, 1, 125, 1, 5));      // 64 %
    v.push_back(SSFFT_FUSED_PF(float, 6000, 10, 10, 10, 6, 200, 1, 2));  // 50 %
}
}  // namespace ssfft
