// fused kernels, fp64: fast sizes 2^k * 3 and 2^k * 9 (see fused_f32_d.cu)
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f64_c(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED_X(double, 96, 8, 12, 1, 1, 12, 16, 2, 3, 1));     // TMA prefetch: 78 -> 93 %
    v.push_back(SSFFT_FUSED_X(double, 192, 8, 8, 3, 1, 24, 8, 2, 3, 1));      // 74 -> 89 %
    v.push_back(SSFFT_FUSED_X(double, 384, 8, 8, 6, 1, 48, 4, 2, 3, 1));      // 70 -> 91 %
    v.push_back(SSFFT_FUSED_X(double, 768, 8, 8, 12, 1, 96, 2, 2, 3, 1));     // TMA prefetch: 69 -> 92 %
    v.push_back(SSFFT_FUSED_X(double, 1536, 8, 8, 8, 3, 192, 1, 2, 3, 1));    // 58 -> 73 %
    v.push_back(SSFFT_FUSED_X(double, 3072, 16, 16, 12, 1, 192, 1, 2, 3, 2));  // 50 -> 75 %
    v.push_back(SSFFT_FUSED_X(double, 144, 8, 18, 1, 1, 18, 8, 2, 3, 1));     // 60 -> 76 %
    v.push_back(SSFFT_FUSED_X(double, 288, 8, 4, 9, 1, 36, 4, 2, 3, 1));      // 62 -> 80 %
    v.push_back(SSFFT_FUSED_X(double, 576, 8, 8, 9, 1, 72, 2, 2, 3, 1));      // 63 -> 83 %
    v.push_back(SSFFT_FUSED_X(double, 1152, 8, 8, 18, 1, 144, 1, 2, 3, 1));   // 54 -> 76 %
    v.push_back(SSFFT_FUSED_X(double, 2304, 8, 8, 4, 9, 288, 1, 2, 3, 1));    // 60 -> 78 %
}
}  // namespace ssfft
