// fused kernels, fp64, power-of-two sizes (16-byte elements: one padding element per 8)
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f64_a(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED_X(double, 64, 8, 8, 1, 1, 8, 32, 2, 3, 0));
    v.push_back(SSFFT_FUSED_X(double, 128, 8, 4, 4, 1, 16, 16, 2, 3, 0));
    v.push_back(SSFFT_FUSED_X(double, 256, 8, 8, 4, 1, 32, 8, 2, 3, 0));
    v.push_back(SSFFT_FUSED_X(double, 512, 8, 8, 8, 1, 64, 4, 3, 3, 0));
    v.push_back(SSFFT_FUSED_X(double, 1024, 8, 8, 4, 4, 128, 2, 3, 3, 1));  // 75 %
    v.push_back(SSFFT_FUSED_X(double, 2048, 8, 8, 8, 4, 256, 1, 2, 3, 1));  // 73 %
    v.push_back(SSFFT_FUSED_X(double, 4096, 8, 8, 8, 8, 512, 1, 1, 3, 1));  // 64 %
}
}  // namespace ssfft
