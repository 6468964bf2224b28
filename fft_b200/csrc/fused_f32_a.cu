// fused kernels, fp32, power-of-two sizes
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f32_a(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED(float, 64, 8, 8, 1, 1, 8, 32, 2));
    v.push_back(SSFFT_FUSED(float, 128, 16, 8, 1, 1, 8, 32, 2));
    v.push_back(SSFFT_FUSED(float, 256, 16, 16, 1, 1, 16, 16, 2));
    v.push_back(SSFFT_FUSED(float, 512, 8, 8, 8, 1, 64, 4, 2));
    v.push_back(SSFFT_FUSED(float, 1024, 16, 16, 4, 1, 64, 4, 2));
    v.push_back(SSFFT_FUSED(float, 2048, 16, 16, 8, 1, 128, 2, 2));
}
}  // namespace ssfft
