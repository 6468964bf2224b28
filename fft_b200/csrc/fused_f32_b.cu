// fused kernels, fp32, the headline size and larger
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f32_b(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED(float, 4096, 16, 16, 16, 1, 256, 1, 2));
    v.push_back(SSFFT_FUSED(float, 8192, 32, 16, 16, 1, 256, 1, 1));
    v.push_back(SSFFT_FUSED(float, 16384, 32, 32, 16, 1, 512, 1, 1));
}
}  // namespace ssfft
