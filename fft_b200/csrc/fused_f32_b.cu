// fused kernels, fp32, the headline size and larger (configs chosen from profiles/kbench_r01*.txt)
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f32_b(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED_PF(float, 4096, 16, 16, 16, 1, 256, 1, 2));          // 88 % of HBM peak (84 % without TMA prefetch)
    v.push_back(SSFFT_FUSED_REAL(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 1));  // 3 CTAs/SM: R2C / C2R 68 -> 70 %
    v.push_back(SSFFT_FUSED_X(float, 8192, 32, 16, 16, 1, 256, 1, 1, 5, 1));     // 72 % (51 % without prefetch)
    v.push_back(SSFFT_FUSED_X(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 0));    // 55 %; staging buffer does not fit beside 128 KiB
}
}  // namespace ssfft
