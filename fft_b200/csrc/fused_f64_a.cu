// fused kernels, fp64, power-of-two sizes (configs chosen from profiles/kbench_r01*.txt)
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f64_a(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED(double, 64, 8, 8, 1, 1, 8, 32, 2));
    v.push_back(SSFFT_FUSED(double, 128, 8, 4, 4, 1, 16, 16, 2));
    v.push_back(SSFFT_FUSED(double, 256, 8, 8, 4, 1, 32, 8, 2));
    v.push_back(SSFFT_FUSED(double, 512, 8, 8, 8, 1, 64, 4, 3));
    v.push_back(SSFFT_FUSED_PF(double, 1024, 8, 8, 4, 4, 128, 2, 3));  // 75 %
    v.push_back(SSFFT_FUSED_PF(double, 2048, 8, 8, 8, 4, 256, 1, 2));  // 73 %
    v.push_back(SSFFT_FUSED_PF(double, 4096, 8, 8, 8, 8, 512, 1, 1));  // 64 %
}
}  // namespace ssfft
