// fused kernels, fp32, power-of-two sizes below the headline (configs chosen from profiles/kbench_r01*.txt,
// shared-memory padding from the offline bank-conflict model)
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f32_a(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED_X(float, 64, 8, 8, 1, 1, 8, 32, 2, 4, 1));        // TMA prefetch: 72 -> 82 %
    v.push_back(SSFFT_FUSED_X(float, 128, 16, 8, 1, 1, 8, 32, 2, 4, 1));      // 83 -> 94 %
    v.push_back(SSFFT_FUSED_X(float, 256, 16, 16, 1, 1, 16, 8, 4, 4, 1));     // 97 -> 100 % of the measured copy peak
    v.push_back(SSFFT_FUSED_X(float, 512, 32, 16, 1, 1, 16, 8, 3, 5, 1));     // TMA prefetch, 3 CTAs/SM: 96 -> 100 % (real: 82/60 -> 87/87 %)
    v.push_back(SSFFT_FUSED_REAL(float, 64, 8, 8, 1, 1, 8, 32, 3, 4, 1));     // RealFFT 128: 3 CTAs/SM: 68/74 -> 70/81 %
    v.push_back(SSFFT_FUSED_REAL(float, 128, 16, 8, 1, 1, 8, 16, 4, 4, 1));   // RealFFT 256: 69/77 -> 73/80 %
    v.push_back(SSFFT_FUSED_REAL(float, 512, 32, 16, 1, 1, 16, 8, 4, 5, 2));  // RealFFT 1024: in-place staging, 4 CTAs/SM: 83/84 -> 89/92 %
    v.push_back(SSFFT_FUSED_X(float, 1024, 32, 32, 1, 1, 32, 4, 3, 5, 1));    // 3 CTAs/SM: 99 -> 101 % of the measured copy peak; C2R 77 -> 84 %
    v.push_back(SSFFT_FUSED_X(float, 2048, 16, 16, 8, 1, 128, 1, 6, 4, 2));   // in-place staging, 6 CTAs/SM: 90 -> 94 %
}
}  // namespace ssfft
