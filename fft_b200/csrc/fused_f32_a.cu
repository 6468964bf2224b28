// fused kernels, fp32, power-of-two sizes below the headline (configs chosen from profiles/kbench_r01*.txt,
// shared-memory padding from the offline bank-conflict model)
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f32_a(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED(float, 64, 8, 8, 1, 1, 8, 32, 2));
    v.push_back(SSFFT_FUSED(float, 128, 16, 8, 1, 1, 8, 32, 2));
    v.push_back(SSFFT_FUSED(float, 256, 16, 16, 1, 1, 16, 8, 4));             // 96 % of HBM peak
    v.push_back(SSFFT_FUSED_X(float, 512, 32, 16, 1, 1, 16, 8, 4, 5, 0));     // 93 %+
    v.push_back(SSFFT_FUSED_X(float, 1024, 32, 32, 1, 1, 32, 4, 2, 5, 1));    // 97 %
    v.push_back(SSFFT_FUSED(float, 2048, 16, 16, 8, 1, 128, 1, 6));           // 91 %
    // RealFFT plans (complex core of half the real length), profiles/kbench_real_r01.txt:
    v.push_back(SSFFT_FUSED_REAL(float, 128, 16, 8, 1, 1, 8, 16, 4, 4, 0));   // C2R 62 -> 69 %
    v.push_back(SSFFT_FUSED_REAL(float, 512, 32, 16, 1, 1, 16, 8, 3, 5, 1));  // TMA staging: R2C 82 -> 87 %, C2R 60 -> 87 %
}
}  // namespace ssfft
