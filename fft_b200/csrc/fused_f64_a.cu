// fused kernels, fp64, power-of-two sizes (16-byte elements: one padding element per 8)
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f64_a(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED_X(double, 64, 8, 8, 1, 1, 8, 32, 2, 3, 1));       // TMA prefetch: 93 -> 97 %
    v.push_back(SSFFT_FUSED_X(double, 128, 16, 8, 1, 1, 8, 16, 3, 3, 1));     // two passes + prefetch: 77 -> 93 %
    v.push_back(SSFFT_FUSED_X(double, 256, 16, 16, 1, 1, 16, 8, 3, 3, 1));    // two radix-16 passes + prefetch: 71 -> 92 %
    v.push_back(SSFFT_FUSED_X(double, 512, 8, 8, 8, 1, 64, 2, 4, 3, 1));      // 80 -> 96 %
    v.push_back(SSFFT_FUSED_X(double, 1024, 8, 8, 16, 1, 64, 2, 3, 3, 1));   // three passes: 64-72 -> 97 % (kbench_tune4)
    v.push_back(SSFFT_FUSED_X(double, 2048, 8, 16, 16, 1, 128, 1, 2, 3, 1));  // three passes: 76 -> 90 %
    v.push_back(SSFFT_FUSED_X(double, 4096, 16, 16, 16, 1, 256, 1, 1, 4, 1));  // three radix-16 passes: 66 -> 80 %
    v.push_back(SSFFT_FUSED_X(double, 8192, 16, 16, 32, 1, 256, 1, 1, 3, 2));  // in-place staging: 69 % (no prefetch 54 %, four-step tiles 33 %)
}
}  // namespace ssfft
