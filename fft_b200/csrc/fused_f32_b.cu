// fused kernels, fp32, the headline size and larger (configs chosen from profiles/kbench_r01*.txt)
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f32_b(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED_PF(float, 4096, 16, 16, 16, 1, 256, 1, 2));          // 88 % of HBM peak (84 % without TMA prefetch)
    v.push_back(SSFFT_FUSED_REAL(float, 4096, 16, 16, 16, 1, 256, 1, 4, 4, 2));  // 4 CTAs/SM, in-place staging: R2C / C2R 68 -> 74 %
    // PF = 2 (prefetch lands in the exchange buffer): half the shared memory -> 2 CTAs/SM for 8192, and prefetch at all for 16384
    v.push_back(SSFFT_FUSED_X(float, 8192, 32, 16, 16, 1, 256, 1, 2, 5, 2));     // 79 % (PF = 1, 1 CTA/SM: 77 %; no prefetch: 51 %)
    v.push_back(SSFFT_FUSED_X(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 2));    // 60 % (55 % without prefetch); real 49/43 -> 51/52 %
}
}  // namespace ssfft
