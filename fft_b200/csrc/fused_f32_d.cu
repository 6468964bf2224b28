// fused kernels, fp32: the reference's "fast sizes" 2^k * 3 and 2^k * 9 (FFT::sizeMinimum / sizeMaximum,
// signalsmith-fft.h:317-348 steer users to these) and a 3-pass variant of 6000.  Passes whose butterfly
// count is not a multiple of the thread count are "ragged" (excess threads idle behind a predicate).
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f32_d(std::vector<FusedEntry> &v) {
    //                         N     radices        TX  FPB MINB pad PF
    v.push_back(SSFFT_FUSED_X(float, 96, 16, 6, 1, 1, 6, 32, 2, 4, 1));        // TMA prefetch: 68 -> 78 %
    v.push_back(SSFFT_FUSED_X(float, 192, 16, 12, 1, 1, 12, 16, 2, 4, 1));     // 67 -> 85 %
    v.push_back(SSFFT_FUSED_X(float, 384, 16, 24, 1, 1, 24, 8, 2, 4, 1));      // TMA prefetch: 74 -> 91 %
    v.push_back(SSFFT_FUSED_X(float, 768, 16, 16, 3, 1, 48, 4, 3, 4, 1));      // 77 -> 92 %
    v.push_back(SSFFT_FUSED_X(float, 1536, 16, 16, 6, 1, 96, 2, 3, 4, 1));     // TMA prefetch: 78 -> 90 %
    v.push_back(SSFFT_FUSED_X(float, 3072, 16, 16, 12, 1, 192, 1, 3, 4, 1));   // 74 -> 89 %
    v.push_back(SSFFT_FUSED_X(float, 6144, 16, 16, 24, 1, 384, 1, 2, 4, 2));   // 52 -> 78 % (in-place staging, 2 CTAs/SM)
    v.push_back(SSFFT_FUSED_X(float, 144, 16, 9, 1, 1, 9, 16, 2, 4, 1));       // 63 -> 74 %
    v.push_back(SSFFT_FUSED_X(float, 288, 16, 18, 1, 1, 18, 8, 2, 4, 1));      // 66 -> 79 %
    v.push_back(SSFFT_FUSED_X(float, 576, 16, 4, 9, 1, 36, 4, 3, 4, 1));       // 65 -> 76 %
    v.push_back(SSFFT_FUSED_X(float, 1152, 16, 8, 9, 1, 72, 2, 3, 4, 1));      // 69 -> 81 %
    v.push_back(SSFFT_FUSED_X(float, 2304, 16, 16, 9, 1, 144, 2, 2, 4, 1));    // 68 -> 80 %
    v.push_back(SSFFT_FUSED_X(float, 4608, 16, 16, 18, 1, 288, 1, 3, 4, 2));   // 67 -> 83 % (in-place staging, 3 CTAs/SM)
    v.push_back(SSFFT_FUSED_X(float, 9216, 32, 16, 18, 1, 288, 1, 2, 5, 2));   // 51 -> 70 % (in-place staging, 2 CTAs/SM)
}
}  // namespace ssfft
