// fused kernels, fp64, mixed-radix sizes of BASELINE config 4 (no padding, see fused_f32_c.cu)
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f64_b(std::vector<FusedEntry> &v) {
    v.push_back(SSFFT_FUSED_X(double, 500, 25, 20, 1, 1, 25, 4, 2, 31, 1));     // core of RealFFT<double>(1000)
    v.push_back(SSFFT_FUSED_X(double, 3000, 25, 12, 10, 1, 125, 1, 2, 31, 1));  // core of RealFFT<double>(6000)
    v.push_back(SSFFT_FUSED_X(double, 1000, 10, 10, 10, 1, 100, 2, 3, 31, 1));   // TMA prefetch, 3 CTAs/SM: 67 -> 84 %
    v.push_back(SSFFT_FUSED_X(double, 2187, 9, 9, 9, 3, 243, 1, 2, 31, 1));     // TMA prefetch: 63 -> 72 %
    v.push_back(SSFFT_FUSED_X(double, 3125, 25, 25, 5, 1, 125, 1, 2, 31, 1));    // one transform per CTA + prefetch: 51 -> 77 %
    v.push_back(SSFFT_FUSED_X(double, 6000, 25, 24, 10, 1, 250, 1, 1, 31, 1));   // three passes + prefetch: 41 -> 67 %
}
}  // namespace ssfft
