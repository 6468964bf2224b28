// fused kernels, fp32 and fp64: the lengths of the reference's benchmark set (benchmark/benchmark.h:27-52) between the
// one-thread-per-transform kernel (N <= 24, tiny.cuh) and the first registered single-pass sizes: 32, 36, 48, 72.
// They ran through the pass interpreter at 42-63 % of the roofline.
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f32_e(std::vector<FusedEntry> &v) {
    //                         N     radices        TX  FPB MINB pad PF
    v.push_back(SSFFT_FUSED_X(float, 32, 8, 4, 1, 1, 4, 64, 2, 4, 1));
    v.push_back(SSFFT_FUSED_X(float, 36, 6, 6, 1, 1, 6, 32, 2, 4, 1));
    v.push_back(SSFFT_FUSED_X(float, 48, 8, 6, 1, 1, 6, 32, 2, 4, 1));
    v.push_back(SSFFT_FUSED_X(float, 72, 8, 9, 1, 1, 9, 16, 3, 4, 1));
    v.push_back(SSFFT_FUSED_X(double, 32, 8, 4, 1, 1, 4, 64, 2, 3, 1));
    v.push_back(SSFFT_FUSED_X(double, 36, 6, 6, 1, 1, 6, 32, 2, 3, 1));
    v.push_back(SSFFT_FUSED_X(double, 48, 8, 6, 1, 1, 6, 32, 2, 3, 1));
    v.push_back(SSFFT_FUSED_X(double, 72, 8, 9, 1, 1, 9, 16, 3, 3, 1));
    // fp64 4608 / 6144 / 9216 (the pass interpreter / generic four-step reached 31 / 31 / 23 %): in-place staging, 16 or 18
    // points per thread
    v.push_back(SSFFT_FUSED_X(double, 4608, 16, 16, 18, 1, 288, 1, 2, 3, 2));
    v.push_back(SSFFT_FUSED_X(double, 6144, 8, 8, 8, 12, 512, 1, 1, 3, 2));
    v.push_back(SSFFT_FUSED_X(double, 9216, 8, 8, 12, 12, 768, 1, 1, 3, 2));
}
}  // namespace ssfft
