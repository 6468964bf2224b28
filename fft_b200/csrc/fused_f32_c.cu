// fused kernels, fp32, mixed-radix sizes of BASELINE config 4.  No shared-memory padding: odd strides are
// conflict-free by themselves and the power-of-two padding doubled the wavefronts (bank-conflict model).
#include "fused_launch.cuh"
namespace ssfft {
void register_fused_f32_c(std::vector<FusedEntry> &v) {
    // complex cores of RealFFT<float>(1000) / (6000) (profiles/kbench_tune8_r01.txt): real 1000 / 6000 went from the
    // generic kernel + separate twiddle passes (16 / 22 % of roofline) to 59-68 %
    v.push_back(SSFFT_FUSED_X(float, 500, 10, 10, 5, 1, 50, 4, 4, 31, 1));
    v.push_back(SSFFT_FUSED_X(float, 3000, 25, 12, 10, 1, 125, 1, 4, 31, 1));
    v.push_back(SSFFT_FUSED_X(float, 1000, 10, 10, 10, 1, 100, 2, 5, 31, 1));
    v.push_back(SSFFT_FUSED_X(float, 2187, 9, 9, 27, 1, 81, 2, 3, 31, 1));   // 70 % (27x9x9 without prefetch: 63 %)
    v.push_back(SSFFT_FUSED_X(float, 3125, 25, 25, 5, 1, 125, 1, 5, 31, 0));
    // (a 3-pass 16 x 15 x 25 variant with ragged passes measured 43 % vs 59 % for this one)
    // three passes 25 x 24 x 10 on 250 threads (ragged first and last pass): 77 % -- the four-pass 10x10x10x6 reached 59 %,
    // and the ORDER matters: 24x25x10 only 51 % (profiles/kbench_tune2_r01.txt, kbench_tune3_r01.txt)
    v.push_back(SSFFT_FUSED_X(float, 6000, 25, 24, 10, 1, 250, 1, 2, 31, 1));
}
}  // namespace ssfft
