// ticket-queue four-step kernels (flat.cuh), fp32, 2^16.  The first entry of a size is its default (measured, profiles/
// flat_ab_r02e.txt / _r02f.txt); the other is selected with SSFFT_FLAT_VARIANT="ring,ctas_per_sm,inplace".  Entries with a
// separate exchange buffer also carry the RealFFT kernels of length 2 N1 N2.
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_a(std::vector<FlatEntry> &v) {
    v.push_back(make_flat_entry<TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, 2, 3, true, 0>("float_flat_256x256_r2c3i"));
    v.push_back(make_flat_entry<TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, 1, 3, false, 1>("float_flat_256x256_r1c3x"));   // RealFFT forward of 2^17 (49.7 %)
    v.push_back(make_flat_entry<TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, 1, 3, true, 2>("float_flat_256x256_r1c3i"));   // RealFFT inverse of 2^17 (44.8 % vs 43.7)
}
}  // namespace ssfft
