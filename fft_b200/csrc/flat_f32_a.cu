// ticket-queue four-step kernels (flat.cuh), fp32, 2^16.  The first entry of a size is its default (measured, profiles/
// flat_ab_r02e.txt / _r02f.txt); the other is selected with SSFFT_FLAT_VARIANT="ring,ctas_per_sm,inplace".  Entries with a
// separate exchange buffer also carry the RealFFT kernels of length 2 N1 N2.
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_a(std::vector<FlatEntry> &v) {
    v.push_back(make_flat_entry<TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, 2, 3, true, 3>("float_flat_256x256_r2c3i"));
    v.push_back(make_flat_entry<TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, 1, 3, false>("float_flat_256x256_r1c3x"));
    v.push_back(make_flat_entry<TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, 1, 3, true, 3>("float_flat_256x256_r1c3i"));
}
}  // namespace ssfft
