// ticket-queue four-step kernels (flat.cuh), fp32, 2^19 and 2^20 (1024-point leg: radix 16 x 8 x 8, 4 lanes per tile).  The first entry of a size is its default (measured, profiles/
// flat_ab_r02e.txt / _r02f.txt); the other is selected with SSFFT_FLAT_VARIANT="ring,ctas_per_sm,inplace".  Entries with a
// separate exchange buffer also carry the RealFFT kernels of length 2 N1 N2.
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_d(std::vector<FlatEntry> &v) {
    v.push_back(make_flat_entry<TileCfg<float, 512, 8, 8, 8, 32, 8, 3>, TileCfg<float, 1024, 16, 8, 8, 64, 4, 3>, 1, 3, false>("float_flat_512x1024_r1c3x"));
    v.push_back(make_flat_entry<TileCfg<float, 512, 8, 8, 8, 32, 8, 3>, TileCfg<float, 1024, 16, 8, 8, 64, 4, 3>, 2, 3, true>("float_flat_512x1024_r2c3i"));
    v.push_back(make_flat_entry<TileCfg<float, 1024, 16, 8, 8, 64, 4, 3>, TileCfg<float, 1024, 16, 8, 8, 64, 4, 3>, 1, 3, false>("float_flat_1024x1024_r1c3x"));
    v.push_back(make_flat_entry<TileCfg<float, 1024, 16, 8, 8, 64, 4, 3>, TileCfg<float, 1024, 16, 8, 8, 64, 4, 3>, 2, 3, true>("float_flat_1024x1024_r2c3i"));
    // first pass of radix 4: with 4 lanes per tile the first-pass stores of radix 16 hit rows 16 apart = the same banks
    // (ncu r02c: 25 M conflicts on 47 M shared-memory wavefronts); rows 4 apart spread over all of them
    v.push_back(make_flat_entry<TileCfg<float, 1024, 4, 16, 16, 64, 4, 3>, TileCfg<float, 1024, 4, 16, 16, 64, 4, 3>, 2, 3, true>("float_flat_1024x1024_p4_r2c3i"));
    v.push_back(make_flat_entry<TileCfg<float, 512, 8, 8, 8, 32, 8, 3>, TileCfg<float, 1024, 4, 16, 16, 64, 4, 3>, 2, 3, true>("float_flat_512x1024_p4_r2c3i"));
}
}  // namespace ssfft
