// ticket-queue four-step kernels (flat.cuh), fp32, 2^19 and 2^20 (1024-point leg, 4 lanes per tile).  The first entry of a
// size is its default; the others are selected with SSFFT_FLAT_NAME / SSFFT_FLAT_VARIANT.
// 1024-point leg: radix 4 x 16 x 16.  With 4 lanes per tile the first-pass stores of a radix-16 pass hit rows 16 apart =
// the same shared-memory banks (ncu r02c: 25 M conflicts on 47 M wavefronts, l1tex 86 % busy); rows 4 apart spread over
// all of them: 2^20 30.6 -> 34.5 %, 2^19 33.1 -> 34.2 % of the roofline (profiles/sweep_r02n_p4_f32.txt).
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_d(std::vector<FlatEntry> &v) {
    using L512 = TileCfg<float, 512, 8, 8, 8, 32, 8, 3>;
    using L1024 = TileCfg<float, 1024, 4, 16, 16, 64, 4, 3>;
    using L1024old = TileCfg<float, 1024, 16, 8, 8, 64, 4, 3>;
    v.push_back(make_flat_entry<L512, L1024, 1, 3, false, 3>("float_flat_512x1024_p4_r1c3x"));
    v.push_back(make_flat_entry<L512, L1024, 2, 3, true, 0>("float_flat_512x1024_p4_r2c3i"));
    v.push_back(make_flat_entry<L1024, L1024, 1, 3, false, 3>("float_flat_1024x1024_p4_r1c3x"));
    v.push_back(make_flat_entry<L1024, L1024, 2, 3, true, 0>("float_flat_1024x1024_p4_r2c3i"));
    v.push_back(make_flat_entry<L1024old, L1024old, 2, 3, true, 0>("float_flat_1024x1024_r2c3i"));  // the round-2 baseline of the A/B above
}
}  // namespace ssfft
