// ticket-queue four-step kernels (flat.cuh), fp32, 2^15.  The first entry of a size is its default (measured, profiles/
// flat_ab_r02e.txt / _r02f.txt); the other is selected with SSFFT_FLAT_VARIANT="ring,ctas_per_sm,inplace".  Entries with a
// separate exchange buffer also carry the RealFFT kernels of length 2 N1 N2.
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_b(std::vector<FlatEntry> &v) {
    v.push_back(make_flat_entry<TileCfg<float, 128, 16, 8, 1, 8, 32, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, 1, 3, false, 0>("float_flat_128x256_r1c3x"));
    v.push_back(make_flat_entry<TileCfg<float, 128, 16, 8, 1, 8, 32, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, 2, 3, true, 1>("float_flat_128x256_r2c3i"));   // RealFFT forward (48.5 % vs 47.9 / 46.2, profiles/flat_ab_real_r02l.txt)
    v.push_back(make_flat_entry<TileCfg<float, 128, 16, 8, 1, 8, 32, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, 1, 3, true, 2>("float_flat_128x256_r1c3i"));   // RealFFT inverse: the three boxes of its column tiles fit 3 CTAs/SM only in place (45.4 % vs 42.7)
    // (a ring of three slots at 2 CTAs/SM was measured too: RealFFT 65536 forward 46.4 %, inverse 35.6 % -- profiles/flat_ab_real_r02ac.txt)
    // 2^14 = 128 x 128 with 32-lane tiles on both sides (256-byte runs): only used when SSFFT_FLAT_MIN_LOG2=14 -- the
    // single-pass kernel of 16384 is the default (A/B: profiles/flat_ab_16384_r02s.txt)
    v.push_back(make_flat_entry<TileCfg<float, 128, 16, 8, 1, 8, 32, 3>, TileCfg<float, 128, 16, 8, 1, 8, 32, 3>, 2, 3, true, 3>("float_flat_128x128_r2c3i"));
}
}  // namespace ssfft
