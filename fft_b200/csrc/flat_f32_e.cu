// ticket-queue four-step kernels (flat.cuh), fp32, WIDE tiles: 512 consumer threads, twice the lanes of flat_f32_a..d, so
// every global run is twice as long (tools/l2_ceiling.cu: 128-byte runs on both sides cap a two-pass transform at 63 % of
// the HBM roofline, 256-byte runs at 82 %).  Ring of one slot that doubles as the exchange buffer, 2 CTAs per SM.
// MEASURED SLOWER than the 256-thread tiles (profiles/flat_ab_r02k.txt: 2^16 46.9 vs 59.3 %, 2^17..2^20 -4..0 points; the
// RealFFT flavours 41 vs 48 %): at these sizes the kernels are bound by instruction issue, and two CTAs of 512 threads
// with one ring slot overlap their copies worse than three of 256 with two.  Kept selectable (SSFFT_FLAT_NAME=_w_) for
// 2^15 / 2^16 as the reference point of that experiment; the 2^17..2^20 entries were dropped.
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_e(std::vector<FlatEntry> &v) {
    v.push_back(make_flat_entry<TileCfg<float, 256, 16, 16, 1, 16, 32, 2>, TileCfg<float, 256, 16, 16, 1, 16, 32, 2>, 1, 2, true, 0>("float_flat_256x256_w_r1c2i"));
    v.push_back(make_flat_entry<TileCfg<float, 128, 16, 8, 1, 8, 64, 2>, TileCfg<float, 256, 16, 16, 1, 16, 32, 2>, 1, 2, true, 0>("float_flat_128x256_w_r1c2i"));
}
}  // namespace ssfft
