// ticket-queue four-step kernels (flat.cuh), fp32, WIDE tiles: 512 consumer threads, twice the lanes of flat_f32_a..d, so
// every global run is twice as long (tools/l2_ceiling.cu: 128-byte runs on both sides cap a two-pass transform at 63 % of
// the HBM roofline, 256-byte runs at 82 %).  Ring of one slot that doubles as the exchange buffer, 2 CTAs per SM.
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_e(std::vector<FlatEntry> &v) {
    v.push_back(make_flat_entry<TileCfg<float, 256, 16, 16, 1, 16, 32, 2>, TileCfg<float, 256, 16, 16, 1, 16, 32, 2>, 1, 2, true, 3>("float_flat_256x256_w_r1c2i"));
    v.push_back(make_flat_entry<TileCfg<float, 128, 16, 8, 1, 8, 64, 2>, TileCfg<float, 256, 16, 16, 1, 16, 32, 2>, 1, 2, true, 3>("float_flat_128x256_w_r1c2i"));
}
}  // namespace ssfft
