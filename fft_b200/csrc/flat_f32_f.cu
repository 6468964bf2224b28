// ticket-queue four-step kernels (flat.cuh), fp32, WIDE tiles (see flat_f32_e.cu), 2^17 and 2^18: the 512-point leg with 16 lanes
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_f(std::vector<FlatEntry> &v) {
    v.push_back(make_flat_entry<TileCfg<float, 256, 16, 16, 1, 16, 32, 2>, TileCfg<float, 512, 8, 8, 8, 32, 16, 2>, 1, 2, true, 3>("float_flat_256x512_w_r1c2i"));
    v.push_back(make_flat_entry<TileCfg<float, 512, 8, 8, 8, 32, 16, 2>, TileCfg<float, 512, 8, 8, 8, 32, 16, 2>, 1, 2, true, 3>("float_flat_512x512_w_r1c2i"));
}
}  // namespace ssfft
