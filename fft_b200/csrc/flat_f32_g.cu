// ticket-queue four-step kernels (flat.cuh), fp32, WIDE tiles (see flat_f32_e.cu), 2^19 and 2^20: the 1024-point leg with 8 lanes
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_g(std::vector<FlatEntry> &v) {
    v.push_back(make_flat_entry<TileCfg<float, 512, 8, 8, 8, 32, 16, 2>, TileCfg<float, 1024, 16, 8, 8, 64, 8, 2>, 1, 2, true, 3>("float_flat_512x1024_w_r1c2i"));
    v.push_back(make_flat_entry<TileCfg<float, 1024, 16, 8, 8, 64, 8, 2>, TileCfg<float, 1024, 16, 8, 8, 64, 8, 2>, 1, 2, true, 3>("float_flat_1024x1024_w_r1c2i"));
}
}  // namespace ssfft
