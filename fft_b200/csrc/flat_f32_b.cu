// ticket-queue four-step kernels (flat.cuh), fp32, 2^15
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_b(std::vector<FlatEntry> &v) {
    v.push_back(make_flat_entry<TileCfg<float, 128, 16, 8, 1, 8, 32, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, 1, 3>("float_flat_128x256_r1c3"));
    v.push_back(make_flat_entry<TileCfg<float, 128, 16, 8, 1, 8, 32, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, 2, 2>("float_flat_128x256_r2c2"));
}
}  // namespace ssfft
