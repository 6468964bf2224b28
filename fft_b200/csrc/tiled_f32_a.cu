// four-step tile kernels, fp32, lengths 64..256 (256 threads each so any two can share a cluster kernel)
#include "tiled_launch.cuh"
namespace ssfft {
void register_tile_f32_a(std::vector<TileEntry> &v) {
    v.push_back(SSFFT_TILE(float, 64, 8, 8, 1, 8, 32, 3));
    v.push_back(SSFFT_TILE(float, 128, 16, 8, 1, 8, 32, 3));
    v.push_back(SSFFT_TILE(float, 256, 16, 16, 1, 16, 16, 3));
}
}  // namespace ssfft
