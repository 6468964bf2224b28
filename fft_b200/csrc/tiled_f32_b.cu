// four-step tile kernels, fp32, lengths 512 and 1024
#include "tiled_launch.cuh"
namespace ssfft {
void register_tile_f32_b(std::vector<TileEntry> &v) {
    v.push_back(SSFFT_TILE(float, 512, 32, 16, 1, 16, 16, 2));
    v.push_back(SSFFT_TILE(float, 1024, 32, 32, 1, 32, 8, 2));
}
}  // namespace ssfft
