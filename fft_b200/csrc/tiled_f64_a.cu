// four-step tile kernels, fp64 (8 lanes = 128-byte segments)
#include "tiled_launch.cuh"
namespace ssfft {
void register_tile_f64_a(std::vector<TileEntry> &v) {
    v.push_back(SSFFT_TILE(double, 64, 8, 8, 1, 8, 8, 4));
    v.push_back(SSFFT_TILE(double, 128, 8, 4, 4, 16, 8, 3));
    v.push_back(SSFFT_TILE(double, 256, 8, 8, 4, 32, 8, 2));
    v.push_back(SSFFT_TILE(double, 512, 8, 8, 8, 64, 8, 1));
    v.push_back(SSFFT_TILE(double, 1024, 16, 8, 8, 64, 8, 1));
}
}  // namespace ssfft
