// cluster four-step kernels (both stages in one persistent launch), fp32
#include "tiled_launch.cuh"
namespace ssfft {
void register_fourstep_f32_a(std::vector<FourStepEntry> &v) {
    v.push_back(make_fourstep_entry<TileCfg<float, 128, 16, 8, 1, 8, 32, 3>, TileCfg<float, 128, 16, 8, 1, 8, 32, 3>>("float_cluster_128x128"));
    v.push_back(make_fourstep_entry<TileCfg<float, 128, 16, 8, 1, 8, 32, 3>, TileCfg<float, 256, 16, 16, 1, 16, 16, 3>>("float_cluster_128x256"));
}
}  // namespace ssfft
