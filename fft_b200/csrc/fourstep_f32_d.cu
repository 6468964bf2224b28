// cluster four-step kernels (both stages in one persistent launch), fp32
#include "tiled_launch.cuh"
namespace ssfft {
void register_fourstep_f32_d(std::vector<FourStepEntry> &v) {
    v.push_back(make_fourstep_entry<TileCfg<float, 512, 32, 16, 1, 16, 16, 2>, TileCfg<float, 1024, 32, 32, 1, 32, 8, 2>>("float_cluster_512x1024"));
    v.push_back(make_fourstep_entry<TileCfg<float, 1024, 32, 32, 1, 32, 8, 2>, TileCfg<float, 1024, 32, 32, 1, 32, 8, 2>>("float_cluster_1024x1024"));
}
}  // namespace ssfft
