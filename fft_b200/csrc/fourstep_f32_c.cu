// cluster four-step kernels (both stages in one persistent launch), fp32
#include "tiled_launch.cuh"
namespace ssfft {
void register_fourstep_f32_c(std::vector<FourStepEntry> &v) {
    v.push_back(make_fourstep_entry<TileCfg<float, 256, 16, 16, 1, 16, 16, 3>, TileCfg<float, 512, 32, 16, 1, 16, 16, 2>>("float_cluster_256x512"));
    v.push_back(make_fourstep_entry<TileCfg<float, 512, 32, 16, 1, 16, 16, 2>, TileCfg<float, 512, 32, 16, 1, 16, 16, 2>>("float_cluster_512x512"));
}
}  // namespace ssfft
