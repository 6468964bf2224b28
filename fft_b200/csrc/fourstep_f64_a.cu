// cluster four-step kernels (both stages in one persistent launch), fp64: the pairs whose two tile kernels use the
// same CTA size (2^14, 2^16, 2^18, 2^19, 2^20); 2^15 and 2^17 keep two launches per chunk
#include "tiled_launch.cuh"
namespace ssfft {
void register_fourstep_f64_a(std::vector<FourStepEntry> &v) {
    v.push_back(make_fourstep_entry<TileCfg<double, 128, 8, 4, 4, 16, 8, 3>, TileCfg<double, 128, 8, 4, 4, 16, 8, 3>>("double_cluster_128x128"));
    v.push_back(make_fourstep_entry<TileCfg<double, 256, 8, 8, 4, 32, 8, 2>, TileCfg<double, 256, 8, 8, 4, 32, 8, 2>>("double_cluster_256x256"));
    // 2^15: the column stage runs the 128-point tile on 16 lanes x 16 threads so that both stages have 256 threads
    v.push_back(make_fourstep_entry<TileCfg<double, 128, 8, 4, 4, 16, 16, 2>, TileCfg<double, 256, 8, 8, 4, 32, 8, 2>>("double_cluster_128x256"));
}
}  // namespace ssfft
