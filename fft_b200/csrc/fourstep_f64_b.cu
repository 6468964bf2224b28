// cluster four-step kernels, fp64, 2^18 .. 2^20
#include "tiled_launch.cuh"
namespace ssfft {
void register_fourstep_f64_b(std::vector<FourStepEntry> &v) {
    // 2^17: 256-point column tiles on 16 lanes x 32 threads (512 threads, like the 512-point row stage)
    v.push_back(make_fourstep_entry<TileCfg<double, 256, 8, 8, 4, 32, 16, 1>, TileCfg<double, 512, 8, 8, 8, 64, 8, 1>>("double_cluster_256x512"));
    v.push_back(make_fourstep_entry<TileCfg<double, 512, 8, 8, 8, 64, 8, 1>, TileCfg<double, 512, 8, 8, 8, 64, 8, 1>>("double_cluster_512x512"));
    v.push_back(make_fourstep_entry<TileCfg<double, 512, 8, 8, 8, 64, 8, 1>, TileCfg<double, 1024, 16, 8, 8, 64, 8, 1>>("double_cluster_512x1024"));
    v.push_back(make_fourstep_entry<TileCfg<double, 1024, 16, 8, 8, 64, 8, 1>, TileCfg<double, 1024, 16, 8, 8, 64, 8, 1>>("double_cluster_1024x1024"));
}
}  // namespace ssfft
