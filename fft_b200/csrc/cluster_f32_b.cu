// cluster-resident four-step kernels (cluster.cuh), fp32: complex 2^16 (16-CTA clusters, non-portable size)
#include "cluster_launch.cuh"
namespace ssfft {
void register_cluster_f32_b(std::vector<ClusterEntry> &v) {
    v.push_back(make_cluster_entry<ClusterCfg<float, 256, 16, 16, 256, 16, 16, 16, 3>>("float_dsmem_256x256_c16", 0u, 7u));
}
}  // namespace ssfft
