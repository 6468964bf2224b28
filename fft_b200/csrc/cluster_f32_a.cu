// cluster-resident four-step kernels (cluster.cuh), fp32: complex 2^14 and 2^15
#include "cluster_launch.cuh"
namespace ssfft {
void register_cluster_f32_a(std::vector<ClusterEntry> &v) {
    // kinds: bit 0 C2C, bit 1 R2C, bit 2 C2R.  First mask = what the planner uses by default (only where the kernel
    // measured faster than the alternatives on B200, profiles/README.md), second = everything it can do
    // (SSFFT_DSMEM_ALL=1: parity tests and experiments).
    v.push_back(make_cluster_entry<ClusterCfg<float, 128, 16, 8, 128, 16, 8, 4, 3>>("float_dsmem_128x128_c4", 0u, 7u));
    v.push_back(make_cluster_entry<ClusterCfg<float, 256, 16, 16, 128, 16, 8, 8, 3>>("float_dsmem_256x128_c8", 1u, 3u));
    v.push_back(make_cluster_entry<ClusterCfg<float, 128, 16, 8, 256, 16, 16, 8, 3>>("float_dsmem_128x256_c8", 0u, 4u));
}
}  // namespace ssfft
