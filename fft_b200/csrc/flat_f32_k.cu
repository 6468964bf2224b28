// ticket-queue four-step kernels (flat.cuh), fp32: an UNBALANCED split of 2^18 -- the long leg in the column stage, a
// 256-point row stage.  tools/l2_ceiling.cu: short runs cost more on the way out (transposed stores of the row stage) than on
// the way in (TMA boxes of the column stage); 1024 x 256 stores runs of 128 bytes where 512 x 512 stores 64.
// MEASURED (profiles/sweep_r02ak_unbalanced_float32.txt): 2^18 40.2 -> 45.5 % of the roofline.  The same idea at 2^17
// (512 x 256: 45.7 vs 46.5 %), 2^19 (2048 x 256: 37.2 vs 36.1 %) and 2^20 (2048 x 512: 29.7 vs 35.8 %) did not pay and those
// entries were dropped.  "_dflt" in the name makes the split the default of its length for complex plans.
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_k(std::vector<FlatEntry> &v) {
    using A1024 = TileCfg<float, 1024, 4, 16, 16, 64, 4, 3>;
    using B256 = TileCfg<float, 256, 16, 16, 1, 16, 16, 3>;
    v.push_back(make_flat_entry<A1024, B256, 2, 3, true, 0>("float_flat_1024x256_dflt_r2c3i"));  // 2^18
}
}  // namespace ssfft
