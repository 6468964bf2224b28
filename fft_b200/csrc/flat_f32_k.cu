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
    using B512 = TileCfg<float, 512, 8, 8, 8, 32, 8, 3>;
    v.push_back(make_flat_entry<A1024, B512, 2, 3, true, 0>("float_flat_1024x512_dflt_r2c3i"));  // 2^19: 36.1 -> 37.6 % (profiles/sweep_r02ao_2p19_float32.txt)
    // the same for 3 * 2^k / 9 * 2^k (profiles/sweep_r02am_unbalanced_3x9_float32.txt): only 294912 as 768 x 384 gained (41.3 -> 43.6 %);
    // 98304 as 384 x 256 (46.6 vs 51.2 %), 196608 as 768 x 256 (42.7 vs 44.3), 393216 as 1536 x 256 (38.9 vs 38.8), 589824 as
    // 1536 x 384 (37.3 vs 38.3) did not and were dropped
    using A768 = TileCfg<float, 768, 8, 8, 12, 32, 8, 2>;
    using B384 = TileCfg<float, 384, 8, 6, 8, 16, 16, 2>;
    v.push_back(make_flat_entry<A768, B384, 2, 2, true, 0>("float_flat_768x384_dflt_r2c2i"));  // 294912
}
}  // namespace ssfft
