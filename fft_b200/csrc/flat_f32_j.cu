// ticket-queue four-step kernels (flat.cuh), fp32, 2^21, 2^22, 3 * 2^19, 3 * 2^20, 9 * 2^17, 9 * 2^18: a 2048-point leg with 32 points per thread
// (64 registers of data, ring of one in-place slot, 2 CTAs/SM).  These lengths otherwise run the composite plan
// (radix pass + inner plan + interleave, three trips through HBM: 22-25 % of the roofline).
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_j(std::vector<FlatEntry> &v) {
    using A768 = TileCfg<float, 768, 8, 8, 12, 32, 8, 2>;
    using A1024 = TileCfg<float, 1024, 4, 16, 16, 64, 4, 2>;
    using L2048 = TileCfg<float, 2048, 8, 16, 16, 64, 4, 2>;
    v.push_back(make_flat_entry<A1024, L2048, 1, 2, true, 0>("float_flat_1024x2048_r1c2i"));  // 2^21
    v.push_back(make_flat_entry<L2048, L2048, 1, 2, true, 0>("float_flat_2048x2048_r1c2i"));  // 2^22
    v.push_back(make_flat_entry<A768, L2048, 1, 2, true, 0>("float_flat_768x2048_r1c2i"));    // 3 * 2^19
    // a 1536-point leg (radix 8 x 8 x 24, 24 points per thread): 9 * 2^17, 9 * 2^18, 3 * 2^20
    using L1536 = TileCfg<float, 1536, 8, 8, 24, 64, 4, 2>;
    v.push_back(make_flat_entry<A768, L1536, 1, 2, true, 0>("float_flat_768x1536_r1c2i"));    // 9 * 2^17
    v.push_back(make_flat_entry<L1536, L1536, 1, 2, true, 0>("float_flat_1536x1536_r1c2i"));  // 9 * 2^18
    v.push_back(make_flat_entry<L1536, L2048, 1, 2, true, 0>("float_flat_1536x2048_r1c2i"));  // 3 * 2^20
}
}  // namespace ssfft
