// ticket-queue four-step kernels (flat.cuh), fp32, lengths 9 * 2^k above the single-pass kernels (18432 ... 589824):
// N = (3 * 2^a) x (3 * 2^b), one factor 3 in each stage.  The column stage keeps its twiddles-as-powers for the passes
// before the last (powers of two) and carries the 3 in its LAST pass, which multiplies by a table row anyway; the row stage
// is the one of flat_f32_h.cu.  Before: radix pass + inner plan + interleave (composite.cuh), 19-21 % of the roofline.
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_i(std::vector<FlatEntry> &v) {
    using A96 = TileCfg<float, 96, 4, 4, 6, 8, 32, 3>;
    using A192 = TileCfg<float, 192, 4, 4, 12, 16, 16, 3>;
    using A384 = TileCfg<float, 384, 8, 4, 12, 16, 16, 2>;  // 24 points per thread: 2 CTAs/SM
    using A768 = TileCfg<float, 768, 8, 8, 12, 32, 8, 2>;
    using B192 = TileCfg<float, 192, 4, 4, 12, 16, 16, 3>;
    using B384 = TileCfg<float, 384, 8, 6, 8, 16, 16, 2>;
    using B768 = TileCfg<float, 768, 8, 12, 8, 32, 8, 2>;
    v.push_back(make_flat_entry<A96, B192, 2, 3, true, 3>("float_flat_96x192_r2c3i"));    // 18432
    v.push_back(make_flat_entry<A192, B192, 2, 3, true, 3>("float_flat_192x192_r2c3i"));  // 36864
    v.push_back(make_flat_entry<A192, B384, 2, 2, true, 3>("float_flat_192x384_r2c2i"));  // 73728
    v.push_back(make_flat_entry<A384, B384, 2, 2, true, 3>("float_flat_384x384_r2c2i"));  // 147456
    v.push_back(make_flat_entry<A384, B768, 2, 2, true, 3>("float_flat_384x768_r2c2i"));  // 294912
    v.push_back(make_flat_entry<A768, B768, 2, 2, true, 3>("float_flat_768x768_r2c2i"));  // 589824
}
}  // namespace ssfft
