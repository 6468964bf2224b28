// ticket-queue four-step kernels (flat.cuh), fp32, lengths 3 * 2^k above the single-pass kernels (12288 ... 786432):
// N = N1 x N2 with a power-of-two column stage (its twiddles are powers of one number per butterfly, flat.cuh) and a row
// stage of length 3 * 2^j whose radices carry the factor 3 (pass tables in shared memory or L1).  These lengths are half
// of the reference's benchmark set (benchmark/benchmark.h:27-52); before, they ran radix pass + inner plan + interleave
// (composite.cuh, 15-25 % of the roofline) and the pass interpreter before that (7-9 %).
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f32_h(std::vector<FlatEntry> &v) {
    using A128 = TileCfg<float, 128, 16, 8, 1, 8, 32, 3>;
    using A256 = TileCfg<float, 256, 16, 16, 1, 16, 16, 3>;
    using A512 = TileCfg<float, 512, 8, 8, 8, 32, 8, 3>;
    using A1024 = TileCfg<float, 1024, 4, 16, 16, 64, 4, 3>;
    using B96 = TileCfg<float, 96, 4, 4, 6, 8, 32, 3>;      // 12 points per thread
    using B192 = TileCfg<float, 192, 4, 4, 12, 16, 16, 3>;  // 12 points per thread
    using B384 = TileCfg<float, 384, 8, 6, 8, 16, 16, 2>;   // 24 points per thread: 2 CTAs/SM
    using B768 = TileCfg<float, 768, 8, 12, 8, 32, 8, 2>;   // 24 points per thread: 2 CTAs/SM
    v.push_back(make_flat_entry<A128, B96, 2, 3, true, 3>("float_flat_128x96_r2c3i"));      // 12288
    v.push_back(make_flat_entry<A128, B192, 2, 3, true, 3>("float_flat_128x192_r2c3i"));    // 24576
    v.push_back(make_flat_entry<A256, B192, 2, 3, true, 3>("float_flat_256x192_r2c3i"));    // 49152
    v.push_back(make_flat_entry<A256, B384, 2, 2, true, 3>("float_flat_256x384_r2c2i"));    // 98304
    v.push_back(make_flat_entry<A256, B768, 2, 2, true, 3>("float_flat_256x768_r2c2i"));    // 196608
    v.push_back(make_flat_entry<A512, B768, 2, 2, true, 3>("float_flat_512x768_r2c2i"));    // 393216
    v.push_back(make_flat_entry<A1024, B768, 2, 2, true, 3>("float_flat_1024x768_r2c2i"));  // 786432
}
}  // namespace ssfft
