// ticket-queue four-step kernels (flat.cuh), fp64, 2^14 ... 2^18: the same kernel as fp32 with tiles of half as many
// lanes (16-byte elements: [128][16], [256][8], [512][4] = 32 KB) and 8 points per thread (32 registers of data).
// TMA boxes count 8-byte words, a double-precision complex value is two of them (tma_host.cuh).
#include "flat_launch.cuh"
namespace ssfft {
void register_flat_f64_a(std::vector<FlatEntry> &v) {
    using D128 = TileCfg<double, 128, 8, 4, 4, 16, 16, 2>;
    using D256 = TileCfg<double, 256, 8, 8, 4, 32, 8, 2>;
    using D512 = TileCfg<double, 512, 8, 8, 8, 64, 4, 2>;
    v.push_back(make_flat_entry<D128, D128, 2, 2, true, 3>("double_flat_128x128_r2c2i"));  // 2^14
    v.push_back(make_flat_entry<D128, D256, 2, 2, true, 3>("double_flat_128x256_r2c2i"));  // 2^15
    v.push_back(make_flat_entry<D256, D256, 2, 2, true, 3>("double_flat_256x256_r2c2i"));  // 2^16
    v.push_back(make_flat_entry<D256, D512, 2, 2, true, 3>("double_flat_256x512_r2c2i"));  // 2^17
    v.push_back(make_flat_entry<D512, D512, 2, 2, true, 3>("double_flat_512x512_r2c2i"));  // 2^18
    // 2^19, 2^20: the 1024-point leg holds 16 points per thread (64 registers of data): ring of one in-place slot
    using D1024 = TileCfg<double, 1024, 4, 16, 16, 64, 4, 2>;
    v.push_back(make_flat_entry<D512, D1024, 1, 2, true, 3>("double_flat_512x1024_r1c2i"));    // 2^19
    v.push_back(make_flat_entry<D1024, D1024, 1, 2, true, 3>("double_flat_1024x1024_r1c2i"));  // 2^20
    // 3 * 2^k and 9 * 2^k as far as 12 points per thread reach (the 384- and 768-point legs would need 24)
    using D96 = TileCfg<double, 96, 4, 4, 6, 8, 32, 2>;
    using D192 = TileCfg<double, 192, 4, 4, 12, 16, 16, 2>;
    v.push_back(make_flat_entry<D128, D96, 2, 2, true, 0>("double_flat_128x96_r2c2i"));    // 12288
    v.push_back(make_flat_entry<D128, D192, 2, 2, true, 0>("double_flat_128x192_r2c2i"));  // 24576
    v.push_back(make_flat_entry<D256, D192, 2, 2, true, 0>("double_flat_256x192_r2c2i"));  // 49152
    v.push_back(make_flat_entry<D96, D192, 2, 2, true, 0>("double_flat_96x192_r2c2i"));    // 18432
    v.push_back(make_flat_entry<D192, D192, 2, 2, true, 0>("double_flat_192x192_r2c2i"));  // 36864
    v.push_back(make_flat_entry<D512, D192, 2, 2, true, 0>("double_flat_512x192_r2c2i"));  // 98304
    v.push_back(make_flat_entry<D1024, D192, 1, 2, true, 0>("double_flat_1024x192_r1c2i"));  // 196608
}
}  // namespace ssfft
