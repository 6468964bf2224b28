// test_ex_request.cpp -- every rule of ex_validate (fft_b200/csrc/ex_request.h) on the CPU: defaults, multiplier kinds,
// alignment, overlapping outputs, in-place requests.  Build: g++ -std=c++11 tests/host/test_ex_request.cpp -Iinclude
#include <cstdio>
#include <cstring>

#include "../../fft_b200/csrc/ex_request.h"

using namespace ssfft;
static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { ++failures; printf("FAIL line %d: %s\n", __LINE__, #cond); } } while (0)

int main() {
    alignas(16) static char bufA[4096], bufB[4096], tab[4096];
    const size_t N = 64, NR = 128, F32 = 8, F64 = 16;  // complex length, real length, sizeof(complex<float/double>)
    ssfft_io io;
    ExRequest x;
    auto fresh = [&]() { memset(&io, 0, sizeof(io)); x = ExRequest(); };

    // defaults: an all-zero descriptor is the plain contiguous call
    fresh();
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 5, bufA, bufB, x) == SSFFT_OK);
    CHECK(x.is == 1 && x.os == 1 && x.id == (long long)N && x.od == (long long)N && x.in_plain && x.out_plain && !x.packed);
    fresh();
    CHECK(ex_validate(SSFFT_REAL, N, NR, F32, EX_R2C, &io, 5, bufA, bufB, x) == SSFFT_OK);
    CHECK(x.in_real && !x.out_real && x.in_len == (long long)NR && x.out_len == (long long)N && x.id == (long long)NR && x.packed);
    fresh();
    CHECK(ex_validate(SSFFT_REAL_MODIFIED, N, NR, F32, EX_C2R, &io, 5, bufA, bufB, x) == SSFFT_OK);
    CHECK(!x.in_real && x.out_real && x.od == (long long)NR && !x.packed);  // no (DC, Nyquist) bin in a half-bin-shifted spectrum

    // STFT: overlapping INPUT frames are fine, a window makes the side non-plain
    fresh(); io.in_dist = 32; io.pre = tab; io.pre_kind = SSFFT_MUL_REAL;
    CHECK(ex_validate(SSFFT_REAL, N, NR, F32, EX_R2C, &io, 100, bufA, bufB, x) == SSFFT_OK && !x.in_plain && x.out_plain && x.id == 32);
    // a multiplier pointer without a valid kind, or a kind without a pointer
    fresh(); io.pre = tab; io.pre_kind = 7;
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 2, bufA, bufB, x) == SSFFT_ERR_INVALID);
    fresh(); io.post_kind = SSFFT_MUL_COMPLEX;  // no table: ignored
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 2, bufA, bufB, x) == SSFFT_OK && x.post_kind == SSFFT_MUL_NONE && x.out_plain);
    // complex multipliers cannot act on a real side
    fresh(); io.pre = tab; io.pre_kind = SSFFT_MUL_COMPLEX;
    CHECK(ex_validate(SSFFT_REAL, N, NR, F32, EX_R2C, &io, 2, bufA, bufB, x) == SSFFT_ERR_INVALID);
    fresh(); io.post = tab; io.post_kind = SSFFT_MUL_COMPLEX;
    CHECK(ex_validate(SSFFT_REAL, N, NR, F32, EX_C2R, &io, 2, bufA, bufB, x) == SSFFT_ERR_INVALID);
    CHECK(ex_validate(SSFFT_REAL, N, NR, F32, EX_R2C, &io, 2, bufA, bufB, x) == SSFFT_OK);  // ... but on the spectrum side they can
    // negative fields
    for (int f = 0; f < 6; ++f) {
        fresh();
        int64_t *fields[6] = {&io.in_stride, &io.in_dist, &io.out_stride, &io.out_dist, &io.pre_dist, &io.post_dist};
        *fields[f] = -1;
        CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 2, bufA, bufB, x) == SSFFT_ERR_INVALID);
    }
    // alignment: complex buffers / tables to a whole complex value, real ones to a scalar
    fresh();
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 2, bufA + 4, bufB, x) == SSFFT_ERR_INVALID);
    CHECK(ex_validate(SSFFT_C2C, N, 0, F64, EX_C2C, &io, 2, bufA + 8, bufB, x) == SSFFT_ERR_INVALID);
    CHECK(ex_validate(SSFFT_REAL, N, NR, F32, EX_R2C, &io, 2, bufA + 4, bufB, x) == SSFFT_OK);      // reals: 4-byte aligned is enough
    CHECK(ex_validate(SSFFT_REAL, N, NR, F32, EX_R2C, &io, 2, bufA + 4, bufB + 4, x) == SSFFT_ERR_INVALID);
    fresh(); io.pre = tab + 4; io.pre_kind = SSFFT_MUL_COMPLEX;
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 2, bufA, bufB, x) == SSFFT_ERR_INVALID);
    io.pre_kind = SSFFT_MUL_REAL;
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 2, bufA, bufB, x) == SSFFT_OK);
    // outputs of different transforms must not overlap: rows one after the other, or interleaved columns
    fresh(); io.out_dist = N / 2;
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 2, bufA, bufB, x) == SSFFT_ERR_INVALID);
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 1, bufA, bufB, x) == SSFFT_OK);  // a single transform cannot overlap itself
    fresh(); io.out_stride = 2; io.out_dist = 2 * N - 2;  // strided rows that are one element too close
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 3, bufA, bufB, x) == SSFFT_ERR_INVALID);
    io.out_dist = 2 * N - 1;
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 3, bufA, bufB, x) == SSFFT_OK);
    fresh(); io.out_stride = 10; io.out_dist = 1;  // columns of a matrix with 10 columns: up to 10 transforms
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 10, bufA, bufB, x) == SSFFT_OK);
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 11, bufA, bufB, x) == SSFFT_ERR_INVALID);
    // in place: identical byte layout, or (complex) identical element layout
    fresh();
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 4, bufA, bufA, x) == SSFFT_OK);
    CHECK(ex_validate(SSFFT_REAL, N, NR, F32, EX_R2C, &io, 4, bufA, bufA, x) == SSFFT_OK);  // N reals == N/2 complex values
    fresh(); io.in_stride = io.out_stride = 7; io.in_dist = io.out_dist = 1;
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 7, bufA, bufA, x) == SSFFT_OK);   // column pass of a 2-D transform
    io.out_stride = 1; io.out_dist = 0;
    CHECK(ex_validate(SSFFT_C2C, N, 0, F32, EX_C2C, &io, 7, bufA, bufA, x) == SSFFT_ERR_INVALID);
    fresh(); io.in_dist = 32;  // overlapping frames cannot be transformed in place
    CHECK(ex_validate(SSFFT_REAL, N, NR, F32, EX_R2C, &io, 4, bufA, bufA, x) == SSFFT_ERR_INVALID);
    fresh(); io.in_dist = NR + 2; io.out_dist = N + 1;  // padded rows with the same bytes per row
    CHECK(ex_validate(SSFFT_REAL, N, NR, F32, EX_R2C, &io, 4, bufA, bufA, x) == SSFFT_OK);
    io.out_dist = N + 2;
    CHECK(ex_validate(SSFFT_REAL, N, NR, F32, EX_R2C, &io, 4, bufA, bufA, x) == SSFFT_ERR_INVALID);

    printf(failures ? "EX-REQUEST-TESTS FAILED (%d)\n" : "EX-REQUEST-TESTS OK\n", failures);
    return failures ? 1 : 0;
}
