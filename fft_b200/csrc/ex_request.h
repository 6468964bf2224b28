// ex_request.h -- validation of an extended-execution request (struct ssfft_io, include/ssfft.h): pure host logic, no
// CUDA, so that tests/host/test_ex_request.cpp can exercise every rule on the CPU.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/ssfft.h"

namespace ssfft {

enum { EX_C2C = 0, EX_R2C = 1, EX_C2R = 2 };

struct ExRequest {  // a validated ssfft_io with the defaults filled in; "elements" as in include/ssfft.h
    bool in_real = false, out_real = false;  // the side holds reals (R2C input, C2R output)
    long long in_len = 0, out_len = 0;       // elements per transform
    long long is = 1, id = 0, os = 1, od = 0;
    const void *pre = nullptr, *post = nullptr;
    int pre_kind = SSFFT_MUL_NONE, post_kind = SSFFT_MUL_NONE;
    long long pre_dist = 0, post_dist = 0;
    bool in_plain = true, out_plain = true;  // contiguous batch, no multiplier: nothing to do on that side
    bool packed = false;                     // the complex side is a RealFFT half spectrum (bin 0 = DC, Nyquist)
};

// plan_kind: SSFFT_C2C / SSFFT_REAL / SSFFT_REAL_MODIFIED; n: complex length; n_real: real length (real plans);
// elem: sizeof(complex<V>).  Returns SSFFT_OK or SSFFT_ERR_INVALID.
inline int ex_validate(int plan_kind, size_t n, size_t n_real, size_t elem, int op, const ssfft_io *io, long long batch,
                       const void *in, const void *out, ExRequest &x) {
    x.in_real = op == EX_R2C;
    x.out_real = op == EX_C2R;
    x.in_len = x.in_real ? (long long)n_real : (long long)n;
    x.out_len = x.out_real ? (long long)n_real : (long long)n;
    x.packed = plan_kind == SSFFT_REAL;  // the half-bin-shifted spectrum of a modified plan has no (DC, Nyquist) bin
    if (io->in_stride < 0 || io->in_dist < 0 || io->out_stride < 0 || io->out_dist < 0 || io->pre_dist < 0 || io->post_dist < 0)
        return SSFFT_ERR_INVALID;
    x.is = io->in_stride ? io->in_stride : 1;
    x.id = io->in_dist ? io->in_dist : x.in_len;
    x.os = io->out_stride ? io->out_stride : 1;
    x.od = io->out_dist ? io->out_dist : x.out_len;
    x.pre = io->pre; x.post = io->post;
    x.pre_kind = io->pre ? io->pre_kind : SSFFT_MUL_NONE;
    x.post_kind = io->post ? io->post_kind : SSFFT_MUL_NONE;
    x.pre_dist = io->pre_dist; x.post_dist = io->post_dist;
    if (io->pre && x.pre_kind != SSFFT_MUL_REAL && x.pre_kind != SSFFT_MUL_COMPLEX) return SSFFT_ERR_INVALID;
    if (io->post && x.post_kind != SSFFT_MUL_REAL && x.post_kind != SSFFT_MUL_COMPLEX) return SSFFT_ERR_INVALID;
    if ((x.in_real && x.pre_kind == SSFFT_MUL_COMPLEX) || (x.out_real && x.post_kind == SSFFT_MUL_COMPLEX)) return SSFFT_ERR_INVALID;
    // vector accesses: complex buffers and complex tables must be aligned to a whole complex value
    const size_t cplx = elem, real = elem / 2;
    if ((uintptr_t)in % (x.in_real ? real : cplx) || (uintptr_t)out % (x.out_real ? real : cplx)) return SSFFT_ERR_INVALID;
    if ((uintptr_t)x.pre % (x.pre_kind == SSFFT_MUL_COMPLEX ? cplx : real) || (uintptr_t)x.post % (x.post_kind == SSFFT_MUL_COMPLEX ? cplx : real))
        return SSFFT_ERR_INVALID;
    // the outputs of different transforms must not overlap: rows one after the other, or interleaved columns
    auto disjoint = [&](long long len, long long stride, long long dist) {
        return batch <= 1 || dist >= (len - 1) * stride + 1 || stride >= (batch - 1) * dist + 1;
    };
    if (!disjoint(x.out_len, x.os, x.od)) return SSFFT_ERR_INVALID;
    x.in_plain = x.is == 1 && x.id == x.in_len && !x.pre;
    x.out_plain = x.os == 1 && x.od == x.out_len && !x.post;
    if (in == out) {
        // in place: every transform is read completely before it is written, so it is enough that transform b's output
        // covers transform b's input bytes and nobody else's
        const long long in_sc = x.in_real ? 1 : 2, out_sc = x.out_real ? 1 : 2;
        const bool same_bytes = x.is == 1 && x.os == 1 && x.id * in_sc == x.od * out_sc && x.in_len * in_sc == x.out_len * out_sc;
        const bool same_layout = op == EX_C2C && x.is == x.os && x.id == x.od;
        if (!same_bytes && !same_layout) return SSFFT_ERR_INVALID;
    }
    return SSFFT_OK;
}

}  // namespace ssfft
