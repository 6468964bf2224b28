// signalsmith-fft.h (B200 edition) -- header-only C++ front end over the C ABI in ssfft.h.
//
// Source-compatible with the reference header of the same name: namespace macro SIGNALSMITH_FFT_NAMESPACE,
// classes FFT<V>, RealFFT<V, flags>, ModifiedRealFFT<V>, FFTOptions, and the methods setSize /
// setSizeMinimum / setSizeMaximum / size / fft / ifft / static sizeMinimum / sizeMaximum with the same
// argument meaning and return types (reference: signalsmith-fft.h:326-387, :389-391, :402-503, :505-508).
// The transforms run on the GPU: every call forwards to libssfft.so (hand-written sm_100a kernels).
// There is no CPU fallback; failures (no device, CUDA error, unsupported size) throw std::runtime_error,
// the closest analogue of the std::bad_alloc the reference can throw from setSize.
//
// Two ways to call fft()/ifft():
//   1. the reference's way -- host containers or iterators:   fft.fft(input, output);
//      data is staged through the device (one transform), results are identical in layout and scaling;
//   2. NEW batched device-pointer overloads:                   fft.fft(d_in, d_out, batch, stream);
//      `batch` contiguous transforms, asynchronous on `stream` (a cudaStream_t passed as void*);
//   3. NEW extended overloads:                                 fft.fft(d_in, d_out, batch, io, stream);
//      strided / overlapping layouts and window / filter multipliers fused into the transform (struct ssfft_io).
//
// Conventions kept from the reference: unnormalised in both directions (ifft(fft(x)) == N*x); RealFFT
// packs DC into output[0].real() and Nyquist into output[0].imag() and writes only N/2 bins;
// RealFFT::setSize() returns N/2 while size() returns N; odd real sizes truncate to 2*(N/2).
#ifndef SIGNALSMITH_FFT_B200_V1
#define SIGNALSMITH_FFT_B200_V1
#ifndef SIGNALSMITH_FFT_NAMESPACE
#define SIGNALSMITH_FFT_NAMESPACE signalsmith
#endif

#include <complex>
#include <cstddef>
#include <iterator>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "ssfft.h"

namespace SIGNALSMITH_FFT_NAMESPACE {

namespace b200_detail {
template <typename V> struct Precision;
template <> struct Precision<float> { static constexpr int value = SSFFT_F32; };
template <> struct Precision<double> { static constexpr int value = SSFFT_F64; };

inline void check(int status, const char *what) {
    if (status != SSFFT_OK) {
        std::string msg = std::string(what) + ": " + ssfft_error_string(status);
        if (status == SSFFT_ERR_CUDA) msg += std::string(" [") + ssfft_last_cuda_error() + "]";
        throw std::runtime_error(msg);
    }
}
// Owner of one ssfft_plan.  A plan holds mutable per-call state next to its immutable tables (scratch for the large
// transforms, staging for the host path), so COPIES GET THEIR OWN PLAN, built from (kind, precision, size) -- reference
// objects are independent after a copy (they own std::vectors) and a pattern such as
// std::vector<FFT<float>>(nThreads, FFT<float>(n)) must not make the copies share scratch.  Moves transfer the handle.
class PlanHandle {
    ssfft_plan *p_ = nullptr;
    int kind_ = 0, precision_ = 0;
    std::size_t n_ = 0;

public:
    PlanHandle() = default;
    PlanHandle(int kind, int precision, std::size_t n) : kind_(kind), precision_(precision), n_(n) {
        check(ssfft_plan_create(&p_, kind, precision, n, -1), "ssfft_plan_create");
    }
    PlanHandle(const PlanHandle &o) : kind_(o.kind_), precision_(o.precision_), n_(o.n_) {
        if (o.p_) check(ssfft_plan_create(&p_, kind_, precision_, n_, -1), "ssfft_plan_create");
    }
    PlanHandle(PlanHandle &&o) noexcept : p_(o.p_), kind_(o.kind_), precision_(o.precision_), n_(o.n_) { o.p_ = nullptr; }
    PlanHandle &operator=(PlanHandle o) noexcept {
        std::swap(p_, o.p_); std::swap(kind_, o.kind_); std::swap(precision_, o.precision_); std::swap(n_, o.n_);
        return *this;
    }
    ~PlanHandle() { if (p_) ssfft_plan_destroy(p_); }
    ssfft_plan *get() const { return p_; }
    explicit operator bool() const { return p_ != nullptr; }
};
inline PlanHandle makePlan(int kind, int precision, std::size_t n) { return PlanHandle(kind, precision, n); }

// Accept containers (anything std::begin works on) or iterators/pointers, as the reference does (:56-67).
template <typename T, typename = void>
struct IteratorOf {
    static T get(const T &t) { return t; }
};
template <typename T>
struct IteratorOf<T, decltype((void)std::begin(std::declval<T &>()))> {
    static auto get(T &t) -> decltype(std::begin(t)) { return std::begin(t); }
};
template <typename T>
auto iteratorOf(T &&t) -> decltype(IteratorOf<typename std::remove_reference<T>::type>::get(t)) {
    return IteratorOf<typename std::remove_reference<T>::type>::get(t);
}
}  // namespace b200_detail

template <typename V>
class FFT {
    using complex = std::complex<V>;
    std::size_t _size;
    b200_detail::PlanHandle plan;
    std::vector<complex> hostIn, hostOut;  // staging for the host-iterator path

    template <bool inverse, typename InputIterator, typename OutputIterator>
    void runHost(InputIterator input, OutputIterator output) {
        for (std::size_t i = 0; i < _size; ++i) hostIn[i] = input[i];  // the input is never modified
        if (_size) b200_detail::check(ssfft_exec_host(plan.get(), inverse ? 1 : 0, hostIn.data(), hostOut.data(), 1), "ssfft_exec_host");
        for (std::size_t i = 0; i < _size; ++i) output[i] = hostOut[i];
    }

public:
    static std::size_t sizeMinimum(std::size_t size) { return ssfft_size_minimum(size); }
    static std::size_t sizeMaximum(std::size_t size) { return ssfft_size_maximum(size); }

    FFT(std::size_t size, int fastDirection = 0) : _size(0) {
        if (fastDirection > 0) size = sizeMinimum(size);
        if (fastDirection < 0) size = sizeMaximum(size);
        this->setSize(size);
    }

    std::size_t setSize(std::size_t size) {
        if (size != _size || !plan) {
            _size = size;
            hostIn.resize(size);
            hostOut.resize(size);
            plan = size ? b200_detail::makePlan(SSFFT_C2C, b200_detail::Precision<V>::value, size) : b200_detail::PlanHandle();
        }
        return _size;
    }
    std::size_t setSizeMinimum(std::size_t size) { return setSize(sizeMinimum(size)); }
    std::size_t setSizeMaximum(std::size_t size) { return setSize(sizeMaximum(size)); }
    const std::size_t &size() const { return _size; }

    // ---- the reference's host API (containers or iterators of std::complex<V>)
    template <typename Input, typename Output>
    void fft(Input &&input, Output &&output) {
        runHost<false>(b200_detail::iteratorOf(input), b200_detail::iteratorOf(output));
    }
    template <typename Input, typename Output>
    void ifft(Input &&input, Output &&output) {
        runHost<true>(b200_detail::iteratorOf(input), b200_detail::iteratorOf(output));
    }

    // ---- batched device-pointer overloads (new): `batch` contiguous transforms, async on `stream`
    void fft(const complex *d_in, complex *d_out, std::size_t batch, void *stream = nullptr) {
        if (_size) b200_detail::check(ssfft_exec_c2c(plan.get(), d_in, d_out, batch, SSFFT_FORWARD, stream), "ssfft_exec_c2c");
    }
    void ifft(const complex *d_in, complex *d_out, std::size_t batch, void *stream = nullptr) {
        if (_size) b200_detail::check(ssfft_exec_c2c(plan.get(), d_in, d_out, batch, SSFFT_INVERSE, stream), "ssfft_exec_c2c");
    }
    // ---- extended overloads (new): explicit layouts and fused multipliers, see struct ssfft_io in ssfft.h.
    // e.g. the column pass of a 2-D transform over a row-major [rows][cols] matrix, in place:
    //     ssfft_io io{}; io.in_stride = io.out_stride = cols; io.in_dist = io.out_dist = 1;  fft.fft(d, d, cols, io);
    void fft(const complex *d_in, complex *d_out, std::size_t batch, const ssfft_io &io, void *stream = nullptr) {
        if (_size) b200_detail::check(ssfft_exec_c2c_ex(plan.get(), d_in, d_out, batch, SSFFT_FORWARD, &io, stream), "ssfft_exec_c2c_ex");
    }
    void ifft(const complex *d_in, complex *d_out, std::size_t batch, const ssfft_io &io, void *stream = nullptr) {
        if (_size) b200_detail::check(ssfft_exec_c2c_ex(plan.get(), d_in, d_out, batch, SSFFT_INVERSE, &io, stream), "ssfft_exec_c2c_ex");
    }
    // batched HOST buffers in one call (H2D, kernels and D2H overlap inside the library)
    void fftHostBatch(const complex *h_in, complex *h_out, std::size_t batch) {
        if (_size) b200_detail::check(ssfft_exec_host(plan.get(), 0, h_in, h_out, batch), "ssfft_exec_host");
    }
    void ifftHostBatch(const complex *h_in, complex *h_out, std::size_t batch) {
        if (_size) b200_detail::check(ssfft_exec_host(plan.get(), 1, h_in, h_out, batch), "ssfft_exec_host");
    }
    std::string describe() const {
        char buf[1024] = "empty";
        if (plan) ssfft_plan_describe(plan.get(), buf, sizeof(buf));
        return buf;
    }
};

struct FFTOptions {
    static constexpr int halfFreqShift = 1;
};

template <typename V, int optionFlags = 0>
class RealFFT {
    static constexpr bool modified = (optionFlags & FFTOptions::halfFreqShift);
    using complex = std::complex<V>;
    std::size_t halfSize;
    b200_detail::PlanHandle plan;
    std::vector<V> hostReal;
    std::vector<complex> hostComplex;

public:
    // quirks of the reference reproduced on purpose (signalsmith-fft.h:403-408)
    static std::size_t sizeMinimum(std::size_t size) { return ssfft_real_size_minimum(size); }
    static std::size_t sizeMaximum(std::size_t size) { return ssfft_real_size_maximum(size); }

    RealFFT(std::size_t size, int fastDirection = 0) : halfSize(0) {
        if (fastDirection > 0) size = sizeMinimum(size);
        if (fastDirection < 0) size = sizeMaximum(size);
        this->setSize(size);
    }

    std::size_t setSize(std::size_t size) {
        halfSize = size / 2;
        hostReal.resize(halfSize * 2);
        hostComplex.resize(halfSize);
        plan = halfSize ? b200_detail::makePlan(modified ? SSFFT_REAL_MODIFIED : SSFFT_REAL,
                                                b200_detail::Precision<V>::value, halfSize * 2)
                        : b200_detail::PlanHandle();
        return halfSize;  // the reference returns the COMPLEX size here (:434)
    }
    std::size_t setSizeMinimum(std::size_t size) { return setSize(sizeMinimum(size)); }
    std::size_t setSizeMaximum(std::size_t size) { return setSize(sizeMaximum(size)); }
    std::size_t size() const { return halfSize * 2; }

    template <typename Input, typename Output>
    void fft(Input &&input, Output &&output) {
        auto in = b200_detail::iteratorOf(input);
        auto out = b200_detail::iteratorOf(output);
        for (std::size_t i = 0; i < halfSize * 2; ++i) hostReal[i] = in[i];
        if (halfSize) b200_detail::check(ssfft_exec_host(plan.get(), 2, hostReal.data(), hostComplex.data(), 1), "ssfft_exec_host");
        for (std::size_t i = 0; i < halfSize; ++i) out[i] = hostComplex[i];  // bins [N/2, N) stay untouched
    }
    template <typename Input, typename Output>
    void ifft(Input &&input, Output &&output) {
        auto in = b200_detail::iteratorOf(input);
        auto out = b200_detail::iteratorOf(output);
        for (std::size_t i = 0; i < halfSize; ++i) hostComplex[i] = in[i];
        if (halfSize) b200_detail::check(ssfft_exec_host(plan.get(), 3, hostComplex.data(), hostReal.data(), 1), "ssfft_exec_host");
        for (std::size_t i = 0; i < halfSize * 2; ++i) out[i] = hostReal[i];
    }

    // batched device-pointer overloads: N reals per transform <-> N/2 complex per transform
    void fft(const V *d_in, complex *d_out, std::size_t batch, void *stream = nullptr) {
        if (halfSize) b200_detail::check(ssfft_exec_r2c(plan.get(), d_in, d_out, batch, stream), "ssfft_exec_r2c");
    }
    void ifft(const complex *d_in, V *d_out, std::size_t batch, void *stream = nullptr) {
        if (halfSize) b200_detail::check(ssfft_exec_c2r(plan.get(), d_in, d_out, batch, stream), "ssfft_exec_c2r");
    }
    // extended overloads (new): e.g. an STFT straight out of a signal, window applied on load --
    //     ssfft_io io{}; io.in_dist = hop; io.pre = d_window; io.pre_kind = SSFFT_MUL_REAL;  rfft.fft(d_signal, d_spectra, frames, io);
    void fft(const V *d_in, complex *d_out, std::size_t batch, const ssfft_io &io, void *stream = nullptr) {
        if (halfSize) b200_detail::check(ssfft_exec_r2c_ex(plan.get(), d_in, d_out, batch, &io, stream), "ssfft_exec_r2c_ex");
    }
    void ifft(const complex *d_in, V *d_out, std::size_t batch, const ssfft_io &io, void *stream = nullptr) {
        if (halfSize) b200_detail::check(ssfft_exec_c2r_ex(plan.get(), d_in, d_out, batch, &io, stream), "ssfft_exec_c2r_ex");
    }
    std::string describe() const {
        char buf[1024] = "empty";
        if (plan) ssfft_plan_describe(plan.get(), buf, sizeof(buf));
        return buf;
    }
};

template <typename V>
struct ModifiedRealFFT : public RealFFT<V, FFTOptions::halfFreqShift> {
    using RealFFT<V, FFTOptions::halfFreqShift>::RealFFT;
};

// NEW (no counterpart in the reference): ONE complex transform sharded over several GPUs of this process.
// devices[r] owns block r of the natural order: shard pointers address size() / devices.size() elements in that GPU's
// memory.  The transform is the same unnormalised FFT<V> (fft then ifft scales by size()).  Calls are asynchronous on
// streams of the plan: synchronize() before reading the output.  transposedOutput skips the last exchange (output shard
// r = rows k1 in [r N1/P, (r+1) N1/P) of X[k1 + N1 k2]).  Move-only.
template <typename V>
class DistributedFFT {
    using complex = std::complex<V>;
    ssfft_dist_plan *plan = nullptr;
    std::size_t _size = 0;
    std::size_t _devices = 0;

    void run(complex *const *d_in, complex *const *d_out, int direction) {
        std::vector<void *> in(_devices), out(_devices);
        for (std::size_t r = 0; r < _devices; ++r) { in[r] = d_in[r]; out[r] = d_out[r]; }
        b200_detail::check(ssfft_dist_exec_c2c(plan, in.data(), out.data(), direction), "ssfft_dist_exec_c2c");
    }

public:
    DistributedFFT(std::size_t size, const std::vector<int> &devices, bool transposedOutput = false)
        : _size(size), _devices(devices.size()) {
        b200_detail::check(ssfft_dist_plan_create(&plan, b200_detail::Precision<V>::value, size, (int)devices.size(), devices.data(),
                                                  transposedOutput ? SSFFT_DIST_TRANSPOSED_OUTPUT : 0),
                           "ssfft_dist_plan_create");
    }
    DistributedFFT(const DistributedFFT &) = delete;
    DistributedFFT &operator=(const DistributedFFT &) = delete;
    DistributedFFT(DistributedFFT &&o) noexcept : plan(o.plan), _size(o._size), _devices(o._devices) { o.plan = nullptr; }
    ~DistributedFFT() { if (plan) ssfft_dist_plan_destroy(plan); }

    std::size_t size() const { return _size; }
    std::size_t devices() const { return _devices; }
    void fft(complex *const *d_in_shards, complex *const *d_out_shards) { run(d_in_shards, d_out_shards, SSFFT_FORWARD); }
    void ifft(complex *const *d_in_shards, complex *const *d_out_shards) { run(d_in_shards, d_out_shards, SSFFT_INVERSE); }
    void synchronize() { b200_detail::check(ssfft_dist_synchronize(plan), "ssfft_dist_synchronize"); }
    std::string describe() const {
        char buf[1024] = "empty";
        if (plan) ssfft_dist_plan_describe(plan, buf, sizeof(buf));
        return buf;
    }
};

}  // namespace SIGNALSMITH_FFT_NAMESPACE

#undef SIGNALSMITH_FFT_NAMESPACE
#endif  // SIGNALSMITH_FFT_B200_V1
