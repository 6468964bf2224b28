"""Python mirror of the reference's class API over the C ABI (include/ssfft.h).

Same names and argument meaning as signalsmith::FFT<V> (signalsmith-fft.h:326-387), RealFFT<V>
(:402-503) and ModifiedRealFFT<V> (:505-508): ``setSize``, ``setSizeMinimum``, ``setSizeMaximum``,
``size``, ``fft(input, output)``, ``ifft(input, output)``, static ``sizeMinimum`` / ``sizeMaximum``.
Both directions are unnormalised; RealFFT packs (DC, Nyquist) into bin 0.

``fft`` / ``ifft`` accept
  * CUDA torch tensors (batched device path: ``[..., N]`` contiguous, any leading batch shape), or
  * numpy arrays / CPU tensors (host path: staged through the device by ssfft_exec_host).
There is no CPU implementation behind this module; without libssfft.so or a GPU it raises.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib as L

try:  # torch is plumbing (device memory, streams); the library itself does not depend on it
    import torch
except Exception:  # pragma: no cover
    torch = None

_REAL = {"float32": L.SSFFT_F32, "float64": L.SSFFT_F64, "float": L.SSFFT_F32, "double": L.SSFFT_F64}


def _prec_of(dtype) -> int:
    if isinstance(dtype, str):
        return _REAL[dtype]
    if torch is not None and isinstance(dtype, torch.dtype):
        return {torch.float32: L.SSFFT_F32, torch.float64: L.SSFFT_F64,
                torch.complex64: L.SSFFT_F32, torch.complex128: L.SSFFT_F64}[dtype]
    dt = np.dtype(dtype)
    return {np.dtype(np.float32): L.SSFFT_F32, np.dtype(np.float64): L.SSFFT_F64,
            np.dtype(np.complex64): L.SSFFT_F32, np.dtype(np.complex128): L.SSFFT_F64}[dt]


def _is_cuda_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


class _PlanOwner:
    _kind = L.SSFFT_C2C

    def __init__(self, dtype, device=None):
        self._lib = L.load()
        self._prec = _prec_of(dtype)
        self._device = -1 if device is None else int(device)
        self._plan = ctypes.c_void_p()
        self._n = None

    # -- plan management ------------------------------------------------------------------
    def _replan(self, n: int) -> None:
        self._release()
        plan = ctypes.c_void_p()
        L.check(self._lib.ssfft_plan_create(ctypes.byref(plan), self._kind, self._prec, n, self._device),
                f"ssfft_plan_create(n={n})")
        self._plan = plan
        self._n = n

    def _release(self) -> None:
        if getattr(self, "_plan", None) is not None and self._plan.value:
            self._lib.ssfft_plan_destroy(self._plan)
            self._plan = ctypes.c_void_p()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def describe(self) -> str:
        buf = ctypes.create_string_buffer(1024)
        L.check(self._lib.ssfft_plan_describe(self._plan, buf, 1024), "ssfft_plan_describe")
        return buf.value.decode()

    # -- dtype helpers ----------------------------------------------------------------------
    @property
    def _np_real(self):
        return np.float32 if self._prec == L.SSFFT_F32 else np.float64

    @property
    def _np_cplx(self):
        return np.complex64 if self._prec == L.SSFFT_F32 else np.complex128

    @property
    def _t_real(self):
        return torch.float32 if self._prec == L.SSFFT_F32 else torch.float64

    @property
    def _t_cplx(self):
        return torch.complex64 if self._prec == L.SSFFT_F32 else torch.complex128

    def _stream(self, t):
        return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)

    def _check_dev(self, t, dtype, last, what):
        if t.dtype != dtype:
            raise TypeError(f"{what}: expected {dtype}, got {t.dtype}")
        if not t.is_contiguous():
            raise ValueError(f"{what} must be contiguous")
        if t.shape[-1] != last:
            raise ValueError(f"{what}: last dimension must be {last}, got {t.shape[-1]}")

    # -- extended execution (ssfft_exec_*_ex): explicit layouts + multipliers fused into the first load / last store --
    def _io(self, in_real, out_real, in_stride, in_dist, out_stride, out_dist, pre, post, pre_dist, post_dist):
        """Build a struct ssfft_io.  pre / post are CUDA tensors: a real dtype is a window (SSFFT_MUL_REAL), a complex
        dtype a filter (SSFFT_MUL_COMPLEX).  The tensors must stay alive until the call has been enqueued."""
        io = L.SsfftIo()
        io.in_stride, io.in_dist, io.out_stride, io.out_dist = int(in_stride), int(in_dist), int(out_stride), int(out_dist)
        io.pre_dist, io.post_dist = int(pre_dist), int(post_dist)
        for name, t in (("pre", pre), ("post", post)):
            if t is None:
                setattr(io, name, None)
                setattr(io, name + "_kind", L.SSFFT_MUL_NONE)
                continue
            if not (_is_cuda_tensor(t) and t.is_contiguous()):
                raise TypeError(f"{name} must be a contiguous CUDA tensor")
            if t.dtype not in (self._t_real, self._t_cplx):
                raise TypeError(f"{name}: expected {self._t_real} (window) or {self._t_cplx} (filter), got {t.dtype}")
            setattr(io, name, t.data_ptr())
            setattr(io, name + "_kind", L.SSFFT_MUL_COMPLEX if t.is_complex() else L.SSFFT_MUL_REAL)
        return io

    def _flat(self, t, dtype, what):
        if not _is_cuda_tensor(t):
            raise TypeError(f"{what} must be a CUDA tensor (the extended calls are device-only)")
        if t.dtype != dtype:
            raise TypeError(f"{what}: expected {dtype}, got {t.dtype}")
        if not t.is_contiguous():
            raise ValueError(f"{what} must be contiguous (the layout is given by the stride / dist arguments)")
        return t

    @staticmethod
    def _span(batch, length, stride, dist):
        return 0 if batch == 0 or length == 0 else (batch - 1) * (dist or length) + (length - 1) * (stride or 1) + 1

    def _host(self, op, x, out, in_dt, in_last, out_dt, out_last):
        if torch is not None and isinstance(x, torch.Tensor):
            x = x.numpy()
        if torch is not None and isinstance(out, torch.Tensor):
            out_np = out.numpy()
        else:
            out_np = out
        x = np.ascontiguousarray(x, dtype=in_dt)
        if x.shape[-1] != in_last:
            raise ValueError(f"input: last dimension must be {in_last}, got {x.shape[-1]}")
        batch = x.size // in_last if in_last else 0
        if not (isinstance(out_np, np.ndarray) and out_np.dtype == out_dt and out_np.flags.c_contiguous):
            raise TypeError(f"output must be a C-contiguous numpy array of {np.dtype(out_dt)}")
        # like the reference, only the first out_last entries of each output row are written
        if out_np.shape[-1] != out_last or out_np.size != batch * out_last:
            raise ValueError(f"output: expected shape [..., {out_last}] with {batch} transforms")
        L.check(self._lib.ssfft_exec_host(self._plan, op, x.ctypes.data, out_np.ctypes.data, batch), "ssfft_exec_host")
        return out


class FFT(_PlanOwner):
    """signalsmith::FFT<V> -- complex transform of any length (signalsmith-fft.h:69-387)."""
    _kind = L.SSFFT_C2C

    def __init__(self, size: int, fastDirection: int = 0, dtype="float32", device=None):
        super().__init__(dtype, device)
        if fastDirection > 0:
            size = self.sizeMinimum(size)
        if fastDirection < 0:
            size = self.sizeMaximum(size)
        self._size = None
        self.setSize(size)

    @staticmethod
    def sizeMinimum(size: int) -> int:
        return int(L.load().ssfft_size_minimum(size))

    @staticmethod
    def sizeMaximum(size: int) -> int:
        return int(L.load().ssfft_size_maximum(size))

    def setSize(self, size: int) -> int:
        if size != self._size:  # re-plans only on change (:356-363)
            self._replan(size)
            self._size = size
        return self._size

    def setSizeMinimum(self, size: int) -> int:
        return self.setSize(self.sizeMinimum(size))

    def setSizeMaximum(self, size: int) -> int:
        return self.setSize(self.sizeMaximum(size))

    def size(self) -> int:
        return self._size

    def _run(self, x, out, direction):
        n = self._size
        if _is_cuda_tensor(x):
            self._check_dev(x, self._t_cplx, n, "input")
            self._check_dev(out, self._t_cplx, n, "output")
            batch = x.numel() // n if n else 0
            if out.numel() != x.numel() or out.device != x.device:
                raise ValueError("output must match the input's shape and device")
            L.check(self._lib.ssfft_exec_c2c(self._plan, x.data_ptr(), out.data_ptr(), batch, direction,
                                             self._stream(x)), "ssfft_exec_c2c")
            return out
        return self._host(0 if direction == L.SSFFT_FORWARD else 1, x, out, self._np_cplx, n, self._np_cplx, n)

    def fft(self, input, output):
        return self._run(input, output, L.SSFFT_FORWARD)

    def ifft(self, input, output):
        return self._run(input, output, L.SSFFT_INVERSE)

    def _run_ex(self, x, out, batch, direction, in_stride, in_dist, out_stride, out_dist, pre, post, pre_dist, post_dist):
        n = self._size
        self._flat(x, self._t_cplx, "input")
        self._flat(out, self._t_cplx, "output")
        if x.numel() < self._span(batch, n, in_stride, in_dist) or out.numel() < self._span(batch, n, out_stride, out_dist):
            raise ValueError("buffer too small for the requested layout")
        io = self._io(False, False, in_stride, in_dist, out_stride, out_dist, pre, post, pre_dist, post_dist)
        L.check(self._lib.ssfft_exec_c2c_ex(self._plan, x.data_ptr(), out.data_ptr(), batch, direction, ctypes.byref(io),
                                            self._stream(x)), "ssfft_exec_c2c_ex")
        return out

    def fft_ex(self, input, output, batch, *, in_stride=1, in_dist=0, out_stride=1, out_dist=0, pre=None, post=None,
               pre_dist=0, post_dist=0):
        """Forward transforms with explicit layouts: sample e of transform b is read at ``input[b*in_dist + e*in_stride]``
        (times ``pre``) and bin k written to ``output[b*out_dist + k*out_stride]`` (times ``post``); ssfft_exec_c2c_ex.
        ``in_stride=cols, in_dist=1`` walks the columns of a row-major matrix (second pass of a 2-D transform)."""
        return self._run_ex(input, output, batch, L.SSFFT_FORWARD, in_stride, in_dist, out_stride, out_dist, pre, post,
                            pre_dist, post_dist)

    def ifft_ex(self, input, output, batch, *, in_stride=1, in_dist=0, out_stride=1, out_dist=0, pre=None, post=None,
                pre_dist=0, post_dist=0):
        """Inverse counterpart of :meth:`fft_ex`; ``pre`` = a complex filter gives ``ifft(spectrum * filter)`` in one pass."""
        return self._run_ex(input, output, batch, L.SSFFT_INVERSE, in_stride, in_dist, out_stride, out_dist, pre, post,
                            pre_dist, post_dist)


class FFT2:
    """Two-dimensional complex transform of row-major ``[..., rows, cols]`` CUDA tensors: a batched row pass (contiguous,
    the plain kernels) followed by a column pass through the extended call (``in_stride = cols, in_dist = 1``, in place
    on the result of the row pass) -- two launches per matrix batch for fused sizes, no transposed copy of the data in
    HBM.  Unnormalised in both directions, like the 1-D transforms: ``ifft2(fft2(x)) == rows * cols * x``.
    (The reference has no 2-D call; its users loop over rows and columns on the host.)"""

    def __init__(self, rows: int, cols: int, dtype="float32", device=None):
        self.rows, self.cols = int(rows), int(cols)
        self._row_pass = FFT(self.cols, dtype=dtype, device=device)   # length-cols transforms along each row
        self._col_pass = FFT(self.rows, dtype=dtype, device=device)   # length-rows transforms down each column

    def _run(self, x, out, inverse):
        rp, cp = self._row_pass, self._col_pass
        rp._flat(x, rp._t_cplx, "input")
        rp._flat(out, rp._t_cplx, "output")
        if x.shape[-2:] != (self.rows, self.cols) or out.shape != x.shape or out.device != x.device:
            raise ValueError(f"expected [..., {self.rows}, {self.cols}] tensors of the same shape and device")
        (rp.ifft if inverse else rp.fft)(x, out)
        per = self.rows * self.cols
        flat = out.reshape(-1)
        run = cp.ifft_ex if inverse else cp.fft_ex
        for m in range(x.numel() // per if per else 0):   # one column pass per matrix of the batch
            v = flat[m * per:(m + 1) * per]
            run(v, v, self.cols, in_stride=self.cols, in_dist=1, out_stride=self.cols, out_dist=1)
        return out

    def fft2(self, input, output):
        return self._run(input, output, False)

    def ifft2(self, input, output):
        return self._run(input, output, True)


class RealFFT(_PlanOwner):
    """signalsmith::RealFFT<V> -- even-length real transform, (DC, Nyquist) packed into bin 0 (:393-503)."""
    _kind = L.SSFFT_REAL

    def __init__(self, size: int, fastDirection: int = 0, dtype="float32", device=None):
        super().__init__(dtype, device)
        if fastDirection > 0:
            size = self.sizeMinimum(size)
        if fastDirection < 0:
            size = self.sizeMaximum(size)
        self.setSize(size)

    @staticmethod
    def sizeMinimum(size: int) -> int:
        return int(L.load().ssfft_real_size_minimum(size))

    @staticmethod
    def sizeMaximum(size: int) -> int:
        return int(L.load().ssfft_real_size_maximum(size))

    def setSize(self, size: int) -> int:
        # the reference rebuilds unconditionally and returns the COMPLEX size N/2 (:416-435, quirk kept)
        self._replan(size)
        self._half = size // 2
        return self._half

    def setSizeMinimum(self, size: int) -> int:
        return self.setSize(self.sizeMinimum(size))

    def setSizeMaximum(self, size: int) -> int:
        return self.setSize(self.sizeMaximum(size))

    def size(self) -> int:
        return self._half * 2  # :442-444

    def fft(self, input, output):
        n, h = self.size(), self._half
        if _is_cuda_tensor(input):
            self._check_dev(input, self._t_real, n, "input")
            self._check_dev(output, self._t_cplx, h, "output")
            batch = input.numel() // n if n else 0
            if output.numel() != batch * h:
                raise ValueError("output must hold N/2 complex bins per transform")
            L.check(self._lib.ssfft_exec_r2c(self._plan, input.data_ptr(), output.data_ptr(), batch,
                                             self._stream(input)), "ssfft_exec_r2c")
            return output
        return self._host(2, input, output, self._np_real, n, self._np_cplx, h)

    def ifft(self, input, output):
        n, h = self.size(), self._half
        if _is_cuda_tensor(input):
            self._check_dev(input, self._t_cplx, h, "input")
            self._check_dev(output, self._t_real, n, "output")
            batch = input.numel() // h if h else 0
            if output.numel() != batch * n:
                raise ValueError("output must hold N reals per transform")
            L.check(self._lib.ssfft_exec_c2r(self._plan, input.data_ptr(), output.data_ptr(), batch,
                                             self._stream(input)), "ssfft_exec_c2r")
            return output
        return self._host(3, input, output, self._np_cplx, h, self._np_real, n)


    def fft_ex(self, input, output, batch, *, in_stride=1, in_dist=0, out_stride=1, out_dist=0, pre=None, post=None,
               pre_dist=0, post_dist=0):
        """RealFFT::fft with explicit layouts (ssfft_exec_r2c_ex).  Input side counts REALS: ``in_dist = hop < N`` reads
        overlapping frames straight out of a signal, ``pre`` = a real window of N samples is applied on load; the output
        side counts complex bins, ``post`` = a complex filter multiplies the packed half spectrum on store."""
        n, h = self.size(), self._half
        self._flat(input, self._t_real, "input")
        self._flat(output, self._t_cplx, "output")
        if input.numel() < self._span(batch, n, in_stride, in_dist) or output.numel() < self._span(batch, h, out_stride, out_dist):
            raise ValueError("buffer too small for the requested layout")
        io = self._io(True, False, in_stride, in_dist, out_stride, out_dist, pre, post, pre_dist, post_dist)
        L.check(self._lib.ssfft_exec_r2c_ex(self._plan, input.data_ptr(), output.data_ptr(), batch, ctypes.byref(io),
                                            self._stream(input)), "ssfft_exec_r2c_ex")
        return output

    def ifft_ex(self, input, output, batch, *, in_stride=1, in_dist=0, out_stride=1, out_dist=0, pre=None, post=None,
                pre_dist=0, post_dist=0):
        """RealFFT::ifft with explicit layouts (ssfft_exec_c2r_ex): ``pre`` = complex filter on the packed half spectrum,
        ``post`` = real synthesis window on the N output samples.  Outputs of different transforms must not overlap."""
        n, h = self.size(), self._half
        self._flat(input, self._t_cplx, "input")
        self._flat(output, self._t_real, "output")
        if input.numel() < self._span(batch, h, in_stride, in_dist) or output.numel() < self._span(batch, n, out_stride, out_dist):
            raise ValueError("buffer too small for the requested layout")
        io = self._io(False, True, in_stride, in_dist, out_stride, out_dist, pre, post, pre_dist, post_dist)
        L.check(self._lib.ssfft_exec_c2r_ex(self._plan, input.data_ptr(), output.data_ptr(), batch, ctypes.byref(io),
                                            self._stream(input)), "ssfft_exec_c2r_ex")
        return output

    def stft(self, signal, hop, window=None, output=None):
        """Short-time transform of a 1-D CUDA signal in ONE launch: frame b = ``signal[b*hop : b*hop + N] * window``,
        ``frames = (len - N) // hop + 1``; returns ``[frames, N/2]`` packed half spectra.  No frame matrix is ever
        written to HBM (overlapping frames are read through L2)."""
        n, h = self.size(), self._half
        self._flat(signal, self._t_real, "signal")
        if signal.dim() != 1 or signal.numel() < n or hop < 1:
            raise ValueError("signal must be 1-D with at least N samples, hop >= 1")
        frames = (signal.numel() - n) // hop + 1
        if output is None:
            output = torch.empty((frames, h), dtype=self._t_cplx, device=signal.device)
        return self.fft_ex(signal, output, frames, in_dist=hop, pre=window)


class ModifiedRealFFT(RealFFT):
    """signalsmith::ModifiedRealFFT<V> == RealFFT<V, FFTOptions::halfFreqShift> (:389-391, :505-508)."""
    _kind = L.SSFFT_REAL_MODIFIED


def launch_count() -> int:
    return int(L.load().ssfft_launch_count())


def fill_uniform(t, seed: int, first_idx: int = 0):
    """Fill a CUDA tensor with the shared synthetic distribution (uniform [-0.5, 0.5))."""
    lib = L.load()
    real = torch.view_as_real(t) if t.is_complex() else t
    prec = _prec_of(real.dtype)
    L.check(lib.ssfft_fill_uniform(real.data_ptr(), real.numel(), prec, seed, first_idx,
                                   ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)),
            "ssfft_fill_uniform")
    return t
