"""fft_b200 -- B200-native (sm_100a) implementation of the Signalsmith FFT transform path.

The product is libssfft.so (hand-written CUDA kernels behind the C ABI in include/ssfft.h) and the
header-only C++ front end include/signalsmith-fft.h.  This package is the Python mirror of the same
class API, used by the tests and bench.py.  Importing it does not load the library; the first use
does, and fails loudly if the extension has not been built (no CPU fallback).
"""
from .api import FFT, FFT2, ModifiedRealFFT, RealFFT, fill_uniform, launch_count  # noqa: F401
from ._lib import LIB_PATH, SsfftError, build, declared_symbols, load  # noqa: F401

__all__ = ["FFT", "FFT2", "RealFFT", "ModifiedRealFFT", "fill_uniform", "launch_count", "build", "load", "SsfftError"]
