"""CPU tests of host-side orchestration in the Python mirror, with the 1-D plans replaced by a torch.fft stand-in that
honours the extended calls' layout arguments (the real plans need a GPU; their kernels are covered by the emulation
tests and the -m gpu suite).  Checks what FFT2 asks of its row and column passes: slicing per matrix of the batch,
in-place column pass with stride = cols / dist = 1, unnormalised scaling in both directions."""
import pytest

torch = pytest.importorskip("torch")

from fft_b200 import api  # noqa: E402


class _Fake1D:
    _t_cplx = torch.complex128

    def __init__(self, n, dtype="float64", device=None):
        self.n = n
        self.calls = []

    def _flat(self, t, dt, what):
        assert t.dtype == dt and t.is_contiguous(), what
        return t

    def fft(self, x, out):
        out.copy_(torch.fft.fft(x, dim=-1))
        return out

    def ifft(self, x, out):
        out.copy_(torch.fft.ifft(x, dim=-1) * self.n)
        return out

    def _ex(self, x, out, batch, inverse, in_stride, in_dist, out_stride, out_dist):
        n = self.n
        self.calls.append((batch, in_stride, in_dist, out_stride, out_dist, x.data_ptr() == out.data_ptr()))
        xs = torch.stack([x[b * (in_dist or n): b * (in_dist or n) + (n - 1) * in_stride + 1: in_stride] for b in range(batch)])
        ys = torch.fft.ifft(xs, dim=-1) * n if inverse else torch.fft.fft(xs, dim=-1)
        for b in range(batch):
            out[b * (out_dist or n): b * (out_dist or n) + (n - 1) * out_stride + 1: out_stride] = ys[b]
        return out

    def fft_ex(self, x, out, batch, *, in_stride=1, in_dist=0, out_stride=1, out_dist=0, **_):
        return self._ex(x, out, batch, False, in_stride, in_dist, out_stride, out_dist)

    def ifft_ex(self, x, out, batch, *, in_stride=1, in_dist=0, out_stride=1, out_dist=0, **_):
        return self._ex(x, out, batch, True, in_stride, in_dist, out_stride, out_dist)


@pytest.mark.parametrize("rows,cols,batch", [(6, 10, 3), (16, 4, 1), (5, 7, 2)])
def test_fft2_orchestration(monkeypatch, rows, cols, batch):
    monkeypatch.setattr(api, "FFT", _Fake1D)
    f2 = api.FFT2(rows, cols, dtype="float64")
    x = torch.randn(batch, rows, cols, dtype=torch.complex128)
    x0 = x.clone()
    y = torch.empty_like(x)
    f2.fft2(x, y)
    assert torch.equal(x, x0), "input was changed"
    assert (y - torch.fft.fft2(x)).abs().max() < 1e-12
    # one in-place column pass per matrix: `cols` transforms of length `rows`, stride = cols, dist = 1 on both sides
    assert f2._col_pass.calls == [(cols, cols, 1, cols, 1, True)] * batch
    z = torch.empty_like(x)
    f2.ifft2(y, z)
    assert (z / (rows * cols) - x).abs().max() < 1e-12  # unnormalised both ways, like the 1-D transforms
    with pytest.raises(ValueError):
        f2.fft2(x, torch.empty(batch, rows, cols + 1, dtype=torch.complex128))


def test_span_of_a_layout():
    span = api._PlanOwner._span
    assert span(0, 64, 1, 0) == 0 and span(3, 0, 1, 0) == 0
    assert span(1, 64, 1, 0) == 64                      # one contiguous transform
    assert span(4, 64, 0, 0) == 256                     # defaults: stride 1, dist = length
    assert span(5, 64, 1, 16) == 4 * 16 + 64            # overlapping frames, hop 16
    assert span(10, 64, 10, 1) == 9 + 63 * 10 + 1       # the columns of a [64][10] matrix: exactly the matrix
    assert span(3, 64, 2, 200) == 2 * 200 + 63 * 2 + 1  # strided rows


def test_stft_frames_and_arguments():
    """RealFFT.stft: frames = (len - N) // hop + 1, hop as the input distance, the window as the pre-multiplier."""
    seen = {}

    class _FakeReal(api.RealFFT):
        def __init__(self, n):  # no plan, no library
            self._half = n // 2
            self._prec = api.L.SSFFT_F32

        def _flat(self, t, dt, what):
            assert t.dtype == dt
            return t

        def fft_ex(self, input, output, batch, **kw):
            seen.update(batch=batch, out_shape=tuple(output.shape), **kw)
            return output

        def __del__(self):
            pass

    r = _FakeReal(256)
    sig, win = torch.zeros(256 + 7 * 64 + 5), torch.ones(256)
    out = torch.empty((8, 128), dtype=torch.complex64)
    assert r.stft(sig, 64, win, out) is out
    assert seen["batch"] == 8 and seen["in_dist"] == 64 and seen["pre"] is win and seen["out_shape"] == (8, 128)
    with pytest.raises(ValueError):
        r.stft(torch.zeros(100), 64)          # shorter than one frame
    with pytest.raises(ValueError):
        r.stft(sig, 0)                        # hop must be positive
    with pytest.raises(ValueError):
        r.stft(sig.reshape(1, -1), 64)        # 1-D signals only
