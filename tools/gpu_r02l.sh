# round 2, GPU call l: RealFFT variants of the ticket-queue kernels (in-place rings, per-direction entries, wide tiles)
set -x
mkdir -p gpurun_out
timeout 900 python tools/flat_ab_real.py r02l 65536 131072 2>&1 | tee gpurun_out/flat_ab_real_r02l.txt
timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
SSFFT_FLAT_NAME=_w_ timeout 600 python -m pytest tests/test_gpu_flat.py -x -q 2>&1 | tail -5
SSFFT_FLAT_NAME=r1c3i timeout 600 python -m pytest tests/test_gpu_flat.py -x -q 2>&1 | tail -5
SSFFT_FLAT_NAME=r2c3i timeout 600 python -m pytest tests/test_gpu_flat.py -x -q 2>&1 | tail -5
