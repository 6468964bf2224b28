# round 2, GPU call k: wide tiles (512 consumer threads, runs twice as long) against the registered defaults
set -x
mkdir -p gpurun_out
timeout 1500 python tools/flat_ab.py r02k 32768 65536 131072 262144 524288 1048576 2>&1 | tee gpurun_out/flat_ab_r02k.txt
SSFFT_FLAT_NAME=_w_ timeout 600 python -m pytest tests/test_gpu_flat.py -x -q 2>&1 | tail -5
