# round 2, GPU call af: fp64 2^19 / 2^20 on the ticket-queue kernels (1024-point leg, 16 points per thread)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flat.py -x -q -k "double_precision" 2>&1 | tail -4
(SSFFT_DISABLE_FLAT=1 timeout 600 python tools/sweep.py r02af_a float64 524288 1048576 2>&1 | grep "^N=" | sed "s/^/round-1 kernels  /"
 timeout 600 python tools/sweep.py r02af_b float64 524288 1048576 2>&1 | grep "^N=" | sed "s/^/ticket queue     /") | tee gpurun_out/sweep_r02af_f64.txt
