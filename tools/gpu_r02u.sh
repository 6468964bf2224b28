# round 2, GPU call u: fp64 on the ticket-queue kernels (parity, sweep against the round-1 kernels)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_flat.py -x -q -k "double_precision" 2>&1 | tail -8 | tee gpurun_out/pytest_flat64_r02u.txt
(SSFFT_DISABLE_FLAT=1 timeout 600 python tools/sweep.py r02u_a float64 16384 32768 65536 131072 262144 2>&1 | grep "^N=" | sed "s/^/round-1 kernels  /"
 timeout 600 python tools/sweep.py r02u_b float64 16384 32768 65536 131072 262144 2>&1 | grep "^N=" | sed "s/^/ticket queue     /") | tee gpurun_out/sweep_r02u_f64.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
