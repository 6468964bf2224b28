# round 2, GPU call ag: fp64 98304 / 196608 on the ticket-queue kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flat.py -x -q -k "double_precision" 2>&1 | tail -4
timeout 600 python tools/sweep.py r02ag float64 98304 196608 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02ag_f64.txt
