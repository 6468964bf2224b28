# round 2, GPU call t: 9 * 2^k on the ticket-queue kernels (parity + sweep), composite parity for what is left
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_flat.py -x -q -k "three_times" 2>&1 | tail -6 | tee gpurun_out/pytest_flat39_r02t.txt
timeout 1200 python -m pytest tests/test_gpu_round2.py -x -q -k composite 2>&1 | tail -6 | tee gpurun_out/pytest_composite_r02t.txt
timeout 900 python tools/sweep.py r02t float32 18432 36864 73728 147456 294912 589824 1179648 2097152 4194304 16777216 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02t_f32.txt
