# round 2, GPU call p2: row stages of length 3 * 2^j (parity + sweep) with the ticket-queue path actually first
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_flat.py -x -q -k "three_times" 2>&1 | tail -8 | tee gpurun_out/pytest_flat3_r02p.txt
timeout 900 python tools/sweep.py r02p float32 12288 24576 49152 98304 196608 393216 786432 1572864 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02p_f32.txt
