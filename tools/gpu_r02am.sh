# round 2, GPU call am: unbalanced splits of 3 * 2^k / 9 * 2^k
set -x
mkdir -p gpurun_out
(timeout 600 python tools/sweep.py r02am_a float32 98304 196608 393216 294912 589824 2>&1 | grep "^N=" | cut -c1-130 | sed "s/^/balanced    /"
 SSFFT_FLAT_NAME=_u_ timeout 600 python tools/sweep.py r02am_b float32 98304 196608 393216 294912 589824 2>&1 | grep "^N=" | cut -c1-130 | sed "s/^/unbalanced  /") | tee gpurun_out/sweep_r02am_f32.txt
SSFFT_FLAT_NAME=_u_ timeout 600 python -m pytest tests/test_gpu_flat.py -x -q -k "three_times" 2>&1 | tail -3
