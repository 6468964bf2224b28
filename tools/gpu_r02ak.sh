# round 2, GPU call ak: unbalanced splits of 2^17 ... 2^20 (long column leg, 256- / 512-point row stage) against the balanced defaults
set -x
mkdir -p gpurun_out
SSFFT_FLAT_NAME=_u_ timeout 900 python -m pytest tests/test_gpu_flat.py -x -q -k "flat_vs_oracle or large_batch" 2>&1 | tail -4
(timeout 600 python tools/sweep.py r02ak_a float32 131072 262144 524288 1048576 2>&1 | grep "^N=" | cut -c1-130 | sed "s/^/balanced    /"
 SSFFT_FLAT_NAME=_u_ timeout 600 python tools/sweep.py r02ak_b float32 131072 262144 524288 1048576 2>&1 | grep "^N=" | cut -c1-130 | sed "s/^/unbalanced  /") | tee gpurun_out/sweep_r02ak_f32.txt
