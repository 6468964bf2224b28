# round 2, GPU call ah: 2048-point leg (2^21, 2^22, 3 * 2^19) against the composite plan
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_round2.py -x -q -k "2048_point or composite" 2>&1 | tail -4
(timeout 600 python tools/sweep.py r02ah_a float32 2097152 4194304 1572864 2>&1 | grep "^N=" | sed "s/^/ticket queue  /"
 SSFFT_DISABLE_FLAT=1 timeout 600 python tools/sweep.py r02ah_b float32 2097152 4194304 1572864 2>&1 | grep "^N=" | sed "s/^/composite     /") | tee gpurun_out/sweep_r02ah_f32.txt
