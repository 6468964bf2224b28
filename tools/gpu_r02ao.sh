# round 2, GPU call ao: 2^19 as 1024 x 512 against 512 x 1024
set -x
mkdir -p gpurun_out
(timeout 300 python tools/sweep.py r02ao_a float32 524288 2>&1 | grep "^N=" | cut -c1-130 | sed "s/^/512 x 1024   /"
 SSFFT_FLAT_NAME=_u_ timeout 300 python tools/sweep.py r02ao_b float32 524288 2>&1 | grep "^N=" | cut -c1-130 | sed "s/^/1024 x 512   /") | tee gpurun_out/sweep_r02ao_f32.txt
SSFFT_FLAT_NAME=_u_ timeout 600 python -m pytest tests/test_gpu_flat.py -x -q -k "flat_vs_oracle and 524288" 2>&1 | tail -3
