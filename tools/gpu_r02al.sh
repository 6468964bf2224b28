# round 2, GPU call al: 2^18 as 1024 x 256 by default -- parity + sweep
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 600 python tools/sweep.py r02al float32 131072 262144 524288 2>&1 | grep "^N=" | cut -c1-150 | tee gpurun_out/sweep_r02al_f32.txt
