# round 2, GPU call ai: 1536-point leg (9 * 2^17, 9 * 2^18, 3 * 2^20)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_round2.py -x -q -k "2048_point or composite" 2>&1 | tail -4
timeout 600 python tools/sweep.py r02ai float32 1179648 2359296 3145728 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02ai_f32.txt
