# round 2, GPU call s: 2^14 as a 128 x 128 ticket-queue four-step (256-byte runs both sides) against the single-pass kernel
set -x
mkdir -p gpurun_out
(timeout 300 python tools/sweep.py r02s_a float32 16384 32768 2>&1 | grep "^N=" | sed "s/^/single-pass default  /"
 SSFFT_FLAT_MIN_LOG2=14 timeout 300 python tools/sweep.py r02s_b float32 16384 32768 2>&1 | grep "^N=" | sed "s/^/flat 128x128          /") | tee gpurun_out/flat_ab_16384_r02s.txt
SSFFT_FLAT_MIN_LOG2=14 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "16384 or 32768" 2>&1 | tail -3
