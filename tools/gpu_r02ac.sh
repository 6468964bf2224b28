# round 2, GPU call ac: RealFFT 65536 with a ring of three slots at 2 CTAs/SM
set -x
mkdir -p gpurun_out
timeout 600 python tools/flat_ab_real.py r02ac 65536 2>&1 | tee gpurun_out/flat_ab_real_r02ac.txt
SSFFT_FLAT_NAME=r3c2i timeout 300 python -m pytest tests/test_gpu_flat.py -x -q -k "real_vs_oracle and 65536" 2>&1 | tail -3
