# round 2, GPU call e: ticket-queue four-step with a separate signaller warp; statistics / no-compute / no-dependency builds
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flat.py -x -q 2>&1 | tail -5
timeout 1500 python tools/flat_ab.py r02e 32768 65536 262144 1048576 2>&1 | tee gpurun_out/flat_ab_r02e.txt
timeout 1200 python tools/flat_stats.py 65536 262144 1048576 2>&1 | tee gpurun_out/flat_stats_r02e.txt
