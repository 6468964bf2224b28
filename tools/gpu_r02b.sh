# round 2, GPU call b: first run of the ticket-queue four-step (flat.cuh)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flat.py -x -q 2>&1 | tail -15
timeout 900 python tools/flat_ab.py r02b 32768 65536 2>&1 | tee gpurun_out/flat_ab_r02b.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:fourstep_flat -s 1 -c 1 -o gpurun_out/prof_flat65536_r1c3_r02b python tools/prof_one.py c2c 65536 > gpurun_out/prof_flat_a.log 2>&1
SSFFT_FLAT_VARIANT=2,2 timeout 300 $NCU -k regex:fourstep_flat -s 1 -c 1 -o gpurun_out/prof_flat65536_r2c2_r02b python tools/prof_one.py c2c 65536 > gpurun_out/prof_flat_b.log 2>&1
tail -3 gpurun_out/prof_flat_*.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
