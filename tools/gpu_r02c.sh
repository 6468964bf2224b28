# round 2, GPU call c: ticket-queue four-step with in-place ring slots, 2^15 .. 2^20
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flat.py -x -q 2>&1 | tail -15
timeout 1500 python tools/flat_ab.py r02c 32768 65536 131072 262144 524288 1048576 2>&1 | tee gpurun_out/flat_ab_r02c.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:fourstep_flat -s 1 -c 1 -o gpurun_out/prof_flat65536_r02c python tools/prof_one.py c2c 65536 > gpurun_out/prof_flat_c1.log 2>&1
timeout 300 $NCU -k regex:fourstep_flat -s 1 -c 1 -o gpurun_out/prof_flat1m_r02c python tools/prof_one.py c2c 1048576 > gpurun_out/prof_flat_c2.log 2>&1
tail -n 3 gpurun_out/prof_flat_c1.log gpurun_out/prof_flat_c2.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
