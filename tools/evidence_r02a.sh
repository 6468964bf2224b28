# round 2, first GPU call: baseline sweep of the four-step sizes + ncu --set full of the kernels VERDICT r01 asks for
set -x
python tools/sweep.py r02base float32 16384 32768 65536 131072 262144 524288 1048576 2>&1 | tee gpurun_out/sweep_r02base_f32.txt
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:fourstep_cluster -s 1 -c 1 -o gpurun_out/prof_r2c65536_r02a python tools/prof_one.py r2c 65536 > gpurun_out/prof_r2c.log 2>&1
$NCU -k regex:fourstep_cluster -s 2 -c 1 -o gpurun_out/prof_c2r65536_r02a python tools/prof_one.py c2r 65536 > gpurun_out/prof_c2r.log 2>&1
$NCU -k regex:fourstep_cluster -s 1 -c 1 -o gpurun_out/prof_c2c1m_r02a python tools/prof_one.py c2c 1048576 > gpurun_out/prof_c2c1m.log 2>&1
$NCU -k regex:fused_fft -s 1 -c 1 -o gpurun_out/prof_c2c2187_r02a python tools/prof_one.py c2c 2187 > gpurun_out/prof_c2c2187.log 2>&1
$NCU -k regex:fused_fft -s 1 -c 1 -o gpurun_out/prof_c2c16384_r02a python tools/prof_one.py c2c 16384 > gpurun_out/prof_c2c16384.log 2>&1
tail -3 gpurun_out/prof_*.log
