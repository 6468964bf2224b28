# round 2, GPU call an: ncu --set full of the kernels added at the end of the round
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:fourstep_flat -s 2 -c 1 -o gpurun_out/prof_flat49152_r02an python tools/prof_one.py c2c 49152 > gpurun_out/prof_an1.log 2>&1
timeout 300 $NCU -k regex:fourstep_flat -s 2 -c 1 -o gpurun_out/prof_flat147456_r02an python tools/prof_one.py c2c 147456 > gpurun_out/prof_an2.log 2>&1
SSFFT_PROF_PREC=float64 timeout 300 $NCU -k regex:fourstep_flat -s 2 -c 1 -o gpurun_out/prof_flat65536_f64_r02an python tools/prof_one.py c2c 65536 > gpurun_out/prof_an3.log 2>&1
timeout 300 $NCU -k regex:tiny_fft -s 2 -c 1 -o gpurun_out/prof_tiny16_r02an python tools/prof_one.py c2c 16 > gpurun_out/prof_an4.log 2>&1
timeout 300 $NCU -k regex:"radix_pass|interleave" -s 2 -c 2 -o gpurun_out/prof_composite_passes_r02an python tools/prof_one.py c2c 8388608 > gpurun_out/prof_an5.log 2>&1
timeout 300 $NCU -k regex:fourstep_flat -s 2 -c 1 -o gpurun_out/prof_flat4m_r02an python tools/prof_one.py c2c 4194304 > gpurun_out/prof_an6.log 2>&1
ls -la gpurun_out/*_r02an.ncu-rep
