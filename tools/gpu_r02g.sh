# round 2, GPU call g: RealFFT on the ticket-queue kernels (C3), full GPU suite, sweep of the four-step sizes, bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 900 python tools/sweep.py r02g float32 16384 32768 65536 131072 262144 524288 1048576 2097152 2>&1 | tee gpurun_out/sweep_r02g_f32.txt
SSFFT_DISABLE_FLAT_REAL=1 timeout 600 python tools/sweep.py r02g_noflatreal float32 65536 131072 262144 1048576 2>&1 | tee gpurun_out/sweep_r02g_noflatreal_f32.txt
timeout 600 python bench.py --no-e2e --no-cpu 2> gpurun_out/bench_r02g.err | tee gpurun_out/bench_r02g.json
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:fourstep_flat -s 1 -c 1 -o gpurun_out/prof_flat_r2c65536_r02g python tools/prof_one.py r2c 65536 > gpurun_out/prof_flat_g1.log 2>&1
timeout 300 $NCU -k regex:fourstep_flat -s 2 -c 1 -o gpurun_out/prof_flat_c2r65536_r02g python tools/prof_one.py c2r 65536 > gpurun_out/prof_flat_g2.log 2>&1
