# round 2, GPU call i: C2R column tiles with 16-byte aligned boxes; the memory-system ceiling of a two-pass transform
set -x
mkdir -p gpurun_out
timeout 120 python tools/c2r_check.py 3 2>&1 | tail -3
timeout 120 ./tools/l2_ceiling.bin | tee gpurun_out/l2_ceiling_r02i.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_r02i.txt
timeout 900 python tools/sweep.py r02i float32 16384 32768 65536 131072 262144 524288 1048576 2097152 2>&1 | tee gpurun_out/sweep_r02i_f32.txt
timeout 600 python bench.py 2> gpurun_out/bench_r02i.err | tee gpurun_out/bench_r02i.json
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:fourstep_flat -s 2 -c 1 -o gpurun_out/prof_flat_c2r65536_r02i python tools/prof_one.py c2r 65536 > gpurun_out/prof_flat_i1.log 2>&1
timeout 300 $NCU -k regex:fourstep_flat -s 2 -c 1 -o gpurun_out/prof_flat_r2c65536_r02i python tools/prof_one.py r2c 65536 > gpurun_out/prof_flat_i2.log 2>&1
