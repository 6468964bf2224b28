# round 2, GPU call h: which C2R column-tile variant works (all boxes inside the tensor vs the one-column overhang + 2-column box)
set -x
mkdir -p gpurun_out
timeout 120 python tools/c2r_check.py 3 2>&1 | tail -4
SSFFT_LIB=$PWD/fft_b200/libssfft_c2r_oob.so timeout 120 python tools/c2r_check.py 3 2>&1 | tail -4
SSFFT_LIB=$PWD/fft_b200/libssfft_c2r_oob.so timeout 300 compute-sanitizer --tool memcheck python tools/c2r_check.py 1 2>&1 | grep -v "^$" | head -40 | tee gpurun_out/c2r_oob_memcheck.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 900 python tools/sweep.py r02h float32 32768 65536 131072 262144 524288 1048576 2097152 2>&1 | tee gpurun_out/sweep_r02h_f32.txt
timeout 600 python bench.py --no-e2e --no-cpu 2> gpurun_out/bench_r02h.err | tee gpurun_out/bench_r02h.json
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:fourstep_flat -s 2 -c 1 -o gpurun_out/prof_flat_c2r65536_r02h python tools/prof_one.py c2r 65536 > gpurun_out/prof_flat_h2.log 2>&1
