# round 2, GPU call j: strided variants of the memory-system ceiling; flat plans without a round-1 fallback; bench line
set -x
mkdir -p gpurun_out
timeout 300 ./tools/l2_ceiling.bin | tee gpurun_out/l2_ceiling_r02j.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_r02j.txt
timeout 600 python bench.py 2> gpurun_out/bench_r02j.err | tee gpurun_out/bench_r02j.json
