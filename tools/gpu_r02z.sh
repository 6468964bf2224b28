# round 2, GPU call z2: unaligned fallbacks + alignment contract, then the full suite and the bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_flat.py -x -q -k "tiny or unaligned" 2>&1 | tail -6 | tee gpurun_out/pytest_tiny_r02z.txt
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r02z.txt
timeout 600 python bench.py 2> gpurun_out/bench_r02z.err | tee gpurun_out/bench_r02z.json | cut -c1-300
