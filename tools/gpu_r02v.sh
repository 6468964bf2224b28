# round 2, GPU call v: end-of-round validation -- full GPU suite, sweeps, the reference's benchmark set in its own format (fp64 + fp32 twin), bench line
set -x
mkdir -p gpurun_out gpurun_out/results
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02v.txt
timeout 900 python tools/sweep.py r02v float32 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02v_f32.txt
timeout 900 python tools/sweep.py r02v float64 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02v_f64.txt
timeout 900 python tools/reference_benchmark.py gpurun_out/results 24 float64 2>&1 | tail -75 > gpurun_out/refbench_f64_r02v.txt
timeout 900 python tools/reference_benchmark.py gpurun_out/results 24 float32 2>&1 | tail -75 > gpurun_out/refbench_f32_r02v.txt
timeout 600 python bench.py 2> gpurun_out/bench_r02v.err | tee gpurun_out/bench_r02v.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_r02v.err | tee gpurun_out/bench_reference_r02v.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" 2>&1 | tail -2
