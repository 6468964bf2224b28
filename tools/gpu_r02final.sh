# round 2, final validation: full GPU suite, sweeps, the reference's benchmark set in its own format (fp64 + fp32 twin), bench line + reference arm, smoke
set -x
mkdir -p gpurun_out gpurun_out/results
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r02final.txt
timeout 900 python tools/sweep.py r02final float32 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02final_f32.txt
timeout 900 python tools/sweep.py r02final float64 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02final_f64.txt
timeout 900 python tools/reference_benchmark.py gpurun_out/results 24 float64 > gpurun_out/refbench_f64_r02final.txt 2>&1
timeout 900 python tools/reference_benchmark.py gpurun_out/results 24 float32 > gpurun_out/refbench_f32_r02final.txt 2>&1
timeout 600 python bench.py 2> gpurun_out/bench_r02final.err | tee gpurun_out/bench_r02final.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_r02final.err | tee gpurun_out/bench_reference_r02final.json | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" 2>&1 | tail -2
