# round 2, GPU call o: tensor-core probe with several groups per CTA (+ ncu of both passes), composite parity, full GPU suite, sweep, bench
set -x
mkdir -p gpurun_out
timeout 300 ./tools/tc_probe.bin 2000 2>&1 | tee gpurun_out/tc_probe_r02o.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:tc_pass_kernel -s 11 -c 1 -o gpurun_out/prof_tc_pass_r02o ./tools/tc_probe.bin 500 > gpurun_out/prof_tc_o1.log 2>&1
timeout 300 $NCU -k regex:fp32_pass_kernel -s 21 -c 1 -o gpurun_out/prof_fp32_pass_r02o ./tools/tc_probe.bin 500 > gpurun_out/prof_tc_o2.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02o.txt
timeout 900 python tools/sweep.py r02o float32 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02o_f32.txt
timeout 900 python tools/sweep.py r02o float64 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02o_f64.txt
timeout 600 python bench.py 2> gpurun_out/bench_r02o.err | tee gpurun_out/bench_r02o.json
