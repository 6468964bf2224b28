# end-of-round evidence pass (run on the GPU box through gpurun): tests, bench lines, sweeps, launch lists, profiles
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_default_r01.json 2> gpurun_out/bench_default_r01.err; tail -1 gpurun_out/bench_default_r01.json | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_r01.json 2>/dev/null
for w in c3 c4-1000 c4-2187 c4-3125 c4-6000 c4-1000-f64 c4-2187-f64 c4-3125-f64 c4-6000-f64 c1; do python bench.py --workload $w --no-e2e --no-cpu --steps 10 2>/dev/null | tail -1 > gpurun_out/bench_${w}_r01.json; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${w}_r01.json')); print('$w', d['ms_per_step'], d['value'], d['roofline']['frac'])"; done
timeout 500 python tools/sweep.py r01final3 float32 2>&1 | tee gpurun_out/sweep_final3_f32.txt | tail -20
timeout 500 python tools/sweep.py r01final3 float64 2>&1 | tee gpurun_out/sweep_final3_f64.txt | tail -20
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c2_r01.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_fft -s 3 -c 1 -f -o gpurun_out/prof_c2_tma_r01 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
