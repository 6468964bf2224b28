# round 2, GPU call ad: launch lists of the bench command (C2, C3) and one --set full capture of the headline kernel on this round's build
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_r02.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/launches_c2_r02.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3_r02.csv python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/launches_c3_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:fused_fft -s 3 -c 1 -o gpurun_out/prof_c2_fused4096_r02 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/prof_c2_r02.log 2>&1
tail -3 gpurun_out/launches_c2_r02.csv | cut -c1-200
tail -3 gpurun_out/launches_c3_r02.csv | cut -c1-200
