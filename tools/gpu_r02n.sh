# round 2, GPU call n: composite plans (parity + sweep, work-buffer size), radix-4-first 1024-point tiles, defaults check
set -x
mkdir -p gpurun_out
timeout 120 ./tools/tc_probe.bin 2000 2>&1 | tee gpurun_out/tc_probe_r02n.txt
timeout 1200 python -m pytest tests/test_gpu_round2.py -x -q -k composite 2>&1 | tail -8 | tee gpurun_out/pytest_composite_r02n.txt
SIZES="12288 24576 49152 98304 196608 393216 786432 18432 36864 73728 147456 294912 589824 2097152 4194304 8388608 16777216"
timeout 900 python tools/sweep.py r02n float32 $SIZES 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02n_f32.txt
for MB in 16 96 4096; do
  SSFFT_COMPOSITE_MB=$MB timeout 600 python tools/sweep.py r02n_mb$MB float32 12288 49152 196608 786432 147456 2097152 2>&1 | grep "^N=" | sed "s/^/MB=$MB /" | tee -a gpurun_out/sweep_r02n_mb_f32.txt
done
SSFFT_DISABLE_COMPOSITE=1 timeout 600 python tools/sweep.py r02n_nocomp float32 12288 196608 147456 2097152 16777216 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02n_nocomp_f32.txt
timeout 600 python tools/sweep.py r02n_p4a float32 524288 1048576 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02n_p4_f32.txt
SSFFT_FLAT_NAME=_p4_ timeout 600 python tools/sweep.py r02n_p4b float32 524288 1048576 2>&1 | grep "^N=" | sed "s/^/p4 /" | tee -a gpurun_out/sweep_r02n_p4_f32.txt
timeout 600 python tools/sweep.py r02n_real float32 65536 131072 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02n_real_f32.txt
