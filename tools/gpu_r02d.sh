# round 2, GPU call d: where does the time of the ticket-queue four-step go (statistics / no-compute builds)
# built here with:  make -C fft_b200/csrc OUT=../libssfft_stats.so BUILD=build_stats EXTRA=-DSSFFT_FLAT_STATS=1
#                   make -C fft_b200/csrc OUT=../libssfft_nocompute.so BUILD=build_nocompute EXTRA="-DSSFFT_FLAT_STATS=1 -DSSFFT_FLAT_NOCOMPUTE=1"
set -x
mkdir -p gpurun_out
timeout 1200 python tools/flat_stats.py 65536 262144 1048576 2>&1 | tee gpurun_out/flat_stats_r02d.txt
