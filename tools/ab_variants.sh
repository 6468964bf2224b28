#!/bin/bash
# tools/ab_variants.sh -- A/B measurement of compile-time experiment switches of the four-step kernels.
#   1. here (CPU box):   tools/ab_variants.sh build          # extra builds next to fft_b200/libssfft.so (they travel with gpurun)
#   2. on the GPU box:   gpurun --timeout 600 -- tools/ab_variants.sh run
# Every variant is the SAME library built with one more -D switch (fft_b200/csrc/Makefile EXTRA=...), loaded through
# SSFFT_LIB; results land in gpurun_out/sweep_<variant>_float32.json (tools/sweep.py: % of the HBM roofline per size).
set -e
cd "$(dirname "$0")/.."
VARIANTS="l2pf:-DSSFFT_FOURSTEP_L2PF=1 fusedl2pf:-DSSFFT_FUSED_L2PF=1"
SIZES="2048 4096 8192 16384 6144 9216 32768 65536 131072 262144 524288 1048576"
case "$1" in
build)
    for v in $VARIANTS; do
        name=${v%%:*}; flags=${v#*:}
        make -C fft_b200/csrc -j8 OUT=../libssfft_$name.so BUILD=build_$name EXTRA="$flags" > /dev/null
        echo "built fft_b200/libssfft_$name.so ($flags)"
    done ;;
run)
    mkdir -p gpurun_out
    python tools/sweep.py ab_default float32 $SIZES
    for v in $VARIANTS; do
        name=${v%%:*}
        SSFFT_LIB=$PWD/fft_b200/libssfft_$name.so python tools/sweep.py ab_$name float32 $SIZES
        # run-time knobs on top of the variant: CTAs per cluster of the persistent four-step launch
        for cs in 2 8; do
            SSFFT_CLUSTER=$cs SSFFT_LIB=$PWD/fft_b200/libssfft_$name.so python tools/sweep.py ab_${name}_cluster$cs float32 $SIZES
        done
    done
    for cs in 2 8; do SSFFT_CLUSTER=$cs python tools/sweep.py ab_default_cluster$cs float32 $SIZES; done ;;
*) echo "usage: $0 build|run"; exit 2 ;;
esac
