# round 2, GPU call y (2 GPUs): the TMA-driven exchange kernel -- parity (logical ranks + real devices), then C5 per chunk count, against the old kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_multi.py -x -q -k "distributed or exchange or logical" 2>&1 | tail -6 | tee gpurun_out/pytest_exchange_tma_r02y.txt
fmt='
import sys, json
tag = sys.argv[1]
for l in sys.stdin:
    d = json.loads(l); print(tag, d["gpus"], "GPUs chunks", d["chunks"], "transposed" if d["transposed_output"] else "natural   ", round(d["ms"], 2), "ms", (d.get("checks") or {}).get("ok"), d.get("error", ""))
'
SSFFT_BENCH_DIST_CHUNKS="1,2,4,8" timeout 600 python tools/bench_dist_local.py 30 2>&1 | grep "^{" | python -c "$fmt" "tma exchange, 2 CTAs/SM " | tee gpurun_out/bench_dist_tma_r02y.txt
SSFFT_EXCHANGE_CTAS_PER_SM=1 SSFFT_BENCH_DIST_CHUNKS="1,4" timeout 600 python tools/bench_dist_local.py 30 2>&1 | grep "^{" | python -c "$fmt" "tma exchange, 1 CTA/SM  " | tee -a gpurun_out/bench_dist_tma_r02y.txt
SSFFT_EXCHANGE_CTAS_PER_SM=4 SSFFT_BENCH_DIST_CHUNKS="1,4" timeout 600 python tools/bench_dist_local.py 30 2>&1 | grep "^{" | python -c "$fmt" "tma exchange, 4 CTAs/SM " | tee -a gpurun_out/bench_dist_tma_r02y.txt
SSFFT_EXCHANGE_TMA=0 SSFFT_BENCH_DIST_CHUNKS="1,4" timeout 600 python tools/bench_dist_local.py 30 2>&1 | grep "^{" | python -c "$fmt" "old exchange kernel     " | tee -a gpurun_out/bench_dist_tma_r02y.txt
