# round 2, GPU call x (2 GPUs): ncu of the fused transpose + peer-store exchange kernel with the NVLink byte counters
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
SSFFT_BENCH_DIST_CHUNKS=1 timeout 600 $NCU --section Nvlink --section Nvlink_Tables --metrics nvltx__bytes.sum,nvlrx__bytes.sum \
  -k regex:exchange_transpose -s 8 -c 2 -o gpurun_out/prof_exchange_r02x python tools/bench_dist_local.py 28 > gpurun_out/prof_exchange_x.log 2>&1
tail -5 gpurun_out/prof_exchange_x.log
ncu -i gpurun_out/prof_exchange_r02x.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys, csv
rows = list(csv.reader(sys.stdin))
if len(rows) > 2:
    h = rows[0]
    want = [i for i, n in enumerate(h) if any(k in n for k in ('Kernel Name', 'gpu__time_duration.sum', 'nvltx__bytes', 'nvlrx__bytes', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct', 'launch__grid_size', 'launch__registers'))]
    for r in rows[1:]:
        print({h[i]: r[i] for i in want})
" | tee gpurun_out/exchange_nvlink_r02x.txt
