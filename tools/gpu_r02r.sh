# round 2, GPU call r (2 GPUs): exchange of chunk c beside the transform of chunk c + 1 (2 of 3 CTAs per SM for the transforms, stream priorities)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4
SSFFT_BENCH_DIST_CHUNKS="1,2,4,8" timeout 600 python tools/bench_dist_local.py 30 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('share   ', d['gpus'], 'GPUs chunks', d['chunks'], 'transposed' if d['transposed_output'] else 'natural   ', round(d['ms'], 2), 'ms', (d.get('checks') or {}).get('ok'))
" | tee gpurun_out/bench_dist_share_r02r.txt
SSFFT_DIST_NO_SHARE=1 SSFFT_BENCH_DIST_CHUNKS="4" timeout 600 python tools/bench_dist_local.py 30 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('no share', d['gpus'], 'GPUs chunks', d['chunks'], 'transposed' if d['transposed_output'] else 'natural   ', round(d['ms'], 2), 'ms', (d.get('checks') or {}).get('ok'))
" | tee -a gpurun_out/bench_dist_share_r02r.txt
