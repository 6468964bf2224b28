# round 2, GPU call w: compute-sanitizer memcheck / synccheck over every kernel family incl. the round-2 ones
set -x
mkdir -p gpurun_out
timeout 120 python tools/sanitize_smoke.py 2>&1 | tail -5
timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | grep -v "^$" | tail -45 | tee gpurun_out/sanitize_memcheck_r02w.txt
timeout 1200 compute-sanitizer --tool synccheck python tools/sanitize_smoke.py 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/sanitize_synccheck_r02w.txt
