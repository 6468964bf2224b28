# round 2, GPU call q (8 GPUs): C5 through the C ABI at 2 / 4 / 8 GPUs, the distributed tests with real devices, the bench line at N = 8
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_multi_r02q.txt
SSFFT_BENCH_DIST_CHUNKS="1,4" timeout 600 python tools/bench_dist_local.py 30 2>&1 | tee gpurun_out/bench_dist_local_r02q.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/bench_8gpu_r02q.err | grep "^{" | tee gpurun_out/bench_8gpu_r02q.json
tail -3 gpurun_out/bench_8gpu_r02q.err
