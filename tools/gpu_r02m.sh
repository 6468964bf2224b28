# round 2, GPU call m (2 GPUs): distributed transform with real ranks and through the C ABI; the bench line at N = 2
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_multi_r02m.txt
timeout 600 python tools/bench_dist_local.py 30 2>&1 | tee gpurun_out/bench_dist_local_r02m.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/bench_2gpu_r02m.err | tee gpurun_out/bench_2gpu_r02m.json
tail -5 gpurun_out/bench_2gpu_r02m.err
