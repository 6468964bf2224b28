# round 2, GPU call ap (4 GPUs): the bench line at N = 4 (the driver's scaling run passes through it)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 2> gpurun_out/bench_4gpu_r02ap.err | grep "^{" | tee gpurun_out/bench_4gpu_r02ap.json | cut -c1-300
tail -3 gpurun_out/bench_4gpu_r02ap.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2>/dev/null | grep "^{" | cut -c1-200
