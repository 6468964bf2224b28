# round 2, GPU call f: full GPU suite (Bluestein, host path, single-process distributed plan on logical ranks), the two
# sizes call e skipped, ncu of the 65536 kernel with the signaller warp, the default bench line with its configs object
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 900 python tools/flat_ab.py r02f 131072 524288 2>&1 | tee gpurun_out/flat_ab_r02f.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:fourstep_flat -s 1 -c 1 -o gpurun_out/prof_flat65536_r02f python tools/prof_one.py c2c 65536 > gpurun_out/prof_flat_f1.log 2>&1
timeout 600 python bench.py 2> gpurun_out/bench_r02f.err | tee gpurun_out/bench_r02f.json
tail -n 5 gpurun_out/bench_r02f.err
