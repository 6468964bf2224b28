# round 2, GPU call ae: composite plan with the 16-byte radix pass -- parity, sweep, per-kernel times
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k composite 2>&1 | tail -4
timeout 600 python tools/sweep.py r02ae float32 1179648 2097152 4194304 16777216 1572864 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02ae_f32.txt
timeout 600 python tools/sweep.py r02ae float64 98304 147456 2097152 2>&1 | grep "^N=" | tee gpurun_out/sweep_r02ae_f64.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_composite_r02ae.csv python tools/prof_one.py c2c 2097152 > /dev/null 2>&1
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_composite_r02ae.csv')) if len(r)>5 and r[0].strip('"').isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    agg[r[4][:60]][0]+=1; agg[r[4][:60]][1]+=float(r[-1].replace(',',''))
for k,v in agg.items(): print(v[0], round(v[1]/v[0]/1e3,1), 'us per launch', k)
P
