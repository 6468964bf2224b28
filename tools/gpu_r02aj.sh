# round 2, GPU call aj: scratch slots of the longest four-step lengths (is the schedule's default too generous for L2?)
set -x
mkdir -p gpurun_out
for cfg in "default" "1 2" "1 3" "2 4"; do
  if [ "$cfg" = "default" ]; then env="";
  else set -- $cfg; env="SSFFT_FLAT_DELAY=$1 SSFFT_FLAT_SLOTS=$2"; fi
  env $env timeout 300 python tools/sweep.py r02aj float32 1048576 2097152 4194304 3145728 2>&1 | grep "^N=" | cut -c1-60 | sed "s/^/[$cfg] /" | tee -a gpurun_out/sweep_r02aj_slots.txt
done
