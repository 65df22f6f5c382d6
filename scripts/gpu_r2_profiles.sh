#!/bin/bash
# round 2: the captures behind profiles/*_r2b_* (run under gpurun, one GPU). Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
# launch lists + DRAM bytes of one device-entry call (second call of the script): config 2 (65536 loci), 5 (16384), 3 (65536)
for c in 2 5 3; do
  n=65536; [ $c = 5 ] && n=16384
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/r2b_traffic_cfg$c.csv python scripts/prof_wave.py $n 2 $c > gpurun_out/r2b_prof_cfg$c.log 2>&1
  python scripts/sum_launches.py gpurun_out/r2b_traffic_cfg$c.csv > gpurun_out/r2b_launches_cfg$c.txt
done
# full captures: the octet resident kernel on config 2, the deep resident kernel (class 3) on config 5
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_resident_kernel -s 1 -c 1 \
    -o gpurun_out/ncu_r2b_resident -f python scripts/prof_wave.py 65536 2 2 > gpurun_out/r2b_ncu_res.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_resident_deep_kernel -s 3 -c 1 \
    -o gpurun_out/ncu_r2b_resident_deep -f python scripts/prof_wave.py 16384 2 5 > gpurun_out/r2b_ncu_deep.log 2>&1
cat gpurun_out/r2b_launches_cfg2.txt gpurun_out/r2b_launches_cfg5.txt
