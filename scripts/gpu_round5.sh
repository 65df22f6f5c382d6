#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_prep_kernel -s 1 -c 1 -o gpurun_out/ncu_wave_prep_v5 -f python scripts/prof_wave.py 65536 2 > gpurun_out/ncu_prep.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/traffic_wave.csv python scripts/prof_wave.py 65536 2 > gpurun_out/prof_wave.log 2>&1
tail -3 gpurun_out/traffic_wave.csv
