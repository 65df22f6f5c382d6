#!/bin/bash
# round 2, session 2: cfg-3 all-Set pipeline capture, cfg-5 launch list
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:vlr_sets_lc_kernel -s 1 -c 1 \
    -o gpurun_out/ncu_r2s2_sets_lc -f python scripts/prof_wave.py 65536 2 3 > gpurun_out/r2s2_ncu_sets.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2s2_traffic_cfg5.csv python scripts/prof_wave.py 16384 2 5 > gpurun_out/r2s2_prof_cfg5.log 2>&1
tail -2 gpurun_out/r2s2_ncu_sets.log gpurun_out/r2s2_prof_cfg5.log
