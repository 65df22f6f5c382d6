#!/bin/bash
# round 2, session 2: state of HEAD on the B200 — full -m gpu suite, bench line, launch list + traffic, full capture of the resident kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2s2_gputest.log 2>&1
tail -3 gpurun_out/r2s2_gputest.log
python bench.py > gpurun_out/r2s2_bench.json 2> gpurun_out/r2s2_bench.err
cut -c1-300 gpurun_out/r2s2_bench.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/r2s2_traffic.csv python scripts/prof_wave.py 65536 2 > gpurun_out/r2s2_prof_wave.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_resident_kernel -s 1 -c 1 \
    -o gpurun_out/ncu_r2s2_resident -f python scripts/prof_wave.py 65536 2 > gpurun_out/r2s2_ncu.log 2>&1
tail -2 gpurun_out/r2s2_prof_wave.log
