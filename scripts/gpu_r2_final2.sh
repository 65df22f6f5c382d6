#!/bin/bash
# round 2, closing session (one GPU) after the last kernel change: full -m gpu suite, smoke, default bench line, and the
# capture of the all-Set lc kernel with its MAPQ table. Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2h_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
ncu --set full --clock-control none --import-source on -k regex:vlr_sets_lc_kernel -s 1 -c 1 \
    -o gpurun_out/ncu_r2h_sets_lc -f python scripts/prof_wave.py 65536 2 3 > gpurun_out/r2h_ncu_sets.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2h_traffic_cfg3.csv python scripts/prof_wave.py 65536 2 3 > gpurun_out/r2h_prof_cfg3.log 2>&1
python scripts/sum_launches.py gpurun_out/r2h_traffic_cfg3.csv
cut -c1-300 gpurun_out/r2h_bench.json
