#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_cfg2_1m.json 2> gpurun_out/bench_cfg2_1m.err; tail -c 300 gpurun_out/bench_cfg2_1m.json; tail -2 gpurun_out/bench_cfg2_1m.err
python bench.py --config 5 --loci 100000 --no-cpu-baseline > gpurun_out/bench_cfg5_100k.json 2> gpurun_out/bench_cfg5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg5_100k.json')); print('cfg5', d['value'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/traffic_wave.csv python scripts/prof_wave.py 65536 2 > gpurun_out/prof_wave.log 2>&1
