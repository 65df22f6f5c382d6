#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_cfg2_1m.json 2> gpurun_out/bench_cfg2_1m.err; tail -c 2500 gpurun_out/bench_cfg2_1m.json; tail -3 gpurun_out/bench_cfg2_1m.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 700 gpurun_out/bench_reference.json
python bench.py --config 5 --loci 100000 --no-cpu-baseline > gpurun_out/bench_cfg5_100k.json 2> gpurun_out/bench_cfg5.err; tail -c 900 gpurun_out/bench_cfg5_100k.json; tail -2 gpurun_out/bench_cfg5.err
