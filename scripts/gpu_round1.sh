#!/bin/bash
# one GPU session: parity tests, A/B bench (generic vs wavefront), launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --loci 200000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wave_200k.json 2> gpurun_out/bench_wave_200k.err; tail -c 1500 gpurun_out/bench_wave_200k.json; tail -3 gpurun_out/bench_wave_200k.err
VLR_WAVE=0 python bench.py --loci 200000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_generic_200k.json 2> gpurun_out/bench_generic_200k.err; tail -c 600 gpurun_out/bench_generic_200k.json
