#!/bin/bash
for sub in 8192 16384 32768 65536 131072 262144; do echo "sub $sub"; VLR_WAVE_SUB=$sub python scripts/prof_wave.py 262144 3 2>&1 | tail -1; done
