#!/bin/bash
# Round 2, GPU session 16 (1 GPU): ncu --set full of kernel I (a warp per pair) on tract x 24, 2M walks
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --set full --import-source on -k regex:k_sgns_wave -s 1 -c 1 -f -o gpurun_out/r2s16_sgns_wave_tract24 python scripts/prof_path.py tract24 2000000 flags=262144 > gpurun_out/r2s16_ncu.log 2>&1; tail -2 gpurun_out/r2s16_ncu.log | cut -c1-300
