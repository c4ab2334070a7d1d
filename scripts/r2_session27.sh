#!/bin/bash
# Round 2, GPU session 27 (1 GPU): ncu --set full of kernel F in the automatic schedule (tract x 24, 2M walks)
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --set full --import-source on -k regex:k_sgns_sent -s 1 -c 1 -f -o gpurun_out/r2s27_sgns_sent_tract24 python scripts/prof_path.py tract24 2000000 > gpurun_out/r2s27_ncu.log 2>&1; tail -2 gpurun_out/r2s27_ncu.log | cut -c1-300
