#!/bin/bash
# Round 2, GPU session 4 (1 GPU): how much concurrency does stage-2 parity at the full bench size tolerate?
mkdir -p gpurun_out
timeout 1200 python scripts/fullsize_staleness.py 0,384,256,128,64,32,16,8 0 2>&1 | tail -12
echo "== smem negative table variant (flag 1024), auto and c128"
timeout 600 python scripts/fullsize_staleness.py 0,128 1024 2>&1 | tail -3
