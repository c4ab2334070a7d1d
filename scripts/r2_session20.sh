#!/bin/bash
# Round 2, GPU session 20 (1 GPU): kernel F with write-through hot words, ONE block per SM of 12 / 13 / 14 warps (the same
# number of sentences on every SM), 128 / 256 / 512 write-through words.
mkdir -p gpurun_out
F=2048
for w in 12 13 14; do
  timeout 900 python scripts/fullsize_staleness.py $((148 * w)) $((F + (w << 12) + (8 << 20))),$((F + (w << 12) + (9 << 20))),$((F + (w << 12) + (10 << 20))) r2s20_w$w 2>&1 | grep -v Warning | tail -3
done
