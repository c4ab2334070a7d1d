#!/bin/bash
# Round 2, GPU session 24 (1 GPU): kernel F, sentences handed out from a counter, 128-thread blocks: how far can the
# sentences in flight go with 256 / 512 write-through words?
mkdir -p gpurun_out
F=$((2048 + 16 + (1 << 25)))
timeout 1500 python scripts/fullsize_staleness.py 2220,2368,2664,2960 $((F + (9 << 20))),$((F + (10 << 20))) r2s24 2>&1 | grep -v Warning | tail -8
