#!/bin/bash
# Round 2, GPU session 19 (1 GPU): kernel F with write-through hot words, 128-thread blocks (spread over all SMs):
# throughput and parity against sentences in flight and the number of write-through words.
mkdir -p gpurun_out
F=$((2048 + 16))
timeout 1500 python scripts/fullsize_staleness.py 1184,1480,1776,2072,2368 $((F + (7 << 20))),$((F + (9 << 20))) r2s19_small 2>&1 | grep -v Warning | tail -11
timeout 900 python scripts/fullsize_staleness.py 1480,2072 $((F + (5 << 20))),$((F + (11 << 20))) r2s19_small_b 2>&1 | grep -v Warning | tail -4
