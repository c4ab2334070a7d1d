#!/bin/bash
# Round 2, GPU session 25 (1 GPU): kernel F, dynamic hand-out + write-through words: one 640-thread block per SM (tract x 24),
# and the same schedule on the CA workload (V = 1 848, D = 8) against its full-size oracle fixture.
mkdir -p gpurun_out
D=$((2048 + (1 << 25)))
echo "== tract x 24: 640-thread blocks, 2960 in flight, 256 / 512 / 1024 words; 15 warps per SM, 512 words"
timeout 900 python scripts/fullsize_staleness.py 2960 $((D + (9 << 20))),$((D + (10 << 20))),$((D + (11 << 20))) r2s25_big 2>&1 | grep -v Warning | tail -3
timeout 900 python scripts/fullsize_staleness.py 2220 $((D + (15 << 12) + (10 << 20))) r2s25_w15 2>&1 | grep -v Warning | tail -1
echo "== CA: automatic (kernel G), then kernel F dynamic with every word written through at 2960 / 1480 / 592, with 256 words at 592"
timeout 900 python scripts/fullsize_ca.py 0 0 r2s25_auto 2>&1 | grep -v Warning | tail -2
timeout 900 python scripts/fullsize_ca.py 2960,1480,592 $((D + (12 << 20))) r2s25_all 2>&1 | grep -v Warning | tail -3
timeout 900 python scripts/fullsize_ca.py 592 $((D + (9 << 20))) r2s25_256 2>&1 | grep -v Warning | tail -1
