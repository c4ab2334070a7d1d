#!/bin/bash
# Round 2, GPU session 18 (1 GPU): kernel F (a warp per sentence) with WRITE-THROUGH hot words -- does parity hold with a
# full GPU of sentences in flight once the most frequent words' rows are no longer held pending?
mkdir -p gpurun_out
F=2048
echo "== arithmetic test (kernel F unchanged at hot = 0)"; timeout 600 python -m pytest tests/test_sgns_gpu.py -m gpu -q -x --tb=short -k "arithmetic" 2>&1 | tail -3
echo "== full size: kernel F, 2960 sentences in flight, hot = 0 / 32 / 64 / 128 / 256 / 1024"
timeout 900 python scripts/fullsize_staleness.py 2960 $F,$((F + (6 << 20))),$((F + (7 << 20))),$((F + (8 << 20))),$((F + (9 << 20))),$((F + (11 << 20))) r2s18_2960 2>&1 | grep -v Warning | tail -7
echo "== full size: kernel F, 1480 sentences in flight, hot = 64 / 256"
timeout 900 python scripts/fullsize_staleness.py 1480 $((F + (7 << 20))),$((F + (9 << 20))) r2s18_1480 2>&1 | grep -v Warning | tail -2
