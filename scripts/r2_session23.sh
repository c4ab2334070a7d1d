#!/bin/bash
# Round 2, GPU session 23 (1 GPU): kernel F with the sentences handed out from a counter (DGE_SGNS_F_DYNAMIC): do the
# uneven configurations (10 / 13 / 14 warps per SM) agree with the oracle again?  128 write-through words.
mkdir -p gpurun_out
echo "== arithmetic test"; timeout 600 python -m pytest tests/test_sgns_gpu.py -m gpu -q --tb=line -k "arithmetic" 2>&1 | tail -8
F=$((2048 + (1 << 25) + (8 << 20)))
for w in 10 12 13 14 16; do
  timeout 900 python scripts/fullsize_staleness.py $((148 * w)) $((F + (w << 12))) r2s23_w$w 2>&1 | grep -v Warning | tail -1
done
P=$((F + (1 << 24)))
timeout 900 python scripts/fullsize_staleness.py 1480,1776 $((P + (10 << 12))),$((P + (12 << 12))) r2s23_pf 2>&1 | grep -v Warning | tail -4
