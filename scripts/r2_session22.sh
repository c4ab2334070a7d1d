#!/bin/bash
# Round 2, GPU session 22 (1 GPU): kernel F with the rows of the next unit requested ahead (DGE_SGNS_F_ROW_PREFETCH):
# arithmetic tests, then throughput and full-size parity at 12 / 10 / 8 warps per SM with 128 / 256 write-through words.
mkdir -p gpurun_out
echo "== arithmetic test"; timeout 600 python -m pytest tests/test_sgns_gpu.py -m gpu -q -x --tb=short -k "arithmetic" 2>&1 | tail -8
F=$((2048 + (1 << 24)))
for w in 12 10 8; do
  timeout 900 python scripts/fullsize_staleness.py $((148 * w)) $((F + (w << 12) + (8 << 20))),$((F + (w << 12) + (9 << 20))) r2s22_w$w 2>&1 | grep -v Warning | tail -2
done
