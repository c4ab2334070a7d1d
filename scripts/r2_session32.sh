#!/bin/bash
# Round 2, GPU session 32 (1 GPU): kernel F with the rows of the next unit requested into shared memory by cp.async
# (DGE_SGNS_F_ROW_PREFETCH_SMEM, 20 warps per SM kept): arithmetic tests, then A/B against the shipped schedule at full size.
mkdir -p gpurun_out
echo "== arithmetic test"; timeout 600 python -m pytest tests/test_sgns_gpu.py -m gpu -q --tb=line -k "arithmetic" 2>&1 | tail -6
F=$((2048 + (1 << 25) + (10 << 20)))
P=$((F + (1 << 26)))
timeout 900 python scripts/fullsize_staleness.py 2960,2960 $F,$P r2s32 2>&1 | grep -v Warning | tail -4
