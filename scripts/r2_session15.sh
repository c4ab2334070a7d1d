#!/bin/bash
# Round 2, GPU session 15 (1 GPU): kernel I (a warp per pair, dynamic hand-out of the round's pairs) -- arithmetic test,
# throughput and full-size parity against kernel G (flags 0) at 8 / 12 / 6 warps per block.
mkdir -p gpurun_out
PW=262144
echo "== arithmetic test"; timeout 600 python -m pytest tests/test_sgns_gpu.py -m gpu -q -x --tb=short -k "arithmetic or parallel_schedule" 2>&1 | tail -8
echo "== full size: kernel G (0), kernel I with 8 / 12 / 6 / 16 warps"
timeout 900 python scripts/fullsize_staleness.py 0 0,$PW,$((PW + (12 << 12))),$((PW + (6 << 12))),$((PW + (15 << 12))) r2s15 2>&1 | grep -v Warning | tail -8
echo "== full size: kernel I 8 warps, 148 / 222 sentences in flight"
timeout 600 python scripts/fullsize_staleness.py 148,222 $PW r2s15_conc 2>&1 | grep -v Warning | tail -3
