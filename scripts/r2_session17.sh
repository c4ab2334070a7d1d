#!/bin/bash
# Round 2, GPU session 17 (1 GPU): kernel J (critical + helper warps, cp.async row ring) -- arithmetic test, throughput and
# full-size parity against kernel G (flags 0) with 4 / 3 / 2 stages.
mkdir -p gpurun_out
HW=524288
echo "== arithmetic test"; timeout 600 python -m pytest tests/test_sgns_gpu.py -m gpu -q -x --tb=short -k "arithmetic" 2>&1 | tail -8
echo "== full size: kernel G (0), kernel J with 4 / 3 / 2 stages"
timeout 900 python scripts/fullsize_staleness.py 0 0,$HW,$((HW + (3 << 12))),$((HW + (2 << 12))) r2s17 2>&1 | grep -v Warning | tail -6
